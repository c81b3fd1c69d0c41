/*
 * kdlae_b200.h - C ABI of the B200-native (sm_100a) KDLAE / ASDQE forward path.
 *
 * The reference (yangtaihong59/Rethink_Acoustic_Image_Enhancement) has no FFI layer: its hot path is
 * three Python nn.Modules.  This header is the boundary a binding (ctypes / cffi / pybind / cgo) links
 * against; each entry point names the reference interface it replaces.  Conventions:
 *
 *   - plain pointers and sizes only; every device pointer is caller-owned (torch / cudaMalloc);
 *   - no hidden allocation: packed weights and workspace are caller buffers sized by the *_bytes queries;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation;
 *   - return 0 on success, non-zero on error with text in kdlae_last_error() (thread-local);
 *   - `precision`: KDLAE_PREC_FP32 = fp32 storage + fp32 FFMA kernels (reference-grade, <=1e-4),
 *                  KDLAE_PREC_BF16 = NHWC bf16 storage, tcgen05/TMEM tensor-core convs, fp32 accumulate;
 *   - images cross the boundary in the reference's own layout: contiguous NCHW fp32.
 */
#ifndef KDLAE_B200_H_
#define KDLAE_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDLAE_ABI_VERSION 3
#define KDLAE_PREC_FP32 0
#define KDLAE_PREC_BF16 1

/* ---- library / device probes ---------------------------------------------------------------- */
int kdlae_abi_version(void);
const char* kdlae_last_error(void);
/* 0 if `device` is an sm_100 (B200) GPU this build can run on; fails cleanly (non-zero + message) otherwise. */
int kdlae_device_check(int device);
/* number of kernels launched by this library in the calling process since load (bench.py's gpu_launches) */
unsigned long long kdlae_launch_count(void);

/* ---- per-kernel-class profiler (CUDA events on the launching stream; used by bench.py's roofline) ------ */
int kdlae_profile_num_classes(void);
const char* kdlae_profile_class_name(int cls);
/* start recording one cudaEvent pair around every kernel launched by this library (adds a few us per launch) */
int kdlae_profile_begin(void);
/* stop, synchronise the device and return per class: summed kernel milliseconds, algorithmic FLOPs, algorithmic
 * bytes (inputs + outputs + weights of each launch) and the launch count. Arrays have kdlae_profile_num_classes() entries. */
int kdlae_profile_end(int n_classes, double* ms, double* flops, double* bytes, long long* launches);
/* per-launch records since kdlae_profile_begin(), in launch order (call BEFORE kdlae_profile_end): class index, kernel
 * milliseconds, algorithmic FLOPs and bytes of up to max_launches launches; returns the number written. */
int kdlae_profile_launches(int max_launches, int* cls, double* ms, double* flops, double* bytes);

/* ---- KDLAE-T : KDLAE_teacher (KDLAE/KDLAE_model.py:204-336), alias RestormerSuperResolutionParam2 ------ */
typedef struct kdlae_teacher_cfg {
  int inp_channels;          /* KDLAE_model.py:206 */
  int out_channels;          /* :207 */
  int dim;                   /* :208 (48) */
  int num_blocks[4];         /* :209 ([4,6,6,8]) */
  int num_refinement_blocks; /* :210 (4) */
  int heads[4];              /* :211 ([1,2,4,8]) */
  int hidden[4];             /* int(dim*2^l*ffn_expansion_factor) per level l, computed by the caller (:93) */
  int ln_with_bias;          /* LayerNorm_type != 'BiasFree' (:214) */
  int sr_head;               /* static == "train" (:262-267) */
  int params_cat;            /* params == 'cat' (:315) */
} kdlae_teacher_cfg;

/* number of state_dict tensors expected by kdlae_teacher_pack, in state_dict() order (483 for the shipped config) */
int kdlae_teacher_num_tensors(const kdlae_teacher_cfg* cfg);
size_t kdlae_teacher_packed_bytes(const kdlae_teacher_cfg* cfg, int precision);
/* Replaces nn.Module.load_state_dict-time weight placement: re-lays the fp32 state_dict tensors (device
 * pointers, state_dict() order, contiguous) into kernel operands: K-major bf16/fp32 GEMM weights with the
 * LayerNorm weight folded in, FFN halves padded (127->128...), PixelShuffle row order, depthwise [9][C]. */
int kdlae_teacher_pack(const kdlae_teacher_cfg* cfg, const float* const* tensors, int n_tensors, void* packed,
                       size_t packed_bytes, int precision, void* stream);
size_t kdlae_teacher_workspace_bytes(const kdlae_teacher_cfg* cfg, int micro_batch, int H, int W, int precision);
/* Replaces KDLAE_teacher.forward (KDLAE_model.py:270-336).
 *   img  [B, inp_channels, H, W] fp32, H % 8 == 0 and W % 8 == 0 (else error, like pixel_unshuffle raising)
 *   rate input["denoise_rate"]: rate_per_image == 0: the [B, 1, H, W] fp32 map the reference concatenates (:316);
 *        rate_per_image != 0: [B] fp32, one value per image, broadcast inside the dilated conv (the maps the notebook and the
 *        dataset build are constant per image - KDLAE_T.ipynb cell 5, paired_image_dataset.py:961 - so the H x W map never has to
 *        exist); may be NULL when params_cat == 0
 *   hq   [B, out_channels, H, W] fp32 out;  sr [B, out_channels, 2H, 2W] fp32 out (NULL iff sr_head == 0)
 * Images are processed `micro_batch` at a time inside the call (workspace is sized for micro_batch). */
int kdlae_teacher_forward(const kdlae_teacher_cfg* cfg, const void* packed, const float* img, const float* rate, int rate_per_image,
                          float* hq, float* sr, int B, int H, int W, int micro_batch, void* workspace, size_t workspace_bytes,
                          int precision, void* stream);

/* ---- KDLAE-S : KDLAE_student (KDLAE/KDLAE_model.py:340-430) ------------------------------------------ */
typedef struct kdlae_student_cfg {
  int hidden[3];   /* hidden_channels (:342), 2 levels + fusion: [16,32,64] */
  int residual;    /* :341 */
} kdlae_student_cfg;
size_t kdlae_student_packed_bytes(const kdlae_student_cfg* cfg, int precision);
int kdlae_student_pack(const kdlae_student_cfg* cfg, const float* const* tensors, int n_tensors /* 26 */, void* packed,
                       size_t packed_bytes, int precision, void* stream);
size_t kdlae_student_workspace_bytes(const kdlae_student_cfg* cfg, int micro_batch, int F, int H, int W, int precision);
/* Replaces KDLAE_student.forward (:395-430): x [B, F, H, W] fp32 -> y [B, F, H, W] fp32; H % 4 == 0, W % 4 == 0. */
int kdlae_student_forward(const kdlae_student_cfg* cfg, const void* packed, const float* x, float* y, int B, int F, int H, int W,
                          int micro_batch, void* workspace, size_t workspace_bytes, int precision, void* stream);

/* ---- ASDQE : DenoiseRatePredictor (ASDQE/ASDQE_model.py:123-171), eval mode --------------------------- */
typedef struct kdlae_asdqe_cfg {
  int in_channels; /* :127 (3) */
  int dim;         /* :127 (16) */
} kdlae_asdqe_cfg;
size_t kdlae_asdqe_packed_bytes(const kdlae_asdqe_cfg* cfg, int precision);
/* tensors: the 148 state_dict entries in order; num_batches_tracked entries are passed as NULL. */
int kdlae_asdqe_pack(const kdlae_asdqe_cfg* cfg, const float* const* tensors, int n_tensors, void* packed, size_t packed_bytes,
                     int precision, void* stream);
size_t kdlae_asdqe_workspace_bytes(const kdlae_asdqe_cfg* cfg, int micro_batch, int H, int W, int precision);
/* Replaces DenoiseRatePredictor.forward (:158-171): lq, gt [B, in_channels, H, W] fp32 -> score [B] fp32 in (-1,1).
 * feat (nullable): the U-Net output `enhanced_feat` (:167) as [B, 3*dim, Hp, Wp] fp32 with Hp/Wp = H/W padded to %16. */
int kdlae_asdqe_forward(const kdlae_asdqe_cfg* cfg, const void* packed, const float* lq, const float* gt, float* score,
                        float* feat, int B, int H, int W, int micro_batch, void* workspace, size_t workspace_bytes,
                        int precision, void* stream);

/* ---- single fused stages (unit parity tests and profiling; same kernels the forwards launch) ---------- */
/* Implicit-GEMM conv over NHWC activations: out[p, n] = act(rs[p] * sum_{tap,c} A[p+tap, c] W[n, tap, c] + bias[n]) (+ res).
 * Replaces nn.Conv2d 1x1 / 3x3 (KDLAE_model.py:95,99,118,120,186,196,238,243; ASDQE_model.py:24-31).
 * `a`,`w`,`res`,`out` are bf16 (precision 1) or fp32 (precision 0); w is [N][kh*kw][C] K-major. */
int kdlae_conv_gemm(const void* a, int C, const void* w, int N, int nimg, int H, int W, int ksize, const float* row_scale,
                    const float* col_bias, int relu, const void* res, void* out, int precision, int force_simt, void* stream);
/* MDTA reductions over all pixels of each image (KDLAE_model.py:134-137): for every (image, head)
 * gram[ch*ch + 2*ch] = { q k^T [ch][ch], |q_i|^2 [ch], |k_j|^2 [ch] }.  qk: [nimg*HW][ld] with q at channel 0 and k at
 * channel C.  scratch: kdlae_mdta_gram_scratch_floats() floats. bf16 path = tcgen05 with MN-major operands. */
size_t kdlae_mdta_gram_scratch_floats(int nimg, int HW, int C, int heads);
int kdlae_mdta_gram(const void* qk, long ld, int nimg, int HW, int C, int heads, float* gram, float* scratch, int precision,
                    void* stream);
/* Per-pixel channel LayerNorm statistics (KDLAE_model.py:50-52,:67-70). x: [rows][C]. */
int kdlae_ln_stats(const void* x, int C, long rows, float* rstd, float* mu, int precision, void* stream);
/* Depthwise 3x3 (+ optional GELU gate) (KDLAE_model.py:97,103-104,119). w9c fp32 [9][C]. */
int kdlae_dwconv3x3(const void* x, void* out, const float* w9c, int nimg, int H, int W, int C, int gate, int precision,
                    void* stream);

/* Fused LayerNorm-folded 1x1 conv (tcgen05) -> depthwise 3x3 on the CUDA cores (packed FFMA2) (-> GELU gate)
 * (KDLAE_model.py:95-104 / :118-119): out = dw3x3(rstd[p] * (x . w1^T)) [gate: gelu(.[:Nt/2]) * .[Nt/2:]].
 * x [nimg,H,W,C] bf16, w1 [Nt][C] bf16 (LayerNorm gamma folded), w9c [9][Nt] fp32, out [nimg,H,W,Nt or Nt/2] bf16.
 * pwdw_f2.cu: bit-identical to kdlae_conv_gemm + kdlae_dwconv3x3 (the t tile is rounded to bf16 in shared memory). */
int kdlae_pwdw_f2(const void* x, const float* rstd, const void* w1, int Nt, const float* w9c, void* out, int nimg, int H, int W, int C,
                  int gate, void* stream);

/* Same fused stage, transposed schedule (pwdw_t.cu): the 1x1 runs as W1 . X^T so the fp32 accumulator has lane = channel and
 * column = pixel, and the depthwise warps read their inputs straight from TMEM (t stays fp32, never leaves the SM). */
int kdlae_pwdw_t(const void* x, const float* rstd, const void* w1, int Nt, const float* w9c, void* out, int nimg, int H, int W, int C,
                 int gate, void* stream);

/* ---- uint8 pre / post-processing around the KDLAE-T forward (KDLAE/KDLAE_T.ipynb cell 5) ------------
 * pre:  src [B,h,w,c] uint8 (HWC) -> img [B,c,H,W] fp32 = src/255 with F.pad(...,'reflect') on the bottom/right edges
 *       (H >= h, W >= w, pad < size); rate_map (nullable) [B,1,H,W] = rates[b] (device floats).
 * post: pred [B,c,Hp,Wp] fp32 (Hp >= h*scale, Wp >= w*scale) -> clamp(0,1), crop to (h*scale, w*scale), rint(x*255)
 *       (skimage.img_as_ubyte), 0 wherever all c channels of src[b, y/scale, x/scale] are 0 -> out [B,h*scale,w*scale,c]. */
int kdlae_preprocess_u8(const unsigned char* src_hwc, int B, int h, int w, int c, const float* rates, float* img_nchw, float* rate_map,
                        int H, int W, void* stream);
int kdlae_postprocess_u8(const float* pred_nchw, const unsigned char* src_hwc, int B, int h, int w, int c, int Hp, int Wp, int scale,
                         unsigned char* out_hwc, void* stream);

/* ---- debug: stage-by-stage checksums of the KDLAE-T forward (two runs that should be bit-identical can be diffed) --------
 * begin: start recording; every stage of kdlae_teacher_forward then appends {sum, position-weighted sum} of its output words.
 * end:   synchronise, copy up to max_points pairs into `sums` (host, 2*max_points entries) and the newline-separated stage tags
 *        into `tags`; returns the number of points.  Costs one small kernel per stage while enabled; off by default. */
int kdlae_debug_trace_begin(void);
int kdlae_debug_trace_end(unsigned long long* sums, int max_points, char* tags, int tags_bytes);

/* ---- validation metric on device (Train/basicsr/metrics/psnr_ssim.py:9-70 calculate_psnr) -------------
 * img1, img2 [B,C,H,W] fp32 (device).  crop_border pixels are dropped on every edge (:57-59).  as_uint8 != 0 first converts
 * both images like tensor2img (utils/img_util.py:67-94: clamp(0,1), *255, round) - the `use_image` branch of
 * nondist_validation (image_restoration_model.py:327-334).  mse_max (device, [B][2] doubles) receives per image the mean
 * squared error over C x (H-2c) x (W-2c) and max(img1) (the reference picks max_value = 1 if img1.max() <= 1 else 255, :68-69);
 * the host finishes with 20*log10(max_value / sqrt(mse)).  scratch: kdlae_psnr_scratch_bytes(B) bytes of device memory. */
size_t kdlae_psnr_scratch_bytes(int B);
int kdlae_psnr(const float* img1, const float* img2, int B, int C, int H, int W, int crop_border, int as_uint8, double* mse_max,
               void* scratch, void* stream);

/* ---- training loss (Train/basicsr/models/losses/losses.py:135-194 L1LossSr, reduction='mean') --------
 * loss[0] = 0.5*lw*mean|hq-hq_gt| + 0.25*lw*mean|sr-sr_gt| + 0.25*lw*(mean|[hq>0.1]-[hq_gt>0.1]| + mean|[sr>0.1]-[sr_gt>0.1]|)
 * grad_hq / grad_sr (nullable) receive d loss / d pred = 0.5*lw*sign(hq-hq_gt)/n_hq and 0.25*lw*sign(sr-sr_gt)/n_sr (the
 * binarised "shadow" terms are piecewise constant: zero gradient, exactly as autograd sees torch.where).  sr may be NULL
 * (pred['sr'] is None, :166-171).  terms (nullable, device [4] doubles): the four means in the order above.
 * scratch: kdlae_l1_sr_scratch_bytes() bytes of device memory.  Sums are double, combined in a fixed order (deterministic). */
size_t kdlae_l1_sr_scratch_bytes(void);
int kdlae_l1_sr_loss(const float* hq, const float* hq_gt, long n_hq, const float* sr, const float* sr_gt, long n_sr,
                     float loss_weight, float* loss, float* grad_hq, float* grad_sr, double* terms, void* scratch, void* stream);

/* ---- training-step slice, fp32 reference-grade path -----------------------------------------------------
 * GDFN half of a TransformerBlock (KDLAE_model.py:50-52,101-106,163): out = x + project_out(gelu(u[:h]) * u[h:]),
 * u = dwconv3x3(project_in(BiasFree_LayerNorm(x))).  x, out, dout, dx: NHWC fp32 [nimg,H,W,C].  Weights in the packed layouts
 * of the forward path: gamma [C]; w_in [2hp][C] (chunk(2) halves each padded from h to hp rows); w_dw [9][2hp]; w_out [C][hp].
 * forward_train saves its intermediates in `ws` (kdlae_gdfn_train_ws_floats floats), backward consumes them and writes the
 * gradients in the same layouts.  Replaces autograd through FeedForward + LayerNorm (image_restoration_model.py:198-216). */
size_t kdlae_gdfn_train_ws_floats(int nimg, int H, int W, int C, int hp);
int kdlae_gdfn_forward_train(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, float* out,
                             int nimg, int H, int W, int C, int hp, float* ws, void* stream);
int kdlae_gdfn_backward(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, const float* dout,
                        float* dx, float* dgamma, float* dw_in, float* dw_dw, float* dw_out, int nimg, int H, int W, int C, int hp,
                        float* ws, void* stream);
/* MDTA half of a TransformerBlock (KDLAE_model.py:124-145,162): out = x + project_out(softmax(q^ k^T * temperature) v), BiasFree
 * LayerNorm in front.  gamma [C]; w_qkv [3C][C]; w_dw [9][3C]; w_proj [C][C]; temp [heads]; channels per head: multiple of 8, <= 96.
 * Same workspace contract as the GDFN pair; gradients come back in the same layouts (dtemp [heads]). */
size_t kdlae_mdta_train_ws_floats(int nimg, int H, int W, int C, int heads);
int kdlae_mdta_forward_train(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj,
                             const float* temp, float* out, int nimg, int H, int W, int C, int heads, float* ws, void* stream);
int kdlae_mdta_backward(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj, const float* temp,
                        const float* dout, float* dx, float* dgamma, float* dw_qkv, float* dw_dw, float* dw_proj, float* dtemp, int nimg,
                        int H, int W, int C, int heads, float* ws, void* stream);
/* Precision of the 1x1-conv GEMMs (forward, dgrad and wgrad) of the training entry points below: 0 (default) = fp32 on the CUDA cores,
 * 1 = TF32 on tcgen05 with fp32 accumulation (gemm_tf32.cu) - what torch.backends.cudnn.allow_tf32 (PyTorch's default) gives
 * the reference's nn.Conv2d layers on this GPU.  Process-wide; the initial value follows KDLAE_TRAIN_TF32=1.  The dense
 * 3x3 convs, the depthwise convs and every reduction stay fp32. */
int kdlae_set_train_matmul_tf32(int on);
int kdlae_train_matmul_tf32(void);
/* The dense convolutions of KDLAE-T outside its TransformerBlocks in training mode (OverlapPatchEmbed KDLAE_model.py:169-178, the
 * Downsample / Upsample bodies :182-200, reduce_chan_level*, output, output_param (dilation 2), output2, cen, upen, outputen
 * :239-268): 1x1 or 3x3, stride 1, zero padding dilation * (ksize / 2), no bias; fp32 NHWC, w [Cout][ksize*ksize][Cin].
 * backward: dx (nullable: the first layer needs none) [nimg,H,W,Cin], dw like w; ws: kdlae_conv_train_ws_floats() floats. */
size_t kdlae_conv_train_ws_floats(int nimg, int H, int W, int Cin, int Cout, int ksize);
int kdlae_conv_train_forward(const float* x, const float* w, float* out, int nimg, int H, int W, int Cin, int Cout, int ksize, int dilation,
                             void* stream);
int kdlae_conv_train_backward(const float* x, const float* w, const float* dout, float* dx, float* dw, int nimg, int H, int W, int Cin,
                              int Cout, int ksize, int dilation, float* ws, void* stream);
/* Fused torch.nn.utils.clip_grad_norm_(params, max_norm) + AdamW step (image_restoration_model.py:218-220) over flat fp32
 * buffers.  kdlae_grad_norm_sq: *norm_sq (device double) = sum g^2, scratch = 1024 device doubles.  kdlae_adamw_step: decoupled
 * weight decay, bias-corrected moments; the clip coefficient min(1, max_norm / (sqrt(*norm_sq) + 1e-6)) is applied on the fly
 * (norm_sq NULL or max_norm <= 0: no clipping).  step counts from 1. */
int kdlae_grad_norm_sq(const float* grad, long n, double* norm_sq, double* scratch, void* stream);
int kdlae_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int step, float max_norm, const double* norm_sq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KDLAE_B200_H_ */
