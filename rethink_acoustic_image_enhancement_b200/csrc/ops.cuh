// Host-side launchers of every kernel class (one per fused stage).  All launchers enqueue on
// `stream`, never allocate and never synchronise.  T is the storage type (float = reference-grade
// path, bf16 = throughput path); accumulation and statistics are always fp32.
#pragma once
#include "common.cuh"

namespace kd {

// Implicit-GEMM convolution: out[pixel, n] = epi( sum_{tap, c} A[shift(pixel, tap), c] * W[n, tap, c] )
struct ConvOp {
  const void* a0 = nullptr; int c0 = 0; long ld0 = 0;   // source 0: channels [0, c0)
  const void* a1 = nullptr; int c1 = 0; long ld1 = 0;   // source 1 (channel concat): [c0, c0+c1)
  int nimg = 1;                 // images * frames; rows = nimg*H*W
  int D = 1, H = 1, W = 1;      // frames per batch element (3-D convs), height, width
  int kd = 1, kh = 1, kw = 1;   // taps (1 or 3 each)
  int dil = 1;                  // spatial dilation (padding = dil*(k/2), zero fill)
  const void* w = nullptr;      // weights: element (n, tap, c) at w[n*w_ld + tap*w_tap_ld + c]
  long w_ld = 0, w_tap_ld = 0;
  // x-packed narrow conv (conv3x3_tc.cu): a0 is an NHWC tensor of xpack_cin (16 or 32) channels viewed as [.., W/P, 64] with
  // P = 64 / xpack_cin pixels per 128-byte "super-pixel"; W, c0 = 64, ld0 = 64 and epi.N = P * cout are given in super-pixel
  // units and w comes from pack_xconv (centre + halo tiles).  0 = ordinary conv.
  int xpack_cin = 0;
  int groups = 1;               // >1: per-image weights (w + g*w_group_stride), rows split evenly
  long w_group_stride = 0;
  Epilogue epi;
};

template <typename T> int conv_gemm_simt(const ConvOp& op, cudaStream_t s);
// tcgen05/TMEM/TMA implicit GEMM (bf16 storage). Returns -1 if the shape is not eligible.
int conv_gemm_tc(const ConvOp& op, cudaStream_t s);
bool conv_gemm_tc_eligible(const ConvOp& op);
// dense 3x3 / 3x3x3 (dilation 1) with a single halo-tile load per K chunk; N pre-split by conv_gemm_tc (conv3x3_tc.cu)
int conv3x3_tc(const ConvOp& op, int nc, int n_chunks, cudaStream_t s);
// dispatch: bf16 -> tcgen05 when eligible, SIMT otherwise; fp32 -> SIMT
template <typename T> int conv_gemm(const ConvOp& op, cudaStream_t s);

// per-pixel LayerNorm statistics over C channels: rstd = 1/sqrt(var+1e-5) (biased), mu optional
template <typename T> int ln_stats(const T* x, long ld, int C, long rows, float* rstd, float* mu, cudaStream_t s);

// depthwise 3x3 (pad 1) NHWC; w packed [9][C] (fp32).  gate=0: out[...,C]; gate=1: C = 2*hp,
// out[..., j] = gelu(dw(x)[j]) * dw(x)[hp + j], j < hp
template <typename T> int dwconv3x3(const T* x, long ldx, T* out, long ldo, const float* w9c, const float* bias,
                                    int nimg, int H, int W, int C, int gate, cudaStream_t s);
// fused LN-folded 1x1 conv (tcgen05) -> depthwise 3x3 on the CUDA cores (packed FFMA2) (-> GELU gate); w9c: fp32 [9][Nt]
bool pwdw_t_eligible(int C, int Nt, int gate);   // transposed schedule (pwdw_t.cu): depthwise inputs straight from TMEM
int pwdw_t(const bf16* x, long ldx, const float* rstd, const bf16* w1, int Nt, const float* w9c, bf16* out, long ldo, int nimg,
           int H, int W, int C, int gate, cudaStream_t s);
bool pwdw_f2_eligible(int C, int Nt, int gate);
// mu / s1 / s2 (optional, together): WithBias LayerNorm fold, t = rstd * acc - rstd * mu[p] * s1[n] + s2[n] (pack_ln_cols)
int pwdw_f2(const bf16* x, long ldx, const float* rstd, const bf16* w1, int Nt, const float* w9c, bf16* out, long ldo, int nimg,
            int H, int W, int C, int gate, cudaStream_t s, const float* mu = nullptr, const float* s1 = nullptr,
            const float* s2 = nullptr);

// MDTA reductions: partial Gram q k^T and squared norms per (image, head, split)
// qk: [nimg*HW, ld] with q at channel 0 and k at channel C.  part: [nimg][heads][splits][ch*ch + 2*ch] fp32
template <typename T> int mdta_gram(const T* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s);
int mdta_gram_splits(int HW, int nimg_heads);
// softmax(normalised Gram * temperature) folded into project_out: Mb[img][n][head*ch + j] = sum_i Wp[n][head*ch+i]*attn[i][j]
template <typename T> int mdta_fold(float* part /* split 0 is overwritten with the softmax */, int nimg, int C, int heads, int splits, const float* temperature,
                                    const float* wproj /*[C][C] fp32*/, T* mb, long mb_ld, long mb_img_stride, cudaStream_t s);

// direct convolutions for very small channel counts (CUDA cores, HBM-bound)
// few-in: cin <= 4 (input given as planes with explicit strides, optionally a second 1-plane source),
struct SmallConv {
  const float* in0 = nullptr; long in0_img = 0, in0_ch = 0; int cin0 = 0;   // fp32 planar (NCHW) source
  const float* in1 = nullptr; long in1_img = 0, in1_ch = 0; int cin1 = 0;   // concatenated second planar source
  int in1_px = 1;                                                           // its pixel stride: 0 = one value per image (broadcast)
  const float* sub0 = nullptr;                                              // optional: value = in0 - sub0 (same strides)
  int nimg = 1, D = 1, H = 1, W = 1, kd = 1, dil = 1;                       // 3x3 (kd=1) or 3x3x3 (kd=3) taps
  const float* w = nullptr;    // [taps][cin][cout] fp32
  const float* bias = nullptr; // [cout] or null
  int cout = 0, relu = 0;
  void* out = nullptr; long out_ld = 0;
};
template <typename T> int conv_few_in(const SmallConv& op, cudaStream_t s);
// few-out: cout <= 4, NHWC input of C channels, planar fp32 output (+ planar fp32 residual)
struct SmallConvOut {
  const void* in = nullptr; long in_ld = 0; int cin = 0;
  int nimg = 1, H = 1, W = 1, k = 3;      // k = 1 or 3
  const float* w = nullptr;    // [cout][taps][cin] fp32
  const float* bias = nullptr;
  int cout = 0;
  const float* res = nullptr; long res_img = 0, res_ch = 0;   // planar residual
  float* out = nullptr; long out_img = 0, out_ch = 0;          // planar output
};
template <typename T> int conv_few_out(const SmallConvOut& op, cudaStream_t s);
// out[img, co, y, x] = sum_{tap} part[img, co*9 + tap, y+dy-1, x+dx-1] (zero outside) (+ res): finishes a 3x3 conv whose nine
// per-tap channel contractions were computed as ONE 1x1 GEMM with 9*cout output columns (teacher.cu conv_to_planar)
int tap_sum(const float* part, int cout, int nimg, int H, int W, const float* res, long res_img, long res_ch, float* out, long out_img,
            long out_ch, cudaStream_t s);

// uint8 pre / post-processing around the teacher forward (prepost.cu; KDLAE_T.ipynb cell 5)
int preprocess_u8(const uint8_t* src, int B, int h, int w, int c, const float* rates, float* img, float* rate_map, int H, int W,
                  cudaStream_t s);
int postprocess_u8(const float* pred, const uint8_t* src, int B, int h, int w, int c, int Hp, int Wp, int scale, uint8_t* out,
                   cudaStream_t s);

// validation / training reductions behind the forward (metrics.cu; psnr_ssim.py:9-70, losses.py:135-194)
size_t psnr_scratch_bytes(int B);
int psnr_mse(const float* a, const float* b, int B, int C, int H, int W, int crop_border, int as_u8, double* out, void* scratch,
             cudaStream_t s);
size_t l1_sr_scratch_bytes();
int l1_shadow_term(const float* pred, const float* tgt, long n, float w_l1, float w_sh, int accumulate, float* loss, float* grad,
                   double* terms, void* scratch, cudaStream_t s);

// training-step slice (train.cu): GDFN half of a TransformerBlock with saves + backward (fp32), fused
// clip-norm + AdamW over flat buffers
size_t gdfn_train_ws_floats(int nimg, int H, int W, int C, int hp);
int gdfn_forward_train(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, float* out, int nimg,
                       int H, int W, int C, int hp, float* ws, cudaStream_t s);
int gdfn_backward(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, const float* dout,
                  float* dx, float* dgamma, float* dw_in, float* dw_dw, float* dw_out, int nimg, int H, int W, int C, int hp, float* ws,
                  cudaStream_t s);
size_t mdta_train_ws_floats(int nimg, int H, int W, int C, int heads);
int mdta_forward_train(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj, const float* temp,
                       float* out, int nimg, int H, int W, int C, int heads, float* ws, cudaStream_t s);
int mdta_backward(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj, const float* temp,
                  const float* dout, float* dx, float* dgamma, float* dw_qkv, float* dw_dw, float* dw_proj, float* dtemp, int nimg, int H,
                  int W, int C, int heads, float* ws, cudaStream_t s);
// TF32 tensor-core GEMM for plain fp32 1x1 convs (gemm_tf32.cu); the training step routes its forward / dgrad GEMMs here when
// train_matmul_tf32() is on (default off: fp32 CUDA-core kernels)
bool gemm_tf32_eligible(const ConvOp& op);
int gemm_tf32(const ConvOp& op, cudaStream_t s);
bool wgrad_tf32_eligible(const float* A, long lda, int N, const float* B, long ldb, int K);
int wgrad_tf32(const float* A, long lda, int N, const float* B, long ldb, int K, long P, float* part, int splits, int* splits_out,
               cudaStream_t s);
bool gram_tf32_eligible(const float* qk, long ld, int C, int heads);
int gram_tf32(const float* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s);
void set_train_matmul_tf32(int on);
int train_matmul_tf32();
// dense 1x1 / 3x3 conv (dilation, zero padding, no bias) with backward, fp32 NHWC, weights [Cout][k*k][Cin]
size_t conv_train_ws_floats(int nimg, int H, int W, int Cin, int Cout, int ks);
int conv_train_forward(const float* x, const float* w, float* out, int nimg, int H, int W, int Cin, int Cout, int ks, int dil,
                       cudaStream_t s);
int conv_train_backward(const float* x, const float* w, const float* dout, float* dx, float* dw, int nimg, int H, int W, int Cin, int Cout,
                        int ks, int dil, float* ws, cudaStream_t s);
int grad_norm_sq(const float* g, long n, double* out, double* scratch, cudaStream_t s);
int adamw_step(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps, float wd, int step,
               float max_norm, const double* norm_sq, cudaStream_t s);

// pooling / resampling / misc glue
template <typename T> int maxpool2x2(const T* x, T* out, int nimg, int H, int W, int C, cudaStream_t s);
template <typename T> int upsample_bilinear2x(const T* x, T* out, int nimg, int H, int W, int C, int OH, int OW, cudaStream_t s);
// score = tanh(MLP(outc(mean_pixels(feat)))): feat [nimg*HW][C] (C <= 64); outc_w fp32 [Cf][C] + outc_b [Cf] map the channel means
// to the Cf regressor inputs (GAP commuted through the 1x1 outc conv, ASDQE_model.py:72,168); w1 [256][Cf] ...
template <typename T> int gap_mlp_tanh(const T* feat, int nimg, int HW, int C, const float* outc_w, const float* outc_b, int Cf,
                                       const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                                       const float* b3, float* score, float* scratch, cudaStream_t s);
template <typename T> int nhwc_to_planar(const T* x, long ld, float* out, int nimg, int HW, int C, cudaStream_t s);

// weight packing (once per load_state_dict)
enum PackMode { PACK_PLAIN = 0, PACK_HALVES = 1, PACK_PIXEL_SHUFFLE = 2, PACK_CONVT = 3 };
struct PackOp {
  const float* src = nullptr;   // PyTorch layout [N][Cin][taps] (or ConvTranspose [Cin][Cout][4])
  int n_src = 0, c_src = 0, taps = 1;
  int mode = PACK_PLAIN;
  int h = 0, hp = 0;            // PACK_HALVES: source half size / padded half size
  int halves_on_k = 0;          // PACK_HALVES applied to K (dst k -> src k) instead of N
  const float* kscale = nullptr;   // [c_src]  (LayerNorm weight fold)
  const float* nscale = nullptr;   // [n_src]  (BatchNorm scale fold)
  void* dst = nullptr;          // [n_dst][taps][c_dst]
  int n_dst = 0, c_dst = 0;
};
template <typename T> int pack_weights(const PackOp& op, cudaStream_t s);
// WithBias-LN fold column vectors of a packed 1x1 weight: s1[n] = sum_k W'[n][k], s2[n] = sum_k lnb[k]*W[src(n)][k]
template <typename T> int pack_ln_cols(const PackOp& op, const float* lnb, float* s1, float* s2, cudaStream_t s);
// depthwise weights [C][1][3][3] -> [9][Cdst] fp32 (optionally with the HALVES channel map)
// x-packed 3x3 / 3x3x3 conv weights (see ConvOp::xpack_cin): src [cout][cin][kd*9] -> dst [P*cout][xconv_tiles][64] bf16
// (centre tiles: block-Toeplitz over the P pixels of a super-pixel; halo tiles: the neighbours' edge pixels), bias replicated P x
// dst [ng*c][taps][ng*c]: block g (rows and columns g*c .. g*c+c) <- src[g] [c][c][taps] * nscale[g][row]; zeros elsewhere.
// Three independent c -> c convs on the channel slices of one tensor become one dense conv (ASDQE stems, ASDQE_model.py:133-137)
template <typename T> int pack_blockdiag(const float* const* src, const float* const* nscale, int ng, int c, int taps, T* dst,
                                         cudaStream_t s);
int xconv_tiles(int kd, int cin);
int pack_xconv(const float* src, const float* nscale, int cout, int cin, int kd, bf16* dst, cudaStream_t s);
int pack_dw(const float* src, int c_src, int h, int hp, float* dst, int c_dst, cudaStream_t s);
int pack_dw_bias(const float* src, int c_src, int h, int hp, float* dst, int c_dst, cudaStream_t s);
// small-conv weights [Cout][Cin][taps] -> [taps][Cin][Cout] fp32 (few-in) ; -> [Cout][taps][Cin] (few-out)
int pack_few_in(const float* src, int cout, int cin, int taps, const float* nscale, float* dst, cudaStream_t s);
int pack_few_out(const float* src, int cout, int cin, int taps, float* dst, cudaStream_t s);
// BatchNorm(eval) fold: scale = g/sqrt(var+eps); shift = beta + (bias - mean)*scale
int bn_fold(const float* g, const float* beta, const float* mean, const float* var, const float* bias, int n, float eps,
            float* scale, float* shift, cudaStream_t s);
int copy_f32(const float* src, float* dst, long n, cudaStream_t s);

}  // namespace kd
