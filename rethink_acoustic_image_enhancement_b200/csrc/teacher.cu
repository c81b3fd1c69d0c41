// KDLAE-T (KDLAE_teacher, KDLAE/KDLAE_model.py:204-336): packed-weight layout and forward schedule.
//
// Activations are NHWC; the residual stream of each U-Net level stays resident in the workspace.
// A TransformerBlock (:159-163) is nine launches:
//   ln_stats -> [1x1 qkv GEMM, LN folded: weight into W, rstd as epilogue row scale] -> dw3x3
//   -> Gram/norm reduction -> softmax folded into project_out (per-image CxC) -> [1x1 GEMM on v + residual]
//   ln_stats -> [1x1 project_in GEMM, LN folded] -> dw3x3 + GELU gate -> [1x1 project_out GEMM + residual]
#include <cstdlib>
#include <type_traits>
#include <vector>
#include "models.cuh"

namespace kd {

namespace {

template <typename T>
struct BlockW {
  int C, heads, h, hp;
  T* wqkv; float* qkv_s1; float* qkv_s2;   // [3C][C] (+ WithBias column vectors)
  float* wdw_qkv;                          // [9][3C]
  float* wproj;                            // [C][C] fp32 (consumed by mdta_fold)
  float* temp;                             // [heads]
  T* win; float* in_s1; float* in_s2;      // [2hp][C]
  float* wdw_ffn;                          // [9][2hp]
  T* wout;                                 // [C][hp]
};

template <typename T>
struct TeacherW {
  float* patch_embed;                 // few-in [9][ic][dim]
  std::vector<BlockW<T>> enc1, enc2, enc3, latent, dec3, dec2, dec1, refine, refine_out, enhance;
  T *down1, *down2, *down3;           // [C/2][9][C]
  T *up4, *up3, *up2, *upen;          // [2C][9][C], rows packed sub-pixel major
  T *reduce3, *reduce2;               // 1x1 over the [upsampled, skip] concat
  float *output, *output_param, *output2, *cen, *outputen;
  T *output_tc, *output2_tc, *outputen_tc;   // bf16 path: [oc][9][C] for the tcgen05 3x3 kernel with the planar-fp32 epilogue
};

template <typename T>
void layout_block(Bump& b, BlockW<T>& w, int C, int heads, int hidden, bool lnb) {
  w.C = C; w.heads = heads; w.h = hidden; w.hp = (hidden + 7) / 8 * 8;
  w.wqkv = b.take<T>((size_t)3 * C * C);
  w.qkv_s1 = lnb ? b.take<float>(3 * C) : nullptr;
  w.qkv_s2 = lnb ? b.take<float>(3 * C) : nullptr;
  w.wdw_qkv = b.take<float>((size_t)9 * 3 * C);
  w.wproj = b.take<float>((size_t)C * C);
  w.temp = b.take<float>(heads);
  w.win = b.take<T>((size_t)2 * w.hp * C);
  w.in_s1 = lnb ? b.take<float>(2 * w.hp) : nullptr;
  w.in_s2 = lnb ? b.take<float>(2 * w.hp) : nullptr;
  w.wdw_ffn = b.take<float>((size_t)9 * 2 * w.hp);
  w.wout = b.take<T>((size_t)C * w.hp);
}

template <typename T>
void layout_teacher(const kdlae_teacher_cfg& c, Bump& b, TeacherW<T>& w) {
  const bool lnb = c.ln_with_bias != 0;
  const int d = c.dim, ic = c.inp_channels, oc = c.out_channels;
  auto blocks = [&](std::vector<BlockW<T>>& v, int n, int ch_level, int head_level) {
    v.resize(n);
    for (int i = 0; i < n; ++i) layout_block(b, v[i], d << ch_level, c.heads[head_level], c.hidden[ch_level], lnb);
  };
  w.patch_embed = b.take<float>((size_t)9 * ic * d);
  blocks(w.enc1, c.num_blocks[0], 0, 0);
  w.down1 = b.take<T>((size_t)(d / 2) * 9 * d);
  blocks(w.enc2, c.num_blocks[1], 1, 1);
  w.down2 = b.take<T>((size_t)d * 9 * 2 * d);
  blocks(w.enc3, c.num_blocks[2], 2, 2);
  w.down3 = b.take<T>((size_t)(2 * d) * 9 * 4 * d);
  blocks(w.latent, c.num_blocks[3], 3, 3);
  w.up4 = b.take<T>((size_t)(16 * d) * 9 * 8 * d);
  w.reduce3 = b.take<T>((size_t)(4 * d) * 8 * d);
  blocks(w.dec3, c.num_blocks[2], 2, 2);
  w.up3 = b.take<T>((size_t)(8 * d) * 9 * 4 * d);
  w.reduce2 = b.take<T>((size_t)(2 * d) * 4 * d);
  blocks(w.dec2, c.num_blocks[1], 1, 1);
  w.up2 = b.take<T>((size_t)(4 * d) * 9 * 2 * d);
  blocks(w.dec1, c.num_blocks[0], 1, 0);            // dim*2 channels with heads[0] (KDLAE_model.py:246)
  blocks(w.refine, c.num_refinement_blocks, 1, 0);
  const bool tc = std::is_same<T, bf16>::value;
  w.output = b.take<float>((size_t)oc * 9 * 2 * d);
  w.output_tc = tc ? b.take<T>((size_t)oc * 9 * 2 * d) : nullptr;
  w.output_param = b.take<float>((size_t)9 * (oc + 1) * 2 * d);
  blocks(w.refine_out, c.num_refinement_blocks, 1, 0);
  w.output2 = b.take<float>((size_t)oc * 9 * 2 * d);
  w.output2_tc = tc ? b.take<T>((size_t)oc * 9 * 2 * d) : nullptr;
  w.outputen_tc = nullptr;
  w.cen = w.outputen = nullptr;
  w.upen = nullptr;
  if (c.sr_head) {
    w.cen = b.take<float>((size_t)9 * oc * 2 * d);
    w.upen = b.take<T>((size_t)(4 * d) * 9 * 2 * d);
    blocks(w.enhance, c.num_refinement_blocks, 0, 0);
    w.outputen = b.take<float>((size_t)oc * 9 * d);
    w.outputen_tc = tc ? b.take<T>((size_t)oc * 9 * d) : nullptr;
  }
}

struct Cursor {
  const float* const* t; int n; int i;
  const float* next() { const float* p = (i < n) ? t[i] : nullptr; ++i; return p; }
};

template <typename T>
int pack_block(const BlockW<T>& w, Cursor& cur, bool lnb, cudaStream_t s) {
  const int C = w.C;
  const float* ln1w = cur.next(); const float* ln1b = lnb ? cur.next() : nullptr;
  const float* temp = cur.next();
  const float* qkv = cur.next(); const float* qkv_dw = cur.next(); const float* proj = cur.next();
  const float* ln2w = cur.next(); const float* ln2b = lnb ? cur.next() : nullptr;
  const float* pin = cur.next(); const float* dw = cur.next(); const float* pout = cur.next();
  KD_CHECK(pout != nullptr && ln1w != nullptr, "teacher_pack: missing tensors inside a TransformerBlock");
  PackOp p;
  p.src = qkv; p.n_src = 3 * C; p.c_src = C; p.taps = 1; p.kscale = ln1w; p.dst = w.wqkv; p.n_dst = 3 * C; p.c_dst = C;
  KD_TRY(pack_weights<T>(p, s));
  if (lnb) KD_TRY(pack_ln_cols<T>(p, ln1b, w.qkv_s1, w.qkv_s2, s));
  KD_TRY(pack_dw(qkv_dw, 3 * C, 0, 0, w.wdw_qkv, 3 * C, s));
  KD_TRY(copy_f32(proj, w.wproj, (long)C * C, s));
  KD_TRY(copy_f32(temp, w.temp, w.heads, s));
  // project_in: the two chunk(2) halves of h rows are each padded to hp (127 -> 128 ...) so both start 8-aligned
  p = PackOp();
  p.src = pin; p.n_src = 2 * w.h; p.c_src = C; p.taps = 1; p.mode = PACK_HALVES; p.h = w.h; p.hp = w.hp; p.kscale = ln2w;
  p.dst = w.win; p.n_dst = 2 * w.hp; p.c_dst = C;
  KD_TRY(pack_weights<T>(p, s));
  if (lnb) KD_TRY(pack_ln_cols<T>(p, ln2b, w.in_s1, w.in_s2, s));
  KD_TRY(pack_dw(dw, 2 * w.h, w.h, w.hp, w.wdw_ffn, 2 * w.hp, s));
  p = PackOp();
  p.src = pout; p.n_src = C; p.c_src = w.h; p.taps = 1; p.mode = PACK_HALVES; p.halves_on_k = 1; p.h = w.h; p.hp = w.hp;
  p.dst = w.wout; p.n_dst = C; p.c_dst = w.hp;
  KD_TRY(pack_weights<T>(p, s));
  return 0;
}

template <typename T>
int pack_conv3(const float* src, T* dst, int cout, int cin, int mode, cudaStream_t s) {
  KD_CHECK(src != nullptr, "teacher_pack: missing 3x3 conv weight");
  PackOp p;
  p.src = src; p.n_src = cout; p.c_src = cin; p.taps = 9; p.mode = mode; p.dst = dst; p.n_dst = cout; p.c_dst = cin;
  return pack_weights<T>(p, s);
}
template <typename T>
int pack_conv1(const float* src, T* dst, int cout, int cin, cudaStream_t s) {
  KD_CHECK(src != nullptr, "teacher_pack: missing 1x1 conv weight");
  PackOp p;
  p.src = src; p.n_src = cout; p.c_src = cin; p.taps = 1; p.dst = dst; p.n_dst = cout; p.c_dst = cin;
  return pack_weights<T>(p, s);
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <typename T>
struct Scratch {
  T* bufA; T* bufB;          // qkv / project_in output ; dwconv output / gated hidden
  float* rstd; float* mu;    // per-pixel LN statistics
  float* gram;               // Gram partials
  T* mb;                     // per-image folded attention matrices [nimg][C][C]
};

// One TransformerBlock (KDLAE_model.py:159-163) over `nimg` images of H x W pixels, C channels.
// x: residual stream (in place, row stride ldx); the block's result goes to xout (row stride ldo).
// KDLAE_FUSE_PWDW (read per forward) picks the schedule of the two "1x1 conv -> depthwise 3x3" pairs of a block:
//   0  unfused (conv_gemm + dwconv3x3);
//   3  both pairs through pwdw_f2.cu (tcgen05 1x1, packed-FFMA2 depthwise; bit-identical to unfused; stages with
//      C > 128 or a WithBias LayerNorm stay unfused), 4 only qkv, 5 only ffn;
//   6  (default) both pairs through pwdw_t.cu (transposed GEMM, depthwise inputs read from TMEM in fp32; with the 7 x 18 gate
//      tile it is the faster kernel for every pair: scripts/pw_bench.py), 7 qkv via pwdw_t + ffn via pwdw_f2, 8 the other way
//      round, 9 qkv via pwdw_t with the ffn pair unfused.
inline int fuse_pwdw_mode() {
  const char* e = getenv("KDLAE_FUSE_PWDW");
  return e ? atoi(e) : 6;
}
inline bool fuse_f2_qkv(int m) { return m == 3 || m == 4 || m == 8; }
inline bool fuse_f2_ffn(int m) { return m == 3 || m == 5 || m == 7; }
inline bool fuse_t_qkv(int m) { return m == 6 || m == 7 || m == 9; }
inline bool fuse_t_ffn(int m) { return m == 6 || m == 8; }

// In the bf16 path the GEMM that produces the residual stream also emits the LayerNorm statistics of its output
// rows (C <= 256, one accumulator chunk), so only the first norm1 of a stage needs the stand-alone ln_stats pass.
template <typename T>
bool epilogue_emits_stats(int C) { return std::is_same<T, bf16>::value && C % 8 == 0 && C <= 256; }

template <typename T>
int run_block(const BlockW<T>& w, bool lnb, T* x, long ldx, T* xout, long ldo, int nimg, int H, int W, Scratch<T>& sc,
              bool have_stats, bool emit_next_stats, cudaStream_t s) {
  const int C = w.C, HW = H * W;
  const long rows = (long)nimg * HW;
  const bool fused_stats = epilogue_emits_stats<T>(C);
  // ---- x = x + project_out(attn(norm1(x))) ----
  if (!(have_stats && fused_stats)) KD_TRY(ln_stats<T>(x, ldx, C, rows, sc.rstd, lnb ? sc.mu : nullptr, s));
  trace_point("blk.in", x, rows, (long)C * sizeof(T), ldx * sizeof(T), s);
  trace_point("blk.rstd1", sc.rstd, 1, rows * 4, rows * 4, s);
  // bf16 + BiasFree LayerNorm: the 1x1 conv is fused into the tensor-core depthwise kernel (t never reaches HBM)
  const int fmode = fuse_pwdw_mode();
  // WithBias LayerNorm (the constructor default): only pwdw_f2 carries the mean / bias fold, so every fused mode maps to it
  const bool f2ok = std::is_same<T, bf16>::value && pwdw_f2_eligible(C, 3 * C, 0) && pwdw_f2_eligible(C, 2 * w.hp, 1);
  const bool wb_f2 = lnb && fmode != 0;
  ConvOp g;
  if (f2ok && !lnb && fuse_t_qkv(fmode)) {
    KD_TRY(pwdw_t(reinterpret_cast<const bf16*>(x), ldx, sc.rstd, reinterpret_cast<const bf16*>(w.wqkv), 3 * C, w.wdw_qkv,
                  reinterpret_cast<bf16*>(sc.bufB), 3 * C, nimg, H, W, C, 0, s));
  } else if (f2ok && (fuse_f2_qkv(fmode) || wb_f2)) {
    KD_TRY(pwdw_f2(reinterpret_cast<const bf16*>(x), ldx, sc.rstd, reinterpret_cast<const bf16*>(w.wqkv), 3 * C, w.wdw_qkv,
                   reinterpret_cast<bf16*>(sc.bufB), 3 * C, nimg, H, W, C, 0, s, lnb ? sc.mu : nullptr, w.qkv_s1, w.qkv_s2));
  } else {
    g.a0 = x; g.c0 = C; g.ld0 = ldx; g.nimg = nimg; g.H = H; g.W = W;
    g.w = w.wqkv; g.w_ld = C; g.w_tap_ld = C;
    g.epi.row_scale = sc.rstd; g.epi.row_mu = lnb ? sc.mu : nullptr; g.epi.col_s1 = w.qkv_s1; g.epi.col_bias = w.qkv_s2;
    g.epi.out = sc.bufA; g.epi.out_ld = 3 * C; g.epi.N = 3 * C; g.epi.H = H; g.epi.W = W;
    KD_TRY(conv_gemm<T>(g, s));
    KD_TRY(dwconv3x3<T>(sc.bufA, 3 * C, sc.bufB, 3 * C, w.wdw_qkv, nullptr, nimg, H, W, 3 * C, 0, s));
  }
  const int splits = mdta_gram_splits(HW, nimg * w.heads);
  trace_point("blk.qkv_dw", sc.bufB, rows, 3L * C * sizeof(T), 3L * C * sizeof(T), s);
  KD_TRY(mdta_gram<T>(sc.bufB, 3 * C, nimg, HW, C, w.heads, splits, sc.gram, s));
  trace_point("blk.gram", sc.gram, 1, (long)nimg * w.heads * splits * ((C / w.heads) * (C / w.heads) + 2 * (C / w.heads)) * 4, 0, s);
  KD_TRY(mdta_fold<T>(sc.gram, nimg, C, w.heads, splits, w.temp, w.wproj, sc.mb, C, (long)C * C, s));
  trace_point("blk.mb", sc.mb, 1, (long)nimg * C * C * sizeof(T), 0, s);
  g = ConvOp();
  g.a0 = sc.bufB + 2 * C; g.c0 = C; g.ld0 = 3 * C; g.nimg = nimg; g.H = H; g.W = W;
  g.w = sc.mb; g.w_ld = C; g.w_tap_ld = C; g.groups = nimg; g.w_group_stride = (long)C * C;
  g.epi.res = x; g.epi.res_ld = ldx; g.epi.out = x; g.epi.out_ld = ldx; g.epi.N = C; g.epi.H = H; g.epi.W = W;
  if (fused_stats) { g.epi.stat_rstd = sc.rstd; g.epi.stat_mu = lnb ? sc.mu : nullptr; }
  KD_TRY(conv_gemm<T>(g, s));
  trace_point("blk.attn_out", x, rows, (long)C * sizeof(T), ldx * sizeof(T), s);
  // ---- x = x + ffn(norm2(x)) ----
  if (!fused_stats) KD_TRY(ln_stats<T>(x, ldx, C, rows, sc.rstd, lnb ? sc.mu : nullptr, s));
  trace_point("blk.rstd2", sc.rstd, 1, rows * 4, rows * 4, s);
  // (scripts/pw_bench.py, 8 x 512^2: 96 -> 2x256: 1008 us transposed vs 1270 us pwdw_f2; 48 -> 2x128: 590 vs 630 us)
  if (f2ok && !lnb && fuse_t_ffn(fmode)) {
    KD_TRY(pwdw_t(reinterpret_cast<const bf16*>(x), ldx, sc.rstd, reinterpret_cast<const bf16*>(w.win), 2 * w.hp, w.wdw_ffn,
                  reinterpret_cast<bf16*>(sc.bufB), w.hp, nimg, H, W, C, 1, s));
  } else if (f2ok && (fuse_f2_ffn(fmode) || wb_f2)) {
    KD_TRY(pwdw_f2(reinterpret_cast<const bf16*>(x), ldx, sc.rstd, reinterpret_cast<const bf16*>(w.win), 2 * w.hp, w.wdw_ffn,
                   reinterpret_cast<bf16*>(sc.bufB), w.hp, nimg, H, W, C, 1, s, lnb ? sc.mu : nullptr, w.in_s1, w.in_s2));
  } else {
    g = ConvOp();
    g.a0 = x; g.c0 = C; g.ld0 = ldx; g.nimg = nimg; g.H = H; g.W = W;
    g.w = w.win; g.w_ld = C; g.w_tap_ld = C;
    g.epi.row_scale = sc.rstd; g.epi.row_mu = lnb ? sc.mu : nullptr; g.epi.col_s1 = w.in_s1; g.epi.col_bias = w.in_s2;
    g.epi.out = sc.bufA; g.epi.out_ld = 2 * w.hp; g.epi.N = 2 * w.hp; g.epi.H = H; g.epi.W = W;
    KD_TRY(conv_gemm<T>(g, s));
    KD_TRY(dwconv3x3<T>(sc.bufA, 2 * w.hp, sc.bufB, w.hp, w.wdw_ffn, nullptr, nimg, H, W, 2 * w.hp, 1, s));
  }
  g = ConvOp();
  g.a0 = sc.bufB; g.c0 = w.hp; g.ld0 = w.hp; g.nimg = nimg; g.H = H; g.W = W;
  trace_point("blk.gated", sc.bufB, rows, (long)w.hp * sizeof(T), (long)w.hp * sizeof(T), s);
  g.w = w.wout; g.w_ld = w.hp; g.w_tap_ld = w.hp;
  g.epi.res = x; g.epi.res_ld = ldx; g.epi.out = xout; g.epi.out_ld = ldo; g.epi.N = C; g.epi.H = H; g.epi.W = W;
  if (fused_stats && emit_next_stats) { g.epi.stat_rstd = sc.rstd; g.epi.stat_mu = lnb ? sc.mu : nullptr; }
  KD_TRY(conv_gemm<T>(g, s));
  trace_point("blk.out", xout, rows, (long)C * sizeof(T), ldo * sizeof(T), s);
  return 0;
}

template <typename T>
int run_blocks(const std::vector<BlockW<T>>& v, bool lnb, T* x, long ldx, T* last_out, long last_ld, int nimg, int H, int W,
               Scratch<T>& sc, cudaStream_t s) {
  for (size_t i = 0; i < v.size(); ++i) {
    const bool last = (i + 1 == v.size());
    // stats for block i+1's norm1 come from block i's project_out epilogue (same stage, same pixel rows)
    KD_TRY(run_block<T>(v[i], lnb, x, ldx, last ? last_out : x, last ? last_ld : ldx, nimg, H, W, sc, /*have_stats=*/i > 0,
                        /*emit_next_stats=*/!last, s));
  }
  return 0;
}

// dense 3x3 conv (pad 1, no bias) as implicit GEMM with PixelShuffle / PixelUnshuffle addressing in the epilogue
template <typename T>
int conv3x3(const T* a, int cin, long lda, const T* w, int cout, int nimg, int H, int W, int mode, T* out, long ldo, int coff,
            cudaStream_t s) {
  ConvOp g;
  g.a0 = a; g.c0 = cin; g.ld0 = lda; g.nimg = nimg; g.H = H; g.W = W; g.kh = 3; g.kw = 3;
  g.w = w; g.w_ld = 9L * cin; g.w_tap_ld = cin;
  g.epi.out = out; g.epi.out_ld = ldo; g.epi.out_coff = coff; g.epi.N = cout; g.epi.H = H; g.epi.W = W; g.epi.mode = mode;
  g.epi.cq = cout / 4;
  return conv_gemm<T>(g, s);
}

// 3x3 conv to <= 4 planar fp32 output channels (+ planar residual): tcgen05 implicit GEMM (N padded to 16 inside the MMA)
// in the bf16 path, CUDA-core direct conv otherwise
template <typename T>
int conv_to_planar(const T* in, int cin, long ld, const float* w_few, const T* w_tc, int cout, int nimg, int H, int W,
                   const float* res, long res_img, long res_ch, float* out, long out_img, long out_ch, float* part, cudaStream_t s) {
  // bf16: the nine taps' channel contractions as ONE 1x1 tcgen05 GEMM with 9*cout output columns (w_tc [cout][9][cin] read as
  // [9*cout][cin]) into fp32 planes, then a 9-point gather-sum.  The implicit-GEMM form needs 9 N=16 MMAs per K step for a
  // single output channel and is bound by the tensor core's shared-memory operand reads (2.7 ms vs 0.5 ms at 16 x 1024^2).
  if (w_tc != nullptr && part != nullptr) {
    ConvOp g;
    g.a0 = in; g.c0 = cin; g.ld0 = ld; g.nimg = nimg; g.H = H; g.W = W;
    g.w = w_tc; g.w_ld = cin; g.w_tap_ld = cin;
    g.epi.mode = OUT_PLANAR_F32; g.epi.N = 9 * cout; g.epi.H = H; g.epi.W = W;
    g.epi.planar_out = part; g.epi.planar_img = 9L * cout * H * W; g.epi.planar_ch = (long)H * W;
    if (conv_gemm_tc_eligible(g)) {
      KD_TRY(conv_gemm_tc(g, s));
      return tap_sum(part, cout, nimg, H, W, res, res_img, res_ch, out, out_img, out_ch, s);
    }
  }
  if (w_tc != nullptr) {
    ConvOp g;
    g.a0 = in; g.c0 = cin; g.ld0 = ld; g.nimg = nimg; g.H = H; g.W = W; g.kh = 3; g.kw = 3;
    g.w = w_tc; g.w_ld = 9L * cin; g.w_tap_ld = cin;
    g.epi.mode = OUT_PLANAR_F32; g.epi.N = cout; g.epi.H = H; g.epi.W = W;
    g.epi.planar_out = out; g.epi.planar_img = out_img; g.epi.planar_ch = out_ch;
    g.epi.planar_res = res; g.epi.planar_res_img = res_img; g.epi.planar_res_ch = res_ch;
    if (conv_gemm_tc_eligible(g)) return conv_gemm_tc(g, s);
  }
  SmallConvOut fo;
  fo.in = in; fo.in_ld = ld; fo.cin = cin; fo.nimg = nimg; fo.H = H; fo.W = W; fo.k = 3; fo.w = w_few; fo.cout = cout;
  fo.res = res; fo.res_img = res_img; fo.res_ch = res_ch; fo.out = out; fo.out_img = out_img; fo.out_ch = out_ch;
  return conv_few_out<T>(fo, s);
}

struct WsLayout {
  size_t x1, d1, x2, x3, x4, d3, d2, s0, bufA, bufB, o1, rstd, mu, gram, mb, total;
};

template <typename T>
WsLayout ws_layout(const kdlae_teacher_cfg& c, int mb, int H, int W) {
  Bump b;
  WsLayout L;
  const size_t P1 = (size_t)mb * H * W, d = c.dim;
  const bool sr = c.sr_head != 0;
  auto off = [&](size_t elems, size_t esz) { b.off = align_up(b.off, 256); size_t o = b.off; b.off += elems * esz; return o; };
  const size_t hp1 = (c.hidden[0] + 7) / 8 * 8, hp2 = (c.hidden[1] + 7) / 8 * 8, hp3 = (c.hidden[2] + 7) / 8 * 8,
               hp4 = (c.hidden[3] + 7) / 8 * 8;
  L.x1 = off(P1 * d, sizeof(T));
  L.d1 = off(P1 * 2 * d, sizeof(T));
  L.x2 = off(P1 / 4 * 2 * d, sizeof(T));
  L.x3 = off(P1 / 16 * 4 * d, sizeof(T));
  L.x4 = off(P1 / 64 * 8 * d, sizeof(T));
  L.d3 = off(P1 / 16 * 4 * d, sizeof(T));
  L.d2 = off(P1 / 4 * 2 * d, sizeof(T));
  L.s0 = off(sr ? P1 * 4 * d : 0, sizeof(T));
  // bufA holds qkv (3C) or project_in output (2hp); bufB holds dwconv(qkv) (3C), the gated hidden (hp) or an upsampled map
  size_t a = 0, bb = 0;
  auto upd = [&](size_t pix, size_t C, size_t hp) {
    a = std::max(a, pix * std::max(3 * C, 2 * hp));
    bb = std::max(bb, pix * std::max(3 * C, hp));
  };
  upd(P1, d, hp1); upd(P1 / 4, 2 * d, hp2); upd(P1 / 16, 4 * d, hp3); upd(P1 / 64, 8 * d, hp4);
  upd(P1, 2 * d, hp2);
  if (sr) upd(P1 * 4, d, hp1);
  // conv_to_planar borrows bufA for the nine per-tap fp32 partial planes of the output convs (9 * oc floats per pixel, at
  // 2H x 2W for the SR head): a small-dim / many-output-channel config needs more than the qkv / project_in sizing above
  a = std::max(a, ((sr ? 4 : 1) * P1 * 9 * (size_t)c.out_channels * sizeof(float) + sizeof(T) - 1) / sizeof(T));
  L.bufA = off(a, sizeof(T));
  L.bufB = off(bb, sizeof(T));
  L.o1 = off(P1 * c.out_channels, sizeof(float));
  const size_t maxpix = sr ? P1 * 4 : P1;
  L.rstd = off(maxpix, sizeof(float));
  L.mu = off(c.ln_with_bias ? maxpix : 0, sizeof(float));
  // Gram partials: nimg * heads * splits * (ch*ch + 2ch); the largest case is a full-resolution stage
  size_t gmax = 0;
  auto gr = [&](size_t hw, size_t C, size_t heads) {
    const size_t ch = C / heads;
    gmax = std::max(gmax, (size_t)mb * heads * mdta_gram_splits((int)hw, 1) * (ch * ch + 2 * ch));
  };
  const size_t HW = (size_t)H * W;
  gr(HW, d, c.heads[0]); gr(HW / 4, 2 * d, c.heads[1]); gr(HW / 16, 4 * d, c.heads[2]); gr(HW / 64, 8 * d, c.heads[3]);
  gr(HW, 2 * d, c.heads[0]);
  if (sr) gr(HW * 4, d, c.heads[0]);
  L.gram = off(gmax, sizeof(float));
  L.mb = off((size_t)mb * 8 * d * 8 * d, sizeof(T));
  L.total = align_up(b.off, 256);
  return L;
}

}  // namespace

int teacher_num_tensors(const kdlae_teacher_cfg& c) {
  const int per_block = 9 + (c.ln_with_bias ? 2 : 0);
  int nblk = c.num_blocks[0] * 2 + c.num_blocks[1] * 2 + c.num_blocks[2] * 2 + c.num_blocks[3] + c.num_refinement_blocks * 2;
  int n = 1 + 3 + 3 + 2 + 3;  // patch_embed, downs, ups, reduce, output/output_param/output2
  if (c.sr_head) { nblk += c.num_refinement_blocks; n += 3; }
  return n + nblk * per_block;
}

template <typename T>
size_t teacher_packed_bytes(const kdlae_teacher_cfg& cfg) {
  Bump b;
  TeacherW<T> w;
  layout_teacher<T>(cfg, b, w);
  return align_up(b.off, 256);
}

template <typename T>
int teacher_pack(const kdlae_teacher_cfg& c, const float* const* tensors, int n_tensors, void* packed, size_t packed_bytes,
                 cudaStream_t s) {
  KD_CHECK(n_tensors == teacher_num_tensors(c), "teacher_pack: expected %d state_dict tensors, got %d", teacher_num_tensors(c),
           n_tensors);
  KD_CHECK(c.dim % 16 == 0, "teacher_pack: dim=%d must be a multiple of 16", c.dim);
  for (int l = 0; l < 4; ++l)
    KD_CHECK(c.heads[l] > 0 && (c.dim << l) % c.heads[l] == 0 && ((c.dim << l) / c.heads[l]) % 8 == 0,
             "teacher_pack: channels per head must be a multiple of 8 (level %d)", l);
  for (int l = 0; l < 4; ++l) KD_CHECK(c.num_blocks[l] >= 1, "teacher_pack: num_blocks[%d] must be >= 1", l);
  Bump b;
  b.base = reinterpret_cast<uint8_t*>(packed);
  TeacherW<T> w;
  layout_teacher<T>(c, b, w);
  KD_CHECK(b.off <= packed_bytes, "teacher_pack: packed buffer too small (%zu < %zu)", packed_bytes, b.off);
  const bool lnb = c.ln_with_bias != 0;
  const int d = c.dim, ic = c.inp_channels, oc = c.out_channels;
  Cursor cur{tensors, n_tensors, 0};
  auto blocks = [&](std::vector<BlockW<T>>& v) -> int {
    for (auto& bw : v) KD_TRY(pack_block<T>(bw, cur, lnb, s));
    return 0;
  };
  KD_TRY(pack_few_in(cur.next(), d, ic, 9, nullptr, w.patch_embed, s));
  KD_TRY(blocks(w.enc1));
  KD_TRY(pack_conv3<T>(cur.next(), w.down1, d / 2, d, PACK_PLAIN, s));
  KD_TRY(blocks(w.enc2));
  KD_TRY(pack_conv3<T>(cur.next(), w.down2, d, 2 * d, PACK_PLAIN, s));
  KD_TRY(blocks(w.enc3));
  KD_TRY(pack_conv3<T>(cur.next(), w.down3, 2 * d, 4 * d, PACK_PLAIN, s));
  KD_TRY(blocks(w.latent));
  KD_TRY(pack_conv3<T>(cur.next(), w.up4, 16 * d, 8 * d, PACK_PIXEL_SHUFFLE, s));
  KD_TRY(pack_conv1<T>(cur.next(), w.reduce3, 4 * d, 8 * d, s));
  KD_TRY(blocks(w.dec3));
  KD_TRY(pack_conv3<T>(cur.next(), w.up3, 8 * d, 4 * d, PACK_PIXEL_SHUFFLE, s));
  KD_TRY(pack_conv1<T>(cur.next(), w.reduce2, 2 * d, 4 * d, s));
  KD_TRY(blocks(w.dec2));
  KD_TRY(pack_conv3<T>(cur.next(), w.up2, 4 * d, 2 * d, PACK_PIXEL_SHUFFLE, s));
  KD_TRY(blocks(w.dec1));
  KD_TRY(blocks(w.refine));
  { const float* src = cur.next(); KD_TRY(pack_few_out(src, oc, 2 * d, 9, w.output, s)); if (w.output_tc) KD_TRY(pack_conv3<T>(src, w.output_tc, oc, 2 * d, PACK_PLAIN, s)); }
  KD_TRY(pack_few_in(cur.next(), 2 * d, oc + 1, 9, nullptr, w.output_param, s));
  KD_TRY(blocks(w.refine_out));
  { const float* src = cur.next(); KD_TRY(pack_few_out(src, oc, 2 * d, 9, w.output2, s)); if (w.output2_tc) KD_TRY(pack_conv3<T>(src, w.output2_tc, oc, 2 * d, PACK_PLAIN, s)); }
  if (c.sr_head) {
    KD_TRY(pack_few_in(cur.next(), 2 * d, oc, 9, nullptr, w.cen, s));
    KD_TRY(pack_conv3<T>(cur.next(), w.upen, 4 * d, 2 * d, PACK_PIXEL_SHUFFLE, s));
    KD_TRY(blocks(w.enhance));
    { const float* src = cur.next(); KD_TRY(pack_few_out(src, oc, d, 9, w.outputen, s)); if (w.outputen_tc) KD_TRY(pack_conv3<T>(src, w.outputen_tc, oc, d, PACK_PLAIN, s)); }
  }
  KD_CHECK(cur.i == n_tensors, "teacher_pack: consumed %d of %d tensors", cur.i, n_tensors);
  return 0;
}

template <typename T>
size_t teacher_workspace_bytes(const kdlae_teacher_cfg& cfg, int mb, int H, int W) {
  return ws_layout<T>(cfg, mb, H, W).total;
}

template <typename T>
int teacher_forward(const kdlae_teacher_cfg& c, const void* packed, const float* img, const float* rate, int rate_per_image,
                    float* hq, float* sr, int B, int H, int W, int micro_batch, void* ws, size_t ws_bytes, cudaStream_t s) {
  KD_CHECK(H > 0 && W > 0 && H % 8 == 0 && W % 8 == 0,
           "KDLAE_teacher: H and W must be multiples of 8 (got %dx%d) - pixel_unshuffle expects divisible sizes", H, W);
  KD_CHECK(B >= 1 && micro_batch >= 1, "KDLAE_teacher: bad batch %d / micro_batch %d", B, micro_batch);
  KD_CHECK(!c.params_cat || rate != nullptr, "KDLAE_teacher: denoise_rate is required when params == 'cat'");
  KD_CHECK((c.sr_head != 0) == (sr != nullptr), "KDLAE_teacher: sr buffer must be given iff static == 'train'");
  if (micro_batch > B) micro_batch = B;
  const WsLayout L = ws_layout<T>(c, micro_batch, H, W);
  KD_CHECK(ws_bytes >= L.total, "KDLAE_teacher: workspace too small (%zu < %zu)", ws_bytes, L.total);
  Bump b;
  b.base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(packed));
  TeacherW<T> w;
  layout_teacher<T>(c, b, w);

  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  T* x1 = reinterpret_cast<T*>(wsb + L.x1); T* d1 = reinterpret_cast<T*>(wsb + L.d1);
  T* x2 = reinterpret_cast<T*>(wsb + L.x2); T* x3 = reinterpret_cast<T*>(wsb + L.x3);
  T* x4 = reinterpret_cast<T*>(wsb + L.x4); T* d3 = reinterpret_cast<T*>(wsb + L.d3);
  T* d2 = reinterpret_cast<T*>(wsb + L.d2); T* s0 = reinterpret_cast<T*>(wsb + L.s0);
  float* o1 = reinterpret_cast<float*>(wsb + L.o1);
  Scratch<T> sc;
  sc.bufA = reinterpret_cast<T*>(wsb + L.bufA); sc.bufB = reinterpret_cast<T*>(wsb + L.bufB);
  sc.rstd = reinterpret_cast<float*>(wsb + L.rstd); sc.mu = reinterpret_cast<float*>(wsb + L.mu);
  sc.gram = reinterpret_cast<float*>(wsb + L.gram); sc.mb = reinterpret_cast<T*>(wsb + L.mb);

  const bool lnb = c.ln_with_bias != 0;
  const int d = c.dim, ic = c.inp_channels, oc = c.out_channels;
  const long HW = (long)H * W;

  for (int b0 = 0; b0 < B; b0 += micro_batch) {
    const int n = std::min(micro_batch, B - b0);
    const float* img_b = img + (long)b0 * ic * HW;
    const float* rate_b = rate ? rate + (rate_per_image ? (long)b0 : (long)b0 * HW) : nullptr;
    float* hq_b = hq + (long)b0 * oc * HW;
    float* sr_b = sr ? sr + (long)b0 * oc * HW * 4 : nullptr;

    // 1. patch_embed 3x3 ic -> dim (:173)
    SmallConv fi;
    fi.in0 = img_b; fi.in0_img = ic * HW; fi.in0_ch = HW; fi.cin0 = ic; fi.nimg = n; fi.H = H; fi.W = W;
    fi.w = w.patch_embed; fi.cout = d; fi.out = x1; fi.out_ld = d;
    KD_TRY(conv_few_in<T>(fi, s));
    trace_point("patch_embed", x1, (long)n * HW, (long)d * sizeof(T), (long)d * sizeof(T), s);
    // 2. encoder level 1; its last block writes the skip straight into the concat slot d1[..., d:2d]
    KD_TRY(run_blocks<T>(w.enc1, lnb, x1, d, d1 + d, 2 * d, n, H, W, sc, s));
    // 3. down1_2 (conv d -> d/2 + PixelUnshuffle) and level 2
    KD_TRY(conv3x3<T>(d1 + d, d, 2 * d, w.down1, d / 2, n, H, W, OUT_PIXEL_UNSHUFFLE, x2, 2 * d, 0, s));
    trace_point("down1", x2, (long)n * HW / 4, 2L * d * sizeof(T), 2L * d * sizeof(T), s);
    KD_TRY(run_blocks<T>(w.enc2, lnb, x2, 2 * d, x2, 2 * d, n, H / 2, W / 2, sc, s));
    KD_TRY(conv3x3<T>(x2, 2 * d, 2 * d, w.down2, d, n, H / 2, W / 2, OUT_PIXEL_UNSHUFFLE, x3, 4 * d, 0, s));
    trace_point("down2", x3, (long)n * HW / 16, 4L * d * sizeof(T), 4L * d * sizeof(T), s);
    KD_TRY(run_blocks<T>(w.enc3, lnb, x3, 4 * d, x3, 4 * d, n, H / 4, W / 4, sc, s));
    KD_TRY(conv3x3<T>(x3, 4 * d, 4 * d, w.down3, 2 * d, n, H / 4, W / 4, OUT_PIXEL_UNSHUFFLE, x4, 8 * d, 0, s));
    trace_point("down3", x4, (long)n * HW / 64, 8L * d * sizeof(T), 8L * d * sizeof(T), s);
    KD_TRY(run_blocks<T>(w.latent, lnb, x4, 8 * d, x4, 8 * d, n, H / 8, W / 8, sc, s));
    // 4. up4_3 (conv 8d -> 16d + PixelShuffle) -> cat with enc3 -> reduce_chan_level3 (dual-source 1x1)
    KD_TRY(conv3x3<T>(x4, 8 * d, 8 * d, w.up4, 16 * d, n, H / 8, W / 8, OUT_PIXEL_SHUFFLE, sc.bufB, 4 * d, 0, s));
    trace_point("up4", sc.bufB, (long)n * HW / 16, 4L * d * sizeof(T), 4L * d * sizeof(T), s);
    {
      ConvOp g;
      g.a0 = sc.bufB; g.c0 = 4 * d; g.ld0 = 4 * d; g.a1 = x3; g.c1 = 4 * d; g.ld1 = 4 * d;
      g.nimg = n; g.H = H / 4; g.W = W / 4; g.w = w.reduce3; g.w_ld = 8 * d; g.w_tap_ld = 8 * d;
      g.epi.out = d3; g.epi.out_ld = 4 * d; g.epi.N = 4 * d; g.epi.H = H / 4; g.epi.W = W / 4;
      KD_TRY(conv_gemm<T>(g, s));
    }
    KD_TRY(run_blocks<T>(w.dec3, lnb, d3, 4 * d, d3, 4 * d, n, H / 4, W / 4, sc, s));
    trace_point("dec3", d3, (long)n * HW / 16, 4L * d * sizeof(T), 4L * d * sizeof(T), s);
    KD_TRY(conv3x3<T>(d3, 4 * d, 4 * d, w.up3, 8 * d, n, H / 4, W / 4, OUT_PIXEL_SHUFFLE, sc.bufB, 2 * d, 0, s));
    trace_point("up3", sc.bufB, (long)n * HW / 4, 2L * d * sizeof(T), 2L * d * sizeof(T), s);
    {
      ConvOp g;
      g.a0 = sc.bufB; g.c0 = 2 * d; g.ld0 = 2 * d; g.a1 = x2; g.c1 = 2 * d; g.ld1 = 2 * d;
      g.nimg = n; g.H = H / 2; g.W = W / 2; g.w = w.reduce2; g.w_ld = 4 * d; g.w_tap_ld = 4 * d;
      g.epi.out = d2; g.epi.out_ld = 2 * d; g.epi.N = 2 * d; g.epi.H = H / 2; g.epi.W = W / 2;
      KD_TRY(conv_gemm<T>(g, s));
    }
    KD_TRY(run_blocks<T>(w.dec2, lnb, d2, 2 * d, d2, 2 * d, n, H / 2, W / 2, sc, s));
    // 5. up2_1 writes channels [0, d) of d1 (the skip already sits in [d, 2d)); no reduce conv at level 1 (:245)
    KD_TRY(conv3x3<T>(d2, 2 * d, 2 * d, w.up2, 4 * d, n, H / 2, W / 2, OUT_PIXEL_SHUFFLE, d1, 2 * d, 0, s));
    trace_point("up2+skip", d1, (long)n * HW, 2L * d * sizeof(T), 2L * d * sizeof(T), s);
    KD_TRY(run_blocks<T>(w.dec1, lnb, d1, 2 * d, d1, 2 * d, n, H, W, sc, s));
    KD_TRY(run_blocks<T>(w.refine, lnb, d1, 2 * d, d1, 2 * d, n, H, W, sc, s));
    // 6. output 3x3 2d -> oc ; denoise-rate tail (:314-321)
    KD_CHECK(ic == oc, "KDLAE_teacher: out + inp_img needs inp_channels == out_channels");
    const float* last_w = w.output; const T* last_w_tc = w.output_tc;
    if (c.params_cat) {
      KD_TRY(conv_to_planar<T>(d1, 2 * d, 2 * d, w.output, w.output_tc, oc, n, H, W, nullptr, 0, 0, o1, oc * HW, HW, reinterpret_cast<float*>(sc.bufA), s));
      fi = SmallConv();
      fi.in0 = o1; fi.in0_img = oc * HW; fi.in0_ch = HW; fi.cin0 = oc;
      fi.in1 = rate_b; fi.in1_img = rate_per_image ? 1 : HW; fi.in1_ch = 0; fi.cin1 = 1; fi.in1_px = rate_per_image ? 0 : 1;
      fi.nimg = n; fi.H = H; fi.W = W; fi.dil = 2; fi.w = w.output_param; fi.cout = 2 * d; fi.out = d1; fi.out_ld = 2 * d;
      KD_TRY(conv_few_in<T>(fi, s));
      trace_point("output", o1, 1, (long)n * oc * HW * 4, 0, s);
      trace_point("output_param", d1, (long)n * HW, 2L * d * sizeof(T), 2L * d * sizeof(T), s);
      KD_TRY(run_blocks<T>(w.refine_out, lnb, d1, 2 * d, d1, 2 * d, n, H, W, sc, s));
      last_w = w.output2; last_w_tc = w.output2_tc;
    }
    // out_hq = out + inp_img (:321)
    KD_TRY(conv_to_planar<T>(d1, 2 * d, 2 * d, last_w, last_w_tc, oc, n, H, W, img_b, ic * HW, HW, hq_b, oc * HW, HW, reinterpret_cast<float*>(sc.bufA), s));
    trace_point("hq", hq_b, 1, (long)n * oc * HW * 4, 0, s);
    // 7. SR head (:324-329): cen -> upen (PixelShuffle) -> enhance -> outputen
    if (c.sr_head) {
      fi = SmallConv();
      fi.in0 = hq_b; fi.in0_img = oc * HW; fi.in0_ch = HW; fi.cin0 = oc; fi.nimg = n; fi.H = H; fi.W = W;
      fi.w = w.cen; fi.cout = 2 * d; fi.out = d1; fi.out_ld = 2 * d;
      KD_TRY(conv_few_in<T>(fi, s));
      trace_point("cen", d1, (long)n * HW, 2L * d * sizeof(T), 2L * d * sizeof(T), s);
      KD_TRY(conv3x3<T>(d1, 2 * d, 2 * d, w.upen, 4 * d, n, H, W, OUT_PIXEL_SHUFFLE, s0, d, 0, s));
      trace_point("upen", s0, (long)n * HW * 4, (long)d * sizeof(T), (long)d * sizeof(T), s);
      KD_TRY(run_blocks<T>(w.enhance, lnb, s0, d, s0, d, n, 2 * H, 2 * W, sc, s));
      KD_TRY(conv_to_planar<T>(s0, d, d, w.outputen, w.outputen_tc, oc, n, 2 * H, 2 * W, nullptr, 0, 0, sr_b, oc * HW * 4, HW * 4,
                               reinterpret_cast<float*>(sc.bufA), s));
      trace_point("sr", sr_b, 1, (long)n * oc * HW * 4 * 4, 0, s);
    }
  }
  return 0;
}

#define INST(T)                                                                                                              \
  template size_t teacher_packed_bytes<T>(const kdlae_teacher_cfg&);                                                         \
  template int teacher_pack<T>(const kdlae_teacher_cfg&, const float* const*, int, void*, size_t, cudaStream_t);              \
  template size_t teacher_workspace_bytes<T>(const kdlae_teacher_cfg&, int, int, int);                                       \
  template int teacher_forward<T>(const kdlae_teacher_cfg&, const void*, const float*, const float*, int, float*, float*, int, \
                                  int, int, int, void*, size_t, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace kd
