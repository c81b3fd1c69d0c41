// ASDQE DenoiseRatePredictor (ASDQE/ASDQE_model.py:123-171), eval mode.
// Every Conv3x3+BatchNorm+ReLU stage is one implicit GEMM: BN scale folded into the weight rows,
// BN shift (+conv bias) as the epilogue column bias, ReLU in the epilogue.  torch.cat([x2, x1]) in
// `Up` (:65) and the three-stem concat (:166) never materialise: they are dual-source K loops /
// channel-offset stores.  MaxPool2d(2) and the bilinear x2 (align_corners=True) are vectorised passes.
#include <algorithm>
#include "models.cuh"

namespace kd {

namespace {

template <typename T>
struct CBR { T* w; float* scale; float* shift; int cin, cout; };   // conv3x3 + BN + ReLU

template <typename T>
struct AsdqeW {
  struct Stem { float* w1; float* scale1; float* shift1; CBR<T> c2; } stem[3];
  CBR<T> stem2;                             // the three stems' second convs as ONE block-diagonal 3*dim -> 3*dim conv
  CBR<T> inc[2], d1[2], d2[2], d3[2], u1[2], u2[2], u3[2];
  T* outc; float* outc_b; float* outc_f;    // outc_f: fp32 [3*dim][64] for the GAP-commuted score path
  float *w1, *b1, *w2, *b2, *w3, *b3;
};

template <typename T>
void layout_asdqe(const kdlae_asdqe_cfg& c, Bump& b, AsdqeW<T>& w) {
  const int ic = c.in_channels, dm = c.dim, cc = 3 * dm;
  auto cbr = [&](CBR<T>& k, int cin, int cout) {
    k.cin = cin; k.cout = cout;
    k.w = b.take<T>((size_t)cout * 9 * cin);
    k.scale = b.take<float>(cout);
    k.shift = b.take<float>(cout);
  };
  for (int i = 0; i < 3; ++i) {
    w.stem[i].w1 = b.take<float>((size_t)9 * ic * dm);
    w.stem[i].scale1 = b.take<float>(dm);
    w.stem[i].shift1 = b.take<float>(dm);
    cbr(w.stem[i].c2, dm, dm);
  }
  cbr(w.stem2, cc, cc);
  cbr(w.inc[0], cc, 64); cbr(w.inc[1], 64, 64);
  cbr(w.d1[0], 64, 128); cbr(w.d1[1], 128, 128);
  cbr(w.d2[0], 128, 256); cbr(w.d2[1], 256, 256);
  cbr(w.d3[0], 256, 256); cbr(w.d3[1], 256, 256);
  cbr(w.u1[0], 512, 128); cbr(w.u1[1], 128, 128);
  cbr(w.u2[0], 256, 64); cbr(w.u2[1], 64, 64);
  cbr(w.u3[0], 128, 64); cbr(w.u3[1], 64, 64);
  w.outc = b.take<T>((size_t)cc * 64);
  w.outc_b = b.take<float>(cc);
  w.outc_f = b.take<float>((size_t)cc * 64);
  w.w1 = b.take<float>((size_t)256 * cc); w.b1 = b.take<float>(256);
  w.w2 = b.take<float>(64 * 256); w.b2 = b.take<float>(64);
  w.w3 = b.take<float>(64); w.b3 = b.take<float>(1);
}

struct Cur {
  const float* const* t; int n; int i;
  const float* next() { const float* p = (i < n) ? t[i] : nullptr; ++i; return p; }
};

// consumes conv.{weight,bias}, bn.{weight,bias,running_mean,running_var,num_batches_tracked}
template <typename T>
int pack_cbr(const CBR<T>& k, Cur& cur, cudaStream_t s) {
  const float* cw = cur.next(); const float* cb = cur.next();
  const float* g = cur.next(); const float* beta = cur.next(); const float* mean = cur.next(); const float* var = cur.next();
  cur.next();  // num_batches_tracked (unused in eval)
  KD_CHECK(cw && cb && g && beta && mean && var, "asdqe_pack: missing conv/BN tensor");
  KD_TRY(bn_fold(g, beta, mean, var, cb, k.cout, 1e-5f, k.scale, k.shift, s));
  PackOp p;
  p.src = cw; p.n_src = k.cout; p.c_src = k.cin; p.taps = 9; p.nscale = k.scale; p.dst = k.w; p.n_dst = k.cout; p.c_dst = k.cin;
  return pack_weights<T>(p, s);
}

template <typename T>
int cbr_run(const CBR<T>& k, const T* a0, int c0, long ld0, const T* a1, int c1, long ld1, T* out, long ldo, int coff, int nimg,
            int H, int W, cudaStream_t s) {
  ConvOp g;
  g.a0 = a0; g.c0 = c0; g.ld0 = ld0; g.a1 = a1; g.c1 = c1; g.ld1 = ld1; g.nimg = nimg; g.H = H; g.W = W; g.kh = g.kw = 3;
  g.w = k.w; g.w_ld = 9L * k.cin; g.w_tap_ld = k.cin;
  g.epi.col_bias = k.shift; g.epi.relu = 1; g.epi.out = out; g.epi.out_ld = ldo; g.epi.out_coff = coff; g.epi.N = k.cout;
  g.epi.H = H; g.epi.W = W;
  return conv_gemm<T>(g, s);
}

struct AWs { size_t f48, a[4], b[4], c[4], gap, total; };
template <typename T>
AWs aws_layout(const kdlae_asdqe_cfg& cfg, int mb, int Hp, int Wp) {
  Bump b;
  AWs L;
  auto off = [&](size_t bytes) { b.off = align_up(b.off, 256); size_t o = b.off; b.off += bytes; return o; };
  const size_t P = (size_t)mb * Hp * Wp;
  L.f48 = off(P * 3 * cfg.dim * sizeof(T));
  const size_t chans[4] = {64, 128, 256, 256};
  for (int l = 0; l < 4; ++l) {
    const size_t e = (P >> (2 * l)) * chans[l] * sizeof(T);
    L.a[l] = off(e); L.b[l] = off(e); L.c[l] = off(e);
  }
  L.gap = off((size_t)mb * 64 * 64 * sizeof(float));
  L.total = align_up(b.off, 256);
  return L;
}

inline int pad16(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

template <typename T>
size_t asdqe_packed_bytes(const kdlae_asdqe_cfg& cfg) {
  Bump b;
  AsdqeW<T> w;
  layout_asdqe<T>(cfg, b, w);
  return align_up(b.off, 256);
}

template <typename T>
int asdqe_pack(const kdlae_asdqe_cfg& c, const float* const* t, int n_tensors, void* packed, size_t packed_bytes, cudaStream_t s) {
  KD_CHECK(n_tensors == 148, "asdqe_pack: expected 148 state_dict entries, got %d", n_tensors);
  KD_CHECK(c.dim % 8 == 0 && c.in_channels >= 1 && c.in_channels <= 4 && 3 * c.dim <= 64,
           "asdqe_pack: unsupported in_channels=%d dim=%d", c.in_channels, c.dim);
  Bump b;
  b.base = reinterpret_cast<uint8_t*>(packed);
  AsdqeW<T> w;
  layout_asdqe<T>(c, b, w);
  KD_CHECK(b.off <= packed_bytes, "asdqe_pack: packed buffer too small");
  Cur cur{t, n_tensors, 0};
  const float* c2w[3];
  const float* c2s[3];
  for (int i = 0; i < 3; ++i) {  // lq_extractor, gt_extractor, diff_extractor (:133-137)
    const float* cw = cur.next(); const float* cb = cur.next();
    const float* g = cur.next(); const float* beta = cur.next(); const float* mean = cur.next(); const float* var = cur.next();
    cur.next();
    KD_CHECK(cw && cb && g && beta && mean && var, "asdqe_pack: missing stem tensor");
    KD_TRY(bn_fold(g, beta, mean, var, cb, c.dim, 1e-5f, w.stem[i].scale1, w.stem[i].shift1, s));
    KD_TRY(pack_few_in(cw, c.dim, c.in_channels, 9, w.stem[i].scale1, w.stem[i].w1, s));
    c2w[i] = (cur.i < cur.n) ? cur.t[cur.i] : nullptr;      // conv weight of the stem's second conv (consumed by pack_cbr below)
    KD_TRY(pack_cbr<T>(w.stem[i].c2, cur, s));
    c2s[i] = w.stem[i].c2.scale;
    KD_TRY(copy_f32(w.stem[i].c2.shift, w.stem2.shift + i * c.dim, c.dim, s));
  }
  KD_TRY(pack_blockdiag<T>(c2w, c2s, 3, c.dim, 9, w.stem2.w, s));
  CBR<T>* order[7] = {w.inc, w.d1, w.d2, w.d3, w.u1, w.u2, w.u3};
  for (int i = 0; i < 7; ++i) {
    KD_TRY(pack_cbr<T>(order[i][0], cur, s));
    KD_TRY(pack_cbr<T>(order[i][1], cur, s));
  }
  const int cc = 3 * c.dim;
  {
    const float* ow = cur.next(); const float* ob = cur.next();
    KD_CHECK(ow && ob, "asdqe_pack: missing outc tensor");
    PackOp p;
    p.src = ow; p.n_src = cc; p.c_src = 64; p.taps = 1; p.dst = w.outc; p.n_dst = cc; p.c_dst = 64;
    KD_TRY(pack_weights<T>(p, s));
    KD_TRY(copy_f32(ob, w.outc_b, cc, s));
    KD_TRY(copy_f32(ow, w.outc_f, (long)cc * 64, s));
  }
  const float* r[6];
  for (int i = 0; i < 6; ++i) { r[i] = cur.next(); KD_CHECK(r[i], "asdqe_pack: missing regressor tensor"); }
  KD_TRY(copy_f32(r[0], w.w1, 256L * cc, s)); KD_TRY(copy_f32(r[1], w.b1, 256, s));
  KD_TRY(copy_f32(r[2], w.w2, 64 * 256, s)); KD_TRY(copy_f32(r[3], w.b2, 64, s));
  KD_TRY(copy_f32(r[4], w.w3, 64, s)); KD_TRY(copy_f32(r[5], w.b3, 1, s));
  KD_CHECK(cur.i == n_tensors, "asdqe_pack: consumed %d of %d entries", cur.i, n_tensors);
  return 0;
}

template <typename T>
size_t asdqe_workspace_bytes(const kdlae_asdqe_cfg& cfg, int mb, int H, int W) {
  return aws_layout<T>(cfg, mb, pad16(H, cfg.dim), pad16(W, cfg.dim)).total;
}

template <typename T>
int asdqe_forward(const kdlae_asdqe_cfg& c, const void* packed, const float* lq, const float* gt, float* score, float* feat, int B,
                  int H, int W, int micro_batch, void* ws, size_t ws_bytes, cudaStream_t s) {
  KD_CHECK(B >= 1 && H >= 1 && W >= 1 && micro_batch >= 1, "DenoiseRatePredictor: bad shape");
  KD_CHECK(c.dim % 8 == 0 && c.dim >= 8, "DenoiseRatePredictor: dim must be a multiple of 8");
  if (micro_batch > B) micro_batch = B;
  const int Hp = pad16(H, c.dim), Wp = pad16(W, c.dim);   // pad_to_multiple(x, dim) (:113-121,:159-160)
  KD_CHECK(Hp % 8 == 0 && Wp % 8 == 0, "DenoiseRatePredictor: padded size must be divisible by 8");
  const AWs L = aws_layout<T>(c, micro_batch, Hp, Wp);
  KD_CHECK(ws_bytes >= L.total, "DenoiseRatePredictor: workspace too small (%zu < %zu)", ws_bytes, L.total);
  Bump bp;
  bp.base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(packed));
  AsdqeW<T> w;
  layout_asdqe<T>(c, bp, w);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  T* F48 = reinterpret_cast<T*>(wsb + L.f48);
  T *a[4], *b[4], *cb[4];
  for (int l = 0; l < 4; ++l) {
    a[l] = reinterpret_cast<T*>(wsb + L.a[l]); b[l] = reinterpret_cast<T*>(wsb + L.b[l]); cb[l] = reinterpret_cast<T*>(wsb + L.c[l]);
  }
  float* gap = reinterpret_cast<float*>(wsb + L.gap);
  const int ic = c.in_channels, dm = c.dim, cc = 3 * dm;
  const long HWin = (long)H * W, HWp = (long)Hp * Wp;

  for (int b0 = 0; b0 < B; b0 += micro_batch) {
    const int n = std::min(micro_batch, B - b0);
    const float* lqb = lq + (long)b0 * ic * HWin;
    const float* gtb = gt + (long)b0 * ic * HWin;
    // stems on lq, gt, lq - gt (:162-164); zero padding to (Hp, Wp) = out-of-range reads return 0
    for (int i = 0; i < 3; ++i) {
      SmallConv fi;
      fi.in0 = (i == 1) ? gtb : lqb; fi.sub0 = (i == 2) ? gtb : nullptr;
      fi.in0_img = ic * HWin; fi.in0_ch = HWin; fi.cin0 = ic; fi.nimg = n; fi.H = Hp; fi.W = Wp;
      fi.w = w.stem[i].w1; fi.bias = w.stem[i].shift1; fi.cout = dm; fi.relu = 1; fi.out = b[0] + i * dm; fi.out_ld = cc;
      KD_TRY(conv_few_in_sized<T>(fi, H, W, s));
    }
    // second conv of the three stems: one dense 3*dim -> 3*dim conv with block-diagonal weights over the channel slices
    // (3x the useful FLOPs on a layer that was bound by per-tile latency, not by the tensor pipe: 3 x 240 us -> one launch)
    KD_TRY(cbr_run<T>(w.stem2, b[0], cc, cc, nullptr, 0, 0, F48, cc, 0, n, Hp, Wp, s));
    // U-Net encoder (:97-100)
    KD_TRY(cbr_run<T>(w.inc[0], F48, cc, cc, nullptr, 0, 0, b[0], 64, 0, n, Hp, Wp, s));
    KD_TRY(cbr_run<T>(w.inc[1], b[0], 64, 64, nullptr, 0, 0, a[0], 64, 0, n, Hp, Wp, s));
    const int chans[4] = {64, 128, 256, 256};
    CBR<T>* down[3] = {w.d1, w.d2, w.d3};
    for (int l = 1; l <= 3; ++l) {
      const int h = Hp >> l, wd = Wp >> l;
      KD_TRY(maxpool2x2<T>(a[l - 1], b[l], n, h * 2, wd * 2, chans[l - 1], s));
      KD_TRY(cbr_run<T>(down[l - 1][0], b[l], chans[l - 1], chans[l - 1], nullptr, 0, 0, cb[l], chans[l], 0, n, h, wd, s));
      KD_TRY(cbr_run<T>(down[l - 1][1], cb[l], chans[l], chans[l], nullptr, 0, 0, a[l], chans[l], 0, n, h, wd, s));
    }
    // decoder: bilinear x2 (align_corners) then conv over cat([skip, upsampled]) (:60-66)
    CBR<T>* up[3] = {w.u1, w.u2, w.u3};
    const T* cur = a[3];
    int cur_c = 256;
    for (int j = 0; j < 3; ++j) {
      const int l = 2 - j;                      // target level
      const int h = Hp >> l, wd = Wp >> l;
      KD_TRY(upsample_bilinear2x<T>(cur, b[l], n, h / 2, wd / 2, cur_c, h, wd, s));
      KD_TRY(cbr_run<T>(up[j][0], a[l], chans[l], chans[l], b[l], cur_c, cur_c, cb[l], up[j][0].cout, 0, n, h, wd, s));
      KD_TRY(cbr_run<T>(up[j][1], cb[l], up[j][0].cout, up[j][0].cout, nullptr, 0, 0, b[l], up[j][1].cout, 0, n, h, wd, s));
      cur = b[l];
      cur_c = up[j][1].cout;
    }
    // outc 1x1 64 -> 3*dim (+bias) (:72) followed by AdaptiveAvgPool2d(1) (:168): both are linear, so the score path averages
    // the 64-channel decoder output first and applies outc to the [n, 64] means in fp32 inside the head kernel - the
    // 48-channel full-resolution map (a write + a read of HWp * 96 B per image, plus its bf16 rounding in front of the
    // regressor) is only materialised when the caller asks for `enhanced_feat`
    KD_TRY(gap_mlp_tanh<T>(b[0], n, (int)HWp, 64, w.outc_f, w.outc_b, cc, w.w1, w.b1, w.w2, w.b2, w.w3, w.b3, score + b0, gap, s));
    if (feat) {
      ConvOp g;
      g.a0 = b[0]; g.c0 = 64; g.ld0 = 64; g.nimg = n; g.H = Hp; g.W = Wp; g.w = w.outc; g.w_ld = 64; g.w_tap_ld = 64;
      g.epi.col_bias = w.outc_b; g.epi.out = F48; g.epi.out_ld = cc; g.epi.N = cc; g.epi.H = Hp; g.epi.W = Wp;
      KD_TRY(conv_gemm<T>(g, s));
      KD_TRY(nhwc_to_planar<T>(F48, cc, feat + (long)b0 * cc * HWp, n, (int)HWp, cc, s));
    }
  }
  return 0;
}

#define INST(T)                                                                                                            \
  template size_t asdqe_packed_bytes<T>(const kdlae_asdqe_cfg&);                                                           \
  template int asdqe_pack<T>(const kdlae_asdqe_cfg&, const float* const*, int, void*, size_t, cudaStream_t);                \
  template size_t asdqe_workspace_bytes<T>(const kdlae_asdqe_cfg&, int, int, int);                                         \
  template int asdqe_forward<T>(const kdlae_asdqe_cfg&, const void*, const float*, const float*, float*, float*, int, int, \
                                int, int, void*, size_t, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace kd
