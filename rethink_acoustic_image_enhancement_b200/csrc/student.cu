// KDLAE-S (KDLAE_student, KDLAE/KDLAE_model.py:340-430): 3-D-conv U-Net over frame stacks.
// Frames live on the depth axis: rows of every implicit GEMM are (b, frame, y, x) pixels, NHWC storage;
// Conv3d 3x3x3 = 27-tap implicit GEMM with bias+ReLU epilogue, MaxPool3d(1,2,2) = per-frame 2x2 pool,
// ConvTranspose3d (1,2,2)/s(1,2,2) = 1x1 GEMM with 4 output phases scattered by the PixelShuffle-style
// epilogue plus the skip add at the destination.
#include <algorithm>
#include <type_traits>
#include "models.cuh"

namespace kd {

namespace {

// wx / bx: x-packed form (ConvOp::xpack_cin) of the narrow layers (16 or 32 input channels): P = 64/cin pixels per 128-byte
// super-pixel, so TMA boxes, smem rows, MMAs (N = 64) and stores are full width instead of 1/4 - 1/2 used.  A packed output of
// 128 columns (16 -> 32, 32 -> 64) runs as two launches over 64-column halves so the weight tiles stay resident in shared memory
template <typename T>
struct Conv27 { T* w; float* b; int cin, cout; bf16* wx; float* bx; };

template <typename T>
bool xpack_ok(int cin, int cout) {
  return std::is_same<T, bf16>::value && (cin == 16 || cin == 32) && ((64 / cin) * cout == 64 || (64 / cin) * cout == 128);
}

template <typename T>
struct StudentW {
  float* first_w; float* first_b;   // Conv3d 1 -> h0 as a direct conv, [27][1][h0]
  Conv27<T> e0b, e1a, e1b, fa, fb, d0a, d0b, d1a, d1b;
  T* up0; float* up0_b;             // [4*h1][h2], bias replicated per phase
  T* up1; float* up1_b;             // [4*h0][h1]
  float* out_w; float* out_b;       // 1x1x1 h0 -> 1
};

template <typename T>
void layout_student(const kdlae_student_cfg& c, Bump& b, StudentW<T>& w) {
  const int h0 = c.hidden[0], h1 = c.hidden[1], h2 = c.hidden[2];
  auto c27 = [&](Conv27<T>& k, int cin, int cout) {
    k.cin = cin; k.cout = cout;
    k.w = b.take<T>((size_t)cout * 27 * cin);
    k.b = b.take<float>(cout);
    k.wx = nullptr; k.bx = nullptr;
    if (xpack_ok<T>(cin, cout)) {
      k.wx = b.take<bf16>((size_t)(64 / cin) * cout * xconv_tiles(3, cin) * 64);
      k.bx = b.take<float>((64 / cin) * cout);
    }
  };
  w.first_w = b.take<float>((size_t)27 * h0);
  w.first_b = b.take<float>(h0);
  c27(w.e0b, h0, h0); c27(w.e1a, h0, h1); c27(w.e1b, h1, h1); c27(w.fa, h1, h2); c27(w.fb, h2, h2);
  w.up0 = b.take<T>((size_t)4 * h1 * h2); w.up0_b = b.take<float>(4 * h1);
  w.up1 = b.take<T>((size_t)4 * h0 * h1); w.up1_b = b.take<float>(4 * h0);
  c27(w.d0a, h1, h1); c27(w.d0b, h1, h1); c27(w.d1a, h0, h0); c27(w.d1b, h0, h0);
  w.out_w = b.take<float>(h0);
  w.out_b = b.take<float>(1);
}

template <typename T>
int pack27(const Conv27<T>& k, const float* wsrc, const float* bsrc, cudaStream_t s) {
  KD_CHECK(wsrc && bsrc, "student_pack: missing Conv3d tensor");
  PackOp p;
  p.src = wsrc; p.n_src = k.cout; p.c_src = k.cin; p.taps = 27; p.dst = k.w; p.n_dst = k.cout; p.c_dst = k.cin;
  KD_TRY(pack_weights<T>(p, s));
  if (k.wx) {
    KD_TRY(pack_xconv(wsrc, nullptr, k.cout, k.cin, 3, k.wx, s));
    for (int j = 0; j < 64 / k.cin; ++j) KD_TRY(copy_f32(bsrc, k.bx + j * k.cout, k.cout, s));
  }
  return copy_f32(bsrc, k.b, k.cout, s);
}

template <typename T>
int conv27(const Conv27<T>& k, const T* in, T* out, int nimg, int D, int H, int W, cudaStream_t s) {
  ConvOp g;
  if (k.wx && W % (64 / k.cin) == 0 && W / (64 / k.cin) >= 4) {      // x-packed: rows are super-pixels of P pixels
    const int P = 64 / k.cin, NP = P * k.cout;
    g.a0 = in; g.c0 = 64; g.ld0 = 64; g.nimg = nimg; g.D = D; g.H = H; g.W = W / P; g.kd = g.kh = g.kw = 3;
    g.xpack_cin = k.cin;
    g.w_tap_ld = 64; g.w_ld = (long)xconv_tiles(3, k.cin) * 64;
    g.epi.relu = 1; g.epi.out = out; g.epi.out_ld = NP; g.epi.N = 64; g.epi.H = H; g.epi.W = W / P;
    for (int half = 0; half < NP / 64; ++half) {
      g.w = k.wx + (long)half * 64 * g.w_ld;
      g.epi.col_bias = k.bx + half * 64;
      g.epi.out_coff = half * 64;
      KD_TRY(conv_gemm<T>(g, s));
    }
    return 0;
  }
  g.a0 = in; g.c0 = k.cin; g.ld0 = k.cin; g.nimg = nimg; g.D = D; g.H = H; g.W = W; g.kd = g.kh = g.kw = 3;
  g.w = k.w; g.w_ld = 27L * k.cin; g.w_tap_ld = k.cin;
  g.epi.col_bias = k.b; g.epi.relu = 1; g.epi.out = out; g.epi.out_ld = k.cout; g.epi.N = k.cout; g.epi.H = H; g.epi.W = W;
  return conv_gemm<T>(g, s);
}

// ConvTranspose3d kernel (1,2,2) stride (1,2,2) (:378, no overlap) + skip add (:417).
// w: [4*cout][cin] with rows phase-major (s = 2*dy + dx), bias4 replicated per phase.  For a fixed output-row phase dy the two dx
// phases of input pixel (y, x) are the ADJACENT output pixels (2y+dy, 2x), (2y+dy, 2x+1): 2*cout contiguous channels.  So the
// bf16 path runs two 1x1 GEMMs (N = 2*cout) whose destination is the output seen as [nimg, H, W, 2*cout] with a row stride of
// two output rows - plain TMA slab stores and a TMA-loaded skip tile instead of the scattered PixelShuffle epilogue.
template <typename T>
int upconv(const T* in, int cin, const T* w, const float* bias4, int cout, const T* skip, T* out, int nimg, int H, int W,
           cudaStream_t s) {
  if (std::is_same<T, bf16>::value && (2 * cout) % 8 == 0 && cin % 8 == 0) {
    for (int dy = 0; dy < 2; ++dy) {
      ConvOp g;
      g.a0 = in; g.c0 = cin; g.ld0 = cin; g.nimg = nimg; g.H = H; g.W = W;
      g.w = w + (long)dy * 2 * cout * cin; g.w_ld = cin; g.w_tap_ld = cin;
      const long row = 2L * W * cout;                       // one output row (2W pixels of cout channels)
      g.epi.col_bias = bias4 + dy * 2 * cout;
      g.epi.res = skip + dy * row; g.epi.res_ld = 2 * cout; g.epi.res_y_ld = 2 * row;
      g.epi.out = out + dy * row; g.epi.out_ld = 2 * cout; g.epi.out_y_ld = 2 * row;
      g.epi.N = 2 * cout; g.epi.H = H; g.epi.W = W;
      KD_TRY(conv_gemm<T>(g, s));
    }
    return 0;
  }
  ConvOp g;
  g.a0 = in; g.c0 = cin; g.ld0 = cin; g.nimg = nimg; g.H = H; g.W = W;
  g.w = w; g.w_ld = cin; g.w_tap_ld = cin;
  g.epi.col_bias = bias4; g.epi.res = skip; g.epi.res_ld = cout; g.epi.out = out; g.epi.out_ld = cout; g.epi.N = 4 * cout;
  g.epi.mode = OUT_PIXEL_SHUFFLE; g.epi.cq = cout; g.epi.H = H; g.epi.W = W;
  return conv_gemm<T>(g, s);
}

struct SWs { size_t s0, t0, t1, s1, h0, h1, q0, q1, total; };
template <typename T>
SWs sws_layout(const kdlae_student_cfg& c, int mb, int F, int H, int W) {
  Bump b;
  SWs L;
  const size_t P = (size_t)mb * F * H * W;
  auto off = [&](size_t elems) { b.off = align_up(b.off, 256); size_t o = b.off; b.off += elems * sizeof(T); return o; };
  const size_t c0 = c.hidden[0], c1 = std::max(c.hidden[0], c.hidden[1]), c2 = std::max(c.hidden[1], c.hidden[2]);
  L.s0 = off(P * c0); L.t0 = off(P * c0); L.t1 = off(P * c0);
  L.s1 = off(P / 4 * c1); L.h0 = off(P / 4 * c1); L.h1 = off(P / 4 * c1);
  L.q0 = off(P / 16 * c2); L.q1 = off(P / 16 * c2);
  L.total = align_up(b.off, 256);
  return L;
}

}  // namespace

template <typename T>
size_t student_packed_bytes(const kdlae_student_cfg& cfg) {
  Bump b;
  StudentW<T> w;
  layout_student<T>(cfg, b, w);
  return align_up(b.off, 256);
}

template <typename T>
int student_pack(const kdlae_student_cfg& c, const float* const* t, int n_tensors, void* packed, size_t packed_bytes,
                 cudaStream_t s) {
  KD_CHECK(n_tensors == 26, "student_pack: expected 26 state_dict tensors, got %d", n_tensors);
  for (int i = 0; i < 3; ++i) KD_CHECK(c.hidden[i] % 8 == 0 && c.hidden[i] > 0, "student_pack: hidden_channels must be multiples of 8");
  for (int i = 0; i < 26; ++i) KD_CHECK(t[i] != nullptr, "student_pack: tensor %d is NULL", i);
  Bump b;
  b.base = reinterpret_cast<uint8_t*>(packed);
  StudentW<T> w;
  layout_student<T>(c, b, w);
  KD_CHECK(b.off <= packed_bytes, "student_pack: packed buffer too small");
  const int h0 = c.hidden[0], h1 = c.hidden[1], h2 = c.hidden[2];
  // state_dict order: encoders.{0,1}.{0,2}, st_fusion.{0,2}, upconv_layers.{0,1}, decoders.{0,1}.{0,2}, out_conv
  KD_TRY(pack_few_in(t[0], h0, 1, 27, nullptr, w.first_w, s));
  KD_TRY(copy_f32(t[1], w.first_b, h0, s));
  KD_TRY(pack27<T>(w.e0b, t[2], t[3], s));
  KD_TRY(pack27<T>(w.e1a, t[4], t[5], s));
  KD_TRY(pack27<T>(w.e1b, t[6], t[7], s));
  KD_TRY(pack27<T>(w.fa, t[8], t[9], s));
  KD_TRY(pack27<T>(w.fb, t[10], t[11], s));
  PackOp p;
  p.src = t[12]; p.n_src = h1; p.c_src = h2; p.taps = 1; p.mode = PACK_CONVT; p.dst = w.up0; p.n_dst = 4 * h1; p.c_dst = h2;
  KD_TRY(pack_weights<T>(p, s));
  for (int ph = 0; ph < 4; ++ph) KD_TRY(copy_f32(t[13], w.up0_b + ph * h1, h1, s));
  p = PackOp();
  p.src = t[14]; p.n_src = h0; p.c_src = h1; p.taps = 1; p.mode = PACK_CONVT; p.dst = w.up1; p.n_dst = 4 * h0; p.c_dst = h1;
  KD_TRY(pack_weights<T>(p, s));
  for (int ph = 0; ph < 4; ++ph) KD_TRY(copy_f32(t[15], w.up1_b + ph * h0, h0, s));
  KD_TRY(pack27<T>(w.d0a, t[16], t[17], s));
  KD_TRY(pack27<T>(w.d0b, t[18], t[19], s));
  KD_TRY(pack27<T>(w.d1a, t[20], t[21], s));
  KD_TRY(pack27<T>(w.d1b, t[22], t[23], s));
  KD_TRY(pack_few_out(t[24], 1, h0, 1, w.out_w, s));
  KD_TRY(copy_f32(t[25], w.out_b, 1, s));
  return 0;
}

template <typename T>
size_t student_workspace_bytes(const kdlae_student_cfg& cfg, int mb, int F, int H, int W) {
  return sws_layout<T>(cfg, mb, F, H, W).total;
}

template <typename T>
int student_forward(const kdlae_student_cfg& c, const void* packed, const float* x, float* y, int B, int F, int H, int W,
                    int micro_batch, void* ws, size_t ws_bytes, cudaStream_t s) {
  KD_CHECK(H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, "KDLAE_student: H and W must be multiples of 4 (got %dx%d)", H, W);
  KD_CHECK(B >= 1 && F >= 1 && micro_batch >= 1, "KDLAE_student: bad batch/frames");
  if (micro_batch > B) micro_batch = B;
  const SWs L = sws_layout<T>(c, micro_batch, F, H, W);
  KD_CHECK(ws_bytes >= L.total, "KDLAE_student: workspace too small (%zu < %zu)", ws_bytes, L.total);
  Bump b;
  b.base = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(packed));
  StudentW<T> w;
  layout_student<T>(c, b, w);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
  T* S0 = reinterpret_cast<T*>(wsb + L.s0); T* T0 = reinterpret_cast<T*>(wsb + L.t0); T* T1 = reinterpret_cast<T*>(wsb + L.t1);
  T* S1 = reinterpret_cast<T*>(wsb + L.s1); T* H0 = reinterpret_cast<T*>(wsb + L.h0); T* H1 = reinterpret_cast<T*>(wsb + L.h1);
  T* Q0 = reinterpret_cast<T*>(wsb + L.q0); T* Q1 = reinterpret_cast<T*>(wsb + L.q1);
  const int h0 = c.hidden[0], h1 = c.hidden[1], h2 = c.hidden[2];
  const long HW = (long)H * W;

  for (int b0 = 0; b0 < B; b0 += micro_batch) {
    const int n = std::min(micro_batch, B - b0);
    const int nimg = n * F;
    const float* xb = x + (long)b0 * F * HW;
    float* yb = y + (long)b0 * F * HW;
    // encoders[0]: Conv3d 1->h0 (direct) + ReLU, Conv3d h0->h0 + ReLU  (:389-392)
    SmallConv fi;
    fi.in0 = xb; fi.in0_img = HW; fi.in0_ch = 0; fi.cin0 = 1; fi.nimg = nimg; fi.D = F; fi.H = H; fi.W = W; fi.kd = 3;
    fi.w = w.first_w; fi.bias = w.first_b; fi.cout = h0; fi.relu = 1; fi.out = T0; fi.out_ld = h0;
    KD_TRY(conv_few_in<T>(fi, s));
    KD_TRY(conv27<T>(w.e0b, T0, S0, nimg, F, H, W, s));
    KD_TRY(maxpool2x2<T>(S0, H0, nimg, H, W, h0, s));                       // MaxPool3d (1,2,2) (:366)
    KD_TRY(conv27<T>(w.e1a, H0, H1, nimg, F, H / 2, W / 2, s));
    KD_TRY(conv27<T>(w.e1b, H1, S1, nimg, F, H / 2, W / 2, s));
    KD_TRY(maxpool2x2<T>(S1, Q0, nimg, H / 2, W / 2, h1, s));
    KD_TRY(conv27<T>(w.fa, Q0, Q1, nimg, F, H / 4, W / 4, s));              // st_fusion (:373)
    KD_TRY(conv27<T>(w.fb, Q1, Q0, nimg, F, H / 4, W / 4, s));
    KD_TRY(upconv<T>(Q0, h2, w.up0, w.up0_b, h1, S1, H0, nimg, H / 4, W / 4, s));   // + encoder skip (:417)
    KD_TRY(conv27<T>(w.d0a, H0, H1, nimg, F, H / 2, W / 2, s));
    KD_TRY(conv27<T>(w.d0b, H1, H0, nimg, F, H / 2, W / 2, s));
    KD_TRY(upconv<T>(H0, h1, w.up1, w.up1_b, h0, S0, T0, nimg, H / 2, W / 2, s));
    KD_TRY(conv27<T>(w.d1a, T0, T1, nimg, F, H, W, s));
    KD_TRY(conv27<T>(w.d1b, T1, T0, nimg, F, H, W, s));
    // out_conv 1x1x1 h0 -> 1 (+ x when residual) (:384,:425-426)
    SmallConvOut fo;
    fo.in = T0; fo.in_ld = h0; fo.cin = h0; fo.nimg = nimg; fo.H = H; fo.W = W; fo.k = 1; fo.w = w.out_w; fo.bias = w.out_b;
    fo.cout = 1; fo.out = yb; fo.out_img = HW; fo.out_ch = HW;
    if (c.residual) { fo.res = xb; fo.res_img = HW; fo.res_ch = HW; }
    KD_TRY(conv_few_out<T>(fo, s));
  }
  return 0;
}

#define INST(T)                                                                                                         \
  template size_t student_packed_bytes<T>(const kdlae_student_cfg&);                                                    \
  template int student_pack<T>(const kdlae_student_cfg&, const float* const*, int, void*, size_t, cudaStream_t);         \
  template size_t student_workspace_bytes<T>(const kdlae_student_cfg&, int, int, int, int);                             \
  template int student_forward<T>(const kdlae_student_cfg&, const void*, const float*, float*, int, int, int, int, int, \
                                  void*, size_t, cudaStream_t);
INST(float)
INST(bf16)
#undef INST

}  // namespace kd
