// Depthwise 3x3 (+ optional GELU gate) for NHWC bf16 on the CUDA cores with packed fp32 FMAs (FFMA2, fma.rn.f32x2)
//   (KDLAE_model.py:97,103-104 FeedForward.dwconv + gate, :119 Attention.qkv_dwconv).
//
// Why not the tensor cores (round-1 dwconv_tc.cu, removed; numbers in profiles/r01_summary.md): a depthwise conv has no reuse across channels, so the diagonal-weight MMA reads
// its 128 x 16 activation operand from shared memory once per tap - 9x the tile - and ncu shows that kernel pinned on
// L1/shared-memory throughput (82 %) at ~55 % of the HBM rate.  Here every input value is read from shared memory ~1.9x:
//   * lane = one bf16x2 channel pair of a 64-channel block (a warp-wide LDS.32 is exactly one pixel's 128-byte row, so
//     there are no bank conflicts and global stores are full 128-byte lines), warp = PXT adjacent output columns;
//   * the warp walks down the tile rows: each input row is loaded and unpacked once ((PXT+2) LDS.32 + as many shift/mask
//     pairs) and feeds the three output rows it belongs to, which live in 3 x PXT rotating packed accumulators;
//   * the 9 (18 with the gate) weight pairs stay in registers for the whole persistent CTA (a CTA keeps one channel block);
//   * one FFMA2 does two channel FMAs: 4.5 issue slots per output instead of 9, which is what makes the CUDA-core
//     version fit the issue budget at HBM speed (B200 measured: FFMA and FFMA2 both sustain 128 FMA/clk/SM).
// Staging: one 4-D TMA box {64 ch, TW+2, TH+2, 1} per tile (two with the gate: the matching channels of both chunk(2)
// halves); zero padding = TMA out-of-bounds fill; two stages per CTA, two CTAs per SM.
// Weights are the fp32 originals (the tensor-core variant had to round them to bf16).
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// bf16x2 word -> packed {fp32 even channel, fp32 odd channel}
__device__ __forceinline__ u64 unpack2(uint32_t w) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(w << 16), "r"(w & 0xffff0000u));
  return r;
}
__device__ __forceinline__ float2 as_float2(u64 v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ u64 pack2f(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 splat2(float c) { return pack2f(c, c); }

// gelu(a) * b on a packed channel pair (the same function is used by pwdw_f2.cu / pwdw_t.cu, which keeps the fused and the
// unfused schedules bit-identical)
__device__ __forceinline__ u64 gelu_gate2(u64 a, u64 b) {
  // gelu(x) = x * Phi(x) with Phi(x) = 1 / (1 + 2^(x * Q(x^2))):  Q = -2 log2(e) * P and P(x^2) ~ atanh(erf(x / sqrt2)) / x is a
  // degree-4 fit (P > 0 everywhere, so the sigmoid saturates correctly for any |x|).  |gelu error| <= 5e-6 in fp32 (the
  // result is rounded to bf16: half-ulp 2e-3 relative); x^2, the Horner chain and the products run as packed FFMA2 / FMUL2,
  // ex2.approx + rcp.approx are the two MUFU ops: 6.5 issue slots per value (the A&S 7.1.26 erfc form took 9.5).
  const u64 t = fmul2(a, a);
  u64 q = ffma2(t, splat2(-3.583463763e-06f), splat2(9.391572204e-05f));
  q = ffma2(q, t, splat2(3.380707789e-04f));
  q = ffma2(q, t, splat2(-1.052193928e-01f));
  q = ffma2(q, t, splat2(-2.302019163e+00f));
  const float2 z = as_float2(fmul2(a, q));
  float e0, e1, r0, r1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(z.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(z.y));
  const float2 d = as_float2(ffma2(pack2f(e0, e1), splat2(1.0f), splat2(1.0f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
  return fmul2(fmul2(a, pack2f(r0, r1)), b);
}

constexpr int F2_TH = 8, F2_IN_H = F2_TH + 2, F2_WARPS = 8, F2_THREADS = F2_WARPS * 32;
template <int GATE> struct F2Cfg {
  static constexpr int PXT = GATE ? 2 : 4;              // output columns per warp
  static constexpr int TW = F2_WARPS * PXT, IN_W = TW + 2;
  static constexpr uint32_t TILE_BYTES = IN_W * F2_IN_H * 128;
  static constexpr uint32_t STAGE = (GATE ? 2 : 1) * TILE_BYTES;
  static constexpr uint32_t SMEM = 2 * STAGE + 128 + 64;
};

struct F2Params {
  bf16* out; long ldo;
  const float* w9c;
  int H, W, C, Cout, hp;
  int tiles_x, tiles_y, cblocks, G;
  int nsp;                  // spatial tiles (images x tiles_y x tiles_x), < 2^24
  float inv_tiles_x, inv_tiles_img;
};

template <int GATE>
__global__ void __launch_bounds__(F2_THREADS, 2)
k_dwconv_f2(const __grid_constant__ CUtensorMap map, const F2Params p) {
  constexpr int PXT = F2Cfg<GATE>::PXT, TW = F2Cfg<GATE>::TW, IN_W = F2Cfg<GATE>::IN_W;
  constexpr uint32_t TILE_BYTES = F2Cfg<GATE>::TILE_BYTES, STAGE = F2Cfg<GATE>::STAGE;
  constexpr int NH = GATE ? 2 : 1;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t bar0 = sbase + 2 * STAGE;

  const int cb = blockIdx.x % p.cblocks, g = blockIdx.x / p.cblocks;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_img = p.tiles_x * p.tiles_y;

  auto issue = [&](int sp, int stage) {   // one elected thread: arm the barrier and launch the TMA box(es)
    const int img = fast_div(sp, tiles_img, p.inv_tiles_img);
    const int r = sp - img * tiles_img;
    const int tyi = fast_div(r, p.tiles_x, p.inv_tiles_x), txi = r - tyi * p.tiles_x;
    const uint32_t bar = bar0 + 8u * stage, dst = sbase + stage * STAGE;
    mbar_expect_tx(bar, STAGE);
    tma_load_4d(dst, &map, bar, cb * 64, txi * TW - 1, tyi * F2_TH - 1, img);
    if (GATE) tma_load_4d(dst + TILE_BYTES, &map, bar, p.hp + cb * 64, txi * TW - 1, tyi * F2_TH - 1, img);
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&map);
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (g < p.nsp) issue(g, 0);
  }

  // this lane's channel pair and its 9 (x2) packed weights
  const int ch = cb * 64 + lane * 2;
  const bool ch_ok = ch < p.Cout;
  u64 w[NH][9];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float2 f = make_float2(0.f, 0.f);
      if (ch_ok) f = __ldg(reinterpret_cast<const float2*>(p.w9c + (long)t * p.C + h * p.hp + ch));
      asm("mov.b64 %0, {%1, %2};" : "=l"(w[h][t]) : "f"(f.x), "f"(f.y));
    }
  __syncthreads();

  const int xs = warp * PXT;
  const long row_pitch2 = (long)p.W * p.ldo * 2;
  int it = 0;
  const int ldo2 = (int)p.ldo * 2;
  for (int sp = g; sp < p.nsp; sp += p.G, ++it) {
    const int stage = it & 1;
    if (threadIdx.x == 0 && sp + p.G < p.nsp) issue(sp + p.G, stage ^ 1);
    const int img = fast_div(sp, tiles_img, p.inv_tiles_img);
    const int r = sp - img * tiles_img;
    const int tyi = fast_div(r, p.tiles_x, p.inv_tiles_x), txi = r - tyi * p.tiles_x;
    const int x0 = txi * TW + xs, y0 = tyi * F2_TH;
    const int nq = ch_ok ? p.W - x0 : 0, nr = p.H - y0;     // valid output columns / rows of this thread's strip
    uint8_t* orp = reinterpret_cast<uint8_t*>(p.out + (((long)img * p.H + y0) * p.W + x0) * p.ldo + ch);
    const uint32_t sa = sbase + stage * STAGE + xs * 128 + lane * 4;

    mbar_wait(bar0 + 8u * stage, (it >> 1) & 1);

    u64 acc[NH][3][PXT];
#pragma unroll
    for (int ir = 0; ir < F2_IN_H; ++ir) {
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        u64 v[PXT + 2];
#pragma unroll
        for (int j = 0; j < PXT + 2; ++j) v[j] = unpack2(lds32(sa + h * TILE_BYTES + (ir * IN_W + j) * 128));
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {        // dx outermost: consecutive FFMA2s go to different accumulators
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int orow = ir - dy;
            if (orow >= 0 && orow < F2_TH) {
              const int a = orow % 3;
#pragma unroll
              for (int q = 0; q < PXT; ++q) {
                if (dy == 0 && dx == 0) acc[h][a][q] = fmul2(v[q], w[h][0]);
                else acc[h][a][q] = ffma2(v[q + dx], w[h][dy * 3 + dx], acc[h][a][q]);
              }
            }
          }
        }
      }
      if (ir >= 2) {
        const int orow = ir - 2, a = orow % 3;
#pragma unroll
        for (int q = 0; q < PXT; ++q) {
          const float2 f = as_float2(GATE ? gelu_gate2(acc[0][a][q], acc[NH - 1][a][q]) : acc[0][a][q]);
          if (orow < nr && q < nq) *reinterpret_cast<uint32_t*>(orp + q * ldo2) = pack_bf16x2(f.x, f.y);
        }
        orp += row_pitch2;
      }
    }
    __syncthreads();   // everyone is done reading this stage before it is refilled two iterations later
  }
}

template <int GATE>
int launch_f2(const bf16* x, long ldx, bf16* out, long ldo, const float* w9c, int nimg, int H, int W, int C, cudaStream_t s) {
  typedef F2Cfg<GATE> Cfg;
  static DeviceOnce once;
  bool first; int dev, g_f2_sms;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_dwconv_f2<GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    device_mark(once, dev);
  }
  KD_TRY(device_sms(&g_f2_sms));
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
  const cuuint64_t str[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H};
  const cuuint32_t box[4] = {64, (cuuint32_t)Cfg::IN_W, F2_IN_H, 1};
  KD_TRY(make_map(&map, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));
  F2Params p;
  p.out = out; p.ldo = ldo; p.w9c = w9c;
  p.H = H; p.W = W; p.C = C; p.Cout = GATE ? C / 2 : C; p.hp = GATE ? C / 2 : 0;
  p.tiles_x = cdiv(W, Cfg::TW); p.tiles_y = cdiv(H, F2_TH); p.cblocks = cdiv(p.Cout, 64);
  const long nsp = (long)nimg * p.tiles_x * p.tiles_y;
  KD_CHECK(nsp < (1L << 24) && ldo < (1L << 28), "dwconv3x3: too many tiles (%ld)", nsp);
  p.nsp = (int)nsp;
  p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_img = 1.0f / (float)(p.tiles_x * p.tiles_y);
  p.G = (int)std::min<long>(nsp, std::max(1, 2 * g_f2_sms / p.cblocks));
  ProfScope prof(PC_DWCONV, s, 18.0 * nimg * H * W * C, (double)nimg * H * W * (C + p.Cout) * 2.0 + 36.0 * C);
  k_dwconv_f2<GATE><<<p.G * p.cblocks, F2_THREADS, Cfg::SMEM, s>>>(map, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// bf16 entry used by dwconv3x3<bf16>; returns -1 when the shape is not eligible
int dwconv3x3_f2(const bf16* x, long ldx, bf16* out, long ldo, const float* w9c, int nimg, int H, int W, int C, int gate,
                 cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 3) || ldx % 8 || ldo % 2 || C % (gate ? 16 : 8)) return -1;
  return gate ? launch_f2<1>(x, ldx, out, ldo, w9c, nimg, H, W, C, s) : launch_f2<0>(x, ldx, out, ldo, w9c, nimg, H, W, C, s);
}

}  // namespace kd
