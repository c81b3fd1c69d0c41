// Fused  LayerNorm-folded 1x1 conv (tcgen05)  ->  depthwise 3x3 (packed FFMA2 on the CUDA cores)  (-> GELU gate), bf16 path.
//   qkv branch  (KDLAE_model.py:118-119):  qkv' = dw3x3(W_qkv . LN(x))
//   GDFN branch (KDLAE_model.py:95-104):   g    = gelu(dw(t)[:h]) * dw(t)[h:],  t = W_in . LN(x)
// The 3C / 2h wide intermediate t is the largest tensor of a TransformerBlock and the unfused schedule pays for it twice
// (a write-bound GEMM, then an issue-bound depthwise kernel reading it back).  Here it only exists as an 8 x 32 pixel
// shared-memory tile.  Per CTA, per 6 x 30 pixel output tile, for each 64-channel block of the output (the x tile is loaded
// once per tile and stays resident while the W1 row blocks stream through a single buffer; with the x tile reloaded per
// channel block, its TMA latency sat on the per-item critical path and the depthwise warps idled 22 % of the time):
//   1. TMA: x tile with halo {C, 32 px, 8 rows} (zero fill outside the image = the conv's zero padding); W1 rows of the block
//   2. tcgen05 (M = 128 x 2, N = 64 | 128, K = C):  T = X . W1^T  -> TMEM     (halo recomputed: 1.42x of a tiny GEMM)
//   3. conversion warps: tcgen05.ld, * rstd[pixel] (BiasFree LayerNorm folded: gamma in W1, rstd here), bf16, st.shared
//      into a 128B-swizzled t tile with a row pitch of 32 pixels (same rounding point as the unfused schedule, so the
//      two schedules are bit-identical)
//   4. depthwise warps: the sliding-window FFMA2 loop of dwconv_f2.cu over the t tile (lane = channel pair, warp = 3
//      output columns, 9 weight pairs per half in registers), GELU gate, direct 128-byte-per-pixel global stores.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..5 conversion (one per TMEM lane quarter), 6.. depthwise (10 x 3 columns, or
// with the gate 8 x 4 columns in two passes).
// t tiles are double buffered, so the conversion of item n+1 and the GEMM of item n+2 run under the depthwise of item n
// (item = (tile, channel block)).
// Role in the default schedule: every fused pair of a WithBias-LayerNorm model (this kernel carries the mean / bias fold, template
// WB) and the bit-identical-to-unfused reference of the tests; BiasFree models take pwdw_t.cu (KDLAE_FUSE_PWDW, teacher.cu).
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

typedef unsigned long long u64;

constexpr int PF_TW = 32, PF_OW = 30, PF_OH = 6, PF_IH = 8, PF_CB = 64;
constexpr uint32_t PF_XCHUNK = PF_TW * PF_IH * 128;          // 32768: one 64-channel K chunk of the x tile / one t tile
constexpr int PF_E1_WARPS = 4;
// Depthwise warps: plain variant 10 warps x 3 columns (= the 30 output columns, single pass); gate variant 8 warps x 4 columns
// done as two passes of 2 columns (register budget: two accumulator sets + 18 weight pairs; measured 5 % faster than 10 x 3).
template <int GATE> struct PfCfg {
  static constexpr int DW_WARPS = GATE ? 8 : 10;
  static constexpr int PXT = GATE ? 2 : 3;         // output columns per pass
  static constexpr int NPASS = GATE ? 2 : 1;
  static constexpr int THREADS = (2 + PF_E1_WARPS + DW_WARPS) * 32;
};

struct PfParams {
  int H, W, C, Nt, Cout, nimg;     // Nt = rows of W1 (3C or 2hp); Cout = output channels (3C or hp)
  int kc;                          // 64-wide K chunks of C (1 or 2)
  int tiles_x, tiles_y, cblocks;
  long ntiles_all;                 // spatial tiles (images x tiles_y x tiles_x)
  float inv_tiles_x, inv_tiles_y;
  const float* rstd;               // [nimg*H*W]
  const float* mu;                 // WithBias LayerNorm (:67-70): per-pixel mean, and the two per-column fold vectors
  const float* s1; const float* s2;   //   t = rstd * acc - rstd * mu * s1[n] + s2[n]   (pack_ln_cols)
  const float* w9c;                // depthwise weights fp32 [9][Nt]
  bf16* out; long ldo;
};

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
__device__ __forceinline__ u64 pack2f(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 splat2(float c) { return pack2f(c, c); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// gelu(a) * b on a packed channel pair (identical to dwconv_f2.cu)
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

template <int GATE, int WB>
__global__ void __launch_bounds__(PfCfg<GATE>::THREADS, 1)
k_pwdw_f2(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1, const PfParams p) {
  constexpr int NH = GATE ? 2 : 1;
  constexpr int N1 = 64 * NH;                       // GEMM N: t channels of this block (both halves for the gate)
  constexpr int PXT = PfCfg<GATE>::PXT, NPASS = PfCfg<GATE>::NPASS, PF_DW_WARPS = PfCfg<GATE>::DW_WARPS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // smem (all 1024-aligned): W1 [kc][NH][64 rows][128 B] | x chunks | t[2][NH] tiles | barriers
  const uint32_t w1_base = sbase;
  const uint32_t x_base = w1_base + p.kc * NH * 8192;
  const uint32_t t_base = x_base + p.kc * PF_XCHUNK;
  const uint32_t bar_base = t_base + 2 * NH * PF_XCHUNK;
  const uint32_t w_full = bar_base, x_full = bar_base + 8, x_empty = bar_base + 16, d1_full = bar_base + 24, d1_empty = bar_base + 32;
  const uint32_t w_empty = bar_base + 80;
  auto t_ready = [&](int b) { return bar_base + 40 + 8u * b; };
  auto t_free = [&](int b) { return bar_base + 56 + 8u * b; };
  const uint32_t tmem_slot = bar_base + 72;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* t_gen = smem_raw + (t_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = p.Nt / 2;
  const int ksteps = (p.C + 15) / 16;               // K = 16 MMA steps over the C input channels
  const int ntiles = ((long)blockIdx.x < p.ntiles_all) ? (int)((p.ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int ncb = p.cblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_w1);
    mbar_init(w_full, 1); mbar_init(w_empty, 1); mbar_init(x_full, 1); mbar_init(x_empty, 1); mbar_init(d1_full, 1); mbar_init(d1_empty, PF_E1_WARPS);
    for (int b = 0; b < 2; ++b) { mbar_init(t_ready(b), 1); mbar_init(t_free(b), PF_DW_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * N1));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_xy = [&](int i, int& img, int& y0, int& x0) {          // i-th tile of this CTA
    const int t = blockIdx.x + i * gridDim.x;
    const int rowt = fast_div(t, p.tiles_x, p.inv_tiles_x);
    const int txi = t - rowt * p.tiles_x;
    img = fast_div(rowt, p.tiles_y, p.inv_tiles_y);
    const int tyi = rowt - img * p.tiles_y;
    x0 = txi * PF_OW; y0 = tyi * PF_OH;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t n = 0;
      for (int i = 0; i < ntiles; ++i) {
        int img, y0, x0;
        if (i + 1 < ntiles) {   // the x tile is single buffered: keep the next one warm in L2
          tile_xy(i + 1, img, y0, x0);
          for (int k = 0; k < p.kc; ++k) tma_prefetch_4d(&map_x, k * 64, x0 - 1, y0 - 1, img);
        }
        mbar_wait_lazy(x_empty, (i & 1) ^ 1);        // last GEMM of tile i-1 retired
        tile_xy(i, img, y0, x0);
        mbar_expect_tx(x_full, p.kc * PF_XCHUNK);
        for (int k = 0; k < p.kc; ++k) tma_load_4d(x_base + k * PF_XCHUNK, &map_x, x_full, k * 64, x0 - 1, y0 - 1, img);
        for (int cb = 0; cb < ncb; ++cb, ++n) {
          mbar_wait_lazy(w_empty, (n & 1) ^ 1);      // GEMM n-1 retired: the W1 buffer is free
          mbar_expect_tx(w_full, p.kc * NH * 8192);
          for (int k = 0; k < p.kc; ++k)
            for (int h = 0; h < NH; ++h)
              tma_load_3d(w1_base + (k * NH + h) * 8192, &map_w1, w_full, k * 64, (GATE ? h * hp : 0) + cb * PF_CB, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: T = X . W1^T for both M tiles (pixels 0..127, 128..255) =====================
    if (lane == 0) {
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lo_tag = 1u << 16;
      uint32_t n = 0;
      for (int i = 0; i < ntiles; ++i) {
        for (int cb = 0; cb < ncb; ++cb, ++n) {
          mbar_wait_lazy(d1_empty, (n & 1) ^ 1);   // conversion warps have drained D1 of item n-1
          mbar_wait_relaxed(w_full, n & 1);
          if (cb == 0) mbar_wait_relaxed(x_full, i & 1);
          tc_fence_after();
          for (int mt = 0; mt < 2; ++mt) {
            for (int ks = 0; ks < ksteps; ++ks) {
              const int k = ks >> 2, kk = ks & 3;
              const uint32_t a_lo = (((x_base + k * PF_XCHUNK + mt * 16384 + kk * 32) & 0x3FFFF) >> 4) | lo_tag;
              const uint32_t b_lo = (((w1_base + k * NH * 8192 + kk * 32) & 0x3FFFF) >> 4) | lo_tag;
              umma_bf16_lohi(tmem_base + mt * N1, a_lo, b_lo, desc_hi, idesc, ks != 0 ? 1u : 0u);
            }
          }
          umma_commit(w_empty);                       // W1 block consumed
          if (cb == ncb - 1) umma_commit(x_empty);    // x tile consumed
          umma_commit(d1_full);                       // T accumulators ready
        }
      }
    }
  } else if (warp < 2 + PF_E1_WARPS) {
    // ===================== conversion warps: T (fp32, TMEM) * rstd -> bf16 t tile =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;        // TMEM lane = row of the M tile
    uint32_t n = 0;
    for (int i = 0; i < ntiles; ++i) {
      int img, y0, x0;
      tile_xy(i, img, y0, x0);
      float rs[2], rmu[2], inside[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int pix = mt * 128 + r;           // pixel of the 8 x 32 halo tile
        const int y = y0 - 1 + pix / PF_TW, x = x0 - 1 + pix % PF_TW;
        const bool inimg = y >= 0 && y < p.H && x >= 0 && x < p.W;
        rs[mt] = inimg ? __ldg(p.rstd + ((long)img * p.H + y) * p.W + x) : 0.f;   // 0 outside: conv zero padding of t
        rmu[mt] = (WB && inimg) ? __ldg(p.mu + ((long)img * p.H + y) * p.W + x) * rs[mt] : 0.f;
        inside[mt] = inimg ? 1.f : 0.f;
      }
      for (int cb = 0; cb < ncb; ++cb, ++n) {
      const int b = n & 1;
      mbar_wait_lazy(t_free(b), ((n >> 1) & 1) ^ 1);   // t[b] no longer read by the depthwise warps (item n-2)
      mbar_wait_relaxed(d1_full, n & 1);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int pix = mt * 128 + r;
        const uint32_t t_row = tmem_base + mt * N1 + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // two runs of 32 channels per chunk(2) half
          uint32_t v[NH * 2][16];
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            tmem_ld16_issue(t_row + h * 64 + half * 32, v[h * 2]);
            tmem_ld16_issue(t_row + h * 64 + half * 32 + 16, v[h * 2 + 1]);
          }
#pragma unroll
          for (int u = 0; u < NH * 2; ++u) tmem_ld16_wait(v[u]);
          if (mt == 1 && half == 1) {             // all TMEM reads of this tile done: the next GEMM may overwrite D1
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_empty);
          }
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            uint8_t* trow = t_gen + (b * NH + h) * PF_XCHUNK + pix * 128;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              uint32_t (&vv)[16] = v[h * 2 + (c4 >> 1)];
              const int o = (c4 & 1) * 8;
              uint4 w4;       // packed multiply by rstd (FMUL2), then one cvt.rn.bf16x2 per pair
              const u64 rs2 = splat2(rs[mt]);
              float2 f0 = as_float2(fmul2(pack2f(__uint_as_float(vv[o + 0]), __uint_as_float(vv[o + 1])), rs2));
              float2 f1 = as_float2(fmul2(pack2f(__uint_as_float(vv[o + 2]), __uint_as_float(vv[o + 3])), rs2));
              float2 f2 = as_float2(fmul2(pack2f(__uint_as_float(vv[o + 4]), __uint_as_float(vv[o + 5])), rs2));
              float2 f3 = as_float2(fmul2(pack2f(__uint_as_float(vv[o + 6]), __uint_as_float(vv[o + 7])), rs2));
              if (WB) {       // WithBias LayerNorm fold, same operation order as the GEMM epilogue: (v*rs - rmu*s1[n]) + s2[n]
                const int nb = (GATE ? h * hp : 0) + cb * PF_CB + half * 32 + c4 * 8;
                float* ff[4] = {&f0.x, &f1.x, &f2.x, &f3.x};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float& t = ff[e >> 1][e & 1];
                  const bool nok = nb + e < p.Nt;
                  const float a1 = nok ? __ldg(p.s1 + nb + e) : 0.f, a2 = nok ? __ldg(p.s2 + nb + e) : 0.f;
                  t = __fadd_rn(fmaf(-rmu[mt], a1, t), a2 * inside[mt]);
                }
              }
              w4.x = pack_bf16x2(f0.x, f0.y); w4.y = pack_bf16x2(f1.x, f1.y);
              w4.z = pack_bf16x2(f2.x, f2.y); w4.w = pack_bf16x2(f3.x, f3.y);
              const int chunk = half * 4 + c4;
              *reinterpret_cast<uint4*>(trow + ((chunk ^ (pix & 7)) << 4)) = w4;
            }
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(PF_E1_WARPS * 32) : "memory");          // t tile complete
      if (ew == 0 && lane == 0) mbar_arrive(t_ready(b));
      }
    }
  } else {
    // ===================== depthwise warps: sliding-window FFMA2 over the t tile -> (gate) -> global =====================
    const int fw = warp - 2 - PF_E1_WARPS;      // owns output columns 3*fw .. 3*fw+2 of the tile
    const long row_pitch2 = (long)p.W * p.ldo * 2;
    const int ldo2 = (int)p.ldo * 2;
    const uint32_t lsw = (uint32_t)(lane >> 2), lw = (uint32_t)((lane & 3) << 2);
    uint32_t n = 0;
    for (int i = 0; i < ntiles; ++i) {
      int img, y0, x0;
      tile_xy(i, img, y0, x0);
      for (int cb = 0; cb < ncb; ++cb, ++n) {
      const int b = n & 1;
      // this lane's channel pair of the block and its 9 (x2) packed weights (L1-resident after the first tile)
      const int ch = cb * PF_CB + lane * 2;
      const bool ch_ok = ch < p.Cout;
      u64 w[NH][9];
#pragma unroll
      for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float2 f = make_float2(0.f, 0.f);
          if (ch_ok) f = __ldg(reinterpret_cast<const float2*>(p.w9c + (long)t * p.Nt + h * hp + ch));
          w[h][t] = pack2f(f.x, f.y);
        }
      mbar_wait_relaxed(t_ready(b), (n >> 1) & 1);
#pragma unroll 1
      for (int pass = 0; pass < NPASS; ++pass) {
        const int xs = (fw * NPASS + pass) * PXT;
        // byte offsets of this lane's word in tile columns xs .. xs+PXT+1 (128B swizzle: 16-byte chunk ^ (pixel & 7); the
        // row pitch is 32 pixels, so pixel & 7 depends on the column only)
        uint32_t off[PXT + 2];
#pragma unroll
        for (int j = 0; j < PXT + 2; ++j) off[j] = (uint32_t)(xs + j) * 128 + ((lsw ^ (uint32_t)((xs + j) & 7)) << 4) + lw;
        const uint32_t tb = t_base + (b * NH) * PF_XCHUNK;
        const int nq = ch_ok ? min(p.W - x0 - xs, PF_OW - xs) : 0, nr = p.H - y0;
        uint8_t* orp = reinterpret_cast<uint8_t*>(p.out + (((long)img * p.H + y0) * p.W + x0 + xs) * p.ldo + ch);
        u64 acc[NH][3][PXT];
#pragma unroll
        for (int ir = 0; ir < PF_IH; ++ir) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            u64 v[PXT + 2];
#pragma unroll
            for (int j = 0; j < PXT + 2; ++j) v[j] = unpack2(lds32(tb + h * PF_XCHUNK + ir * (PF_TW * 128) + off[j]));
            // taps of this input row, completing output row (dy = 2) first: in the last half its GELU / gate / store chain is
            // issued right after and overlaps the independent dy = 1, 0 FMAs that follow (same per-accumulator order as before)
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const int dy = 2 - dyi;
              const int orow = ir - dy;
              if (orow >= 0 && orow < PF_OH) {
                const int a = orow % 3;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                  for (int q = 0; q < PXT; ++q) {
                    if (dy == 0 && dx == 0) acc[h][a][q] = fmul2(v[q], w[h][0]);
                    else acc[h][a][q] = ffma2(v[q + dx], w[h][dy * 3 + dx], acc[h][a][q]);
                  }
                }
              }
              if (dyi == 0 && h == NH - 1 && ir >= 2) {
                const int a = (ir - 2) % 3;
#pragma unroll
                for (int q = 0; q < PXT; ++q) {
                  const float2 f = as_float2(GATE ? gelu_gate2(acc[0][a][q], acc[NH - 1][a][q]) : acc[0][a][q]);
                  if (ir - 2 < nr && q < nq) *reinterpret_cast<uint32_t*>(orp + q * ldo2) = pack_bf16x2(f.x, f.y);
                }
                orp += row_pitch2;
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free(b));             // this warp no longer reads t[b]
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * N1));
  }
}

template <int GATE, int WB>
int launch_pf(const CUtensorMap& map_x, const CUtensorMap& map_w1, const PfParams& p, int grid, uint32_t smem, cudaStream_t s) {
  static DeviceOnce once;
  bool first; int dev;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_pwdw_f2<GATE, WB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    device_mark(once, dev);
  }
  k_pwdw_f2<GATE, WB><<<grid, PfCfg<GATE>::THREADS, smem, s>>>(map_x, map_w1, p);
  return 0;
}
}  // namespace

bool pwdw_f2_eligible(int C, int Nt, int gate) {
  return C % 16 == 0 && C >= 16 && C <= 128 && Nt % 8 == 0 && (!gate || Nt % 16 == 0);
}

// x [nimg,H,W,C] (row stride ldx) --1x1 (w1: [Nt][C] bf16, LayerNorm gamma folded), * rstd--> t --dw3x3 (w9c fp32 [9][Nt])
// --> [gate] --> out (row stride ldo)
int pwdw_f2(const bf16* x, long ldx, const float* rstd, const bf16* w1, int Nt, const float* w9c, bf16* out, long ldo, int nimg,
            int H, int W, int C, int gate, cudaStream_t s, const float* mu, const float* s1, const float* s2) {
  KD_CHECK((mu == nullptr) == (s1 == nullptr) && (mu == nullptr) == (s2 == nullptr), "pwdw_f2: mu, s1 and s2 go together");
  KD_CHECK(pwdw_f2_eligible(C, Nt, gate), "pwdw_f2: shape not eligible (C=%d Nt=%d)", C, Nt);
  KD_CHECK(!(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(out) & 3) && !(reinterpret_cast<uintptr_t>(w1) & 15) &&
               ldx % 8 == 0 && ldo % 2 == 0 && ldo < (1L << 28),
           "pwdw_f2: misaligned operands");
  PfParams p;
  p.H = H; p.W = W; p.C = C; p.Nt = Nt; p.Cout = gate ? Nt / 2 : Nt; p.nimg = nimg;
  p.kc = (C + 63) / 64;
  p.tiles_x = cdiv(W, PF_OW); p.tiles_y = cdiv(H, PF_OH); p.cblocks = cdiv(p.Cout, PF_CB);
  p.ntiles_all = (long)nimg * p.tiles_x * p.tiles_y;
  KD_CHECK(p.ntiles_all < (1L << 24), "pwdw_f2: too many tiles");
  p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.rstd = rstd; p.mu = mu; p.s1 = s1; p.s2 = s2; p.w9c = w9c; p.out = out; p.ldo = ldo;
  const int NH = gate ? 2 : 1;
  const uint32_t smem = 1024 + p.kc * NH * 8192 + p.kc * PF_XCHUNK + 2 * NH * PF_XCHUNK + 1024;
  int g_pf_sms;
  KD_TRY(device_sms(&g_pf_sms));
  KD_CHECK(smem <= 232448, "pwdw_f2: shared memory budget exceeded (%u)", smem);
  CUtensorMap map_x, map_w1;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H};
    const cuuint32_t box[4] = {64, PF_TW, PF_IH, 1};
    KD_TRY(make_map(&map_x, x, 4, dims, str, box));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Nt, 1};
    const cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Nt};
    const cuuint32_t box[3] = {64, 64, 1};
    KD_TRY(make_map(&map_w1, w1, 3, dims, str, box));
  }
  // one launch does the work of conv_gemm (1x1) + dwconv3x3: report it under its own class
  const double pix = (double)nimg * H * W;
  ProfScope prof(PC_PWDW, s, 2.0 * pix * Nt * C + 18.0 * pix * Nt, pix * (C + p.Cout) * 2.0 + 4.0 * pix + 2.0 * Nt * C);
  const int grid = (int)std::min<long>(p.ntiles_all, (long)g_pf_sms);
  int lr;
  if (mu) lr = gate ? (launch_pf<1, 1>(map_x, map_w1, p, grid, smem, s)) : (launch_pf<0, 1>(map_x, map_w1, p, grid, smem, s));
  else lr = gate ? (launch_pf<1, 0>(map_x, map_w1, p, grid, smem, s)) : (launch_pf<0, 0>(map_x, map_w1, p, grid, smem, s));
  KD_TRY(lr);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
