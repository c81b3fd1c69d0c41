// Fused  LayerNorm-folded 1x1 conv  ->  depthwise 3x3  (-> GELU gate), bf16 path, "transposed" schedule:
//   qkv branch  (KDLAE_model.py:118-119):  qkv' = dw3x3(W_qkv . LN(x))
//   GDFN branch (KDLAE_model.py:95-104):   g    = gelu(dw(t)[:h]) * dw(t)[h:],  t = W_in . LN(x)
// The 1x1 conv runs on tcgen05 as  T^T = W1 . X^T  (M = 128 output channels, N = the pixels of a halo tile, K = C), so the
// fp32 accumulator sits in TMEM with lane = channel and column = pixel.  That is exactly the orientation the depthwise
// stage wants: a thread owns one channel (two with the gate: lane j holds x1-channel j in one column range and the matching
// x2-channel j in another), pulls a row of pixels with tcgen05.ld (a second load shifted by one column gives the odd-aligned
// register pairs), and runs the 3x3 taps as packed FFMA2 over pixel pairs with the tap weight broadcast.
// Compared with pwdw_f2.cu (GEMM in the usual orientation, bf16 t tile in shared memory, channel-pair lanes) this removes the
// TMEM->bf16->smem conversion pass, the t tile, the LDS + bf16 unpack of every input (which was ~40 % of the issue slots)
// and keeps t in fp32.  LayerNorm's rstd[pixel] is applied to the x tile BEFORE the GEMM: two helper warps rescale the TMA-staged
// bf16 tile in place (x^ = bf16(x * rstd), one extra rounding of the GEMM input; a pixel row is 128 contiguous bytes of the
// swizzled tile, so the pass is linear), which costs C values per pixel instead of the 3C / 2h accumulator values the depthwise
// warps used to scale (2 FMUL2 + 2 LDS.64 of every 13 issue slots per output pair).  Pixels outside the image are TMA
// zero fill, so t is exactly 0 there = the conv's zero padding of t.
// Results leave through per-warp [pixel][32 ch] bf16 staging blocks and one small TMA store per warp and item (TMA clips at
// the image edge), so the depthwise warps never synchronise with each other: with CTA-wide barriers per item they ran in
// lockstep, all in their FFMA2 bursts at once (ncu: math-pipe throttle on top) and all idle together in the store phase.
//   GATE = 0: tile = 8 x 32 pixels (6 x 30 outputs), item = (tile, 128 channels), D = 256 TMEM columns, double buffered.
//   GATE = 1: tile = 7 x 18 pixels (5 x 16 outputs), item = (tile, 128 gated channels): x1 rows -> columns [0,128),
//             x2 rows -> columns [128,256), double buffered.
// Warps: 0 TMA producer, 1 MMA issuer, 2-3 x-tile rescale, 4.. depthwise (TMEM lane quarter = warp % 4).
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
__device__ __forceinline__ u64 pack2u(uint32_t lo, uint32_t hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ u64 splat2(float c) { return pack2f(c, c); }
// gelu(a) * b on a packed pair (identical to dwconv_f2.cu / pwdw_f2.cu)
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

// tcgen05.ld 32x32b of 2 / 4 / 8 consecutive columns (this thread's TMEM lane)
__device__ __forceinline__ void tld2(uint32_t a, uint32_t& r0, uint32_t& r1) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ void tld4(uint32_t a, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void tld8(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
}

// tcgen05.wait::ld that names the loaded registers as in/out operands, so no consumer can be scheduled above it
template <int N>
__device__ __forceinline__ void tld_wait(uint32_t (&r)[N]) {
  static_assert(N == 10 || N == 12, "row width");
  if constexpr (N == 12) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]) :: "memory");
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9])
                 :: "memory");
  }
}

constexpr int PT_MB = 128;                           // channels per item
template <int GATE> struct PtCfg {
  // The gate tile is 7 x 18 pixels (5 x 16 outputs): its 126 pixels fit the 128 TMEM columns a half may use, and the 16 output
  // columns split into two groups of four pixel pairs with nothing left over (the earlier 8 x 16 tile gave 14 columns = 8 + 6,
  // and the second group ran a fourth, discarded pair: 1/8 of the gate kernel's arithmetic).
  static constexpr int IH = GATE ? 7 : 8;            // halo rows
  static constexpr int OH = IH - 2;                  // output rows
  static constexpr int PITCH = GATE ? 18 : 32;       // pixels per tile row (= TMEM columns per row)
  static constexpr int OW = PITCH - 2;               // output columns per tile
  static constexpr int NPX_REAL = PITCH * IH;        // pixels per tile
  static constexpr int NPX = (NPX_REAL + 15) / 16 * 16;   // GEMM N (gate: 126 -> 128, the two extra columns are never read)
  static constexpr int NG = GATE ? 2 : 3;            // column groups of depthwise warps
  static constexpr int GP = GATE ? 4 : 5;            // output pixel pairs per group per row.  (Gate with 16 warps of two pairs,
                                                     // 93 registers: 96 -> 2x256 6 % slower, 48 -> 2x128 2.6 % faster: the narrower
                                                     // groups load 6 columns per 4 outputs.  Gate without its four MUFU per pair:
                                                     // 5 % faster - the kernel is bound by the FMA pipe / issue mix, not by MUFU.)
  static constexpr int NDW = NG * 4;                 // depthwise warps
  static constexpr int THREADS = (4 + NDW) * 32;
  static constexpr uint32_t XCHUNK = NPX * 128;      // one 64-channel K chunk of the x tile
  static constexpr uint32_t XLOAD = NPX_REAL * 128;  // bytes TMA delivers into it
  static constexpr uint32_t WSTAGE = 2 * GP * OH * 64;   // per-warp staging block [row][column][32 ch] bf16
  static constexpr uint32_t STAGE = NDW * WSTAGE;
  static_assert(NG * 2 * GP == OW, "column groups must tile the output columns");
};

struct PtParams {
  int H, W, C, Nt, Cout, nimg;     // Nt = rows of W1 (3C or 2hp); Cout = output channels (3C or hp)
  int kc;                          // 64-wide K chunks of C (1 or 2)
  int tiles_x, tiles_y, ncb;
  long ntiles_all;
  float inv_tiles_x, inv_tiles_y;
  const float* rstd;               // [nimg*H*W]
  const float* w9c;                // depthwise weights fp32 [9][Nt]
  int xbufs;                       // x tile buffers: 2 when shared memory allows (the next tile loads and rescales under this one)
  int dbg;                         // bring-up probes (KDLAE_PT_DBG): 1 = skip the output stores, 2 = skip the depthwise arithmetic
};

template <int GATE>
__global__ void __launch_bounds__(PtCfg<GATE>::THREADS, 1)
k_pwdw_t(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
         const __grid_constant__ CUtensorMap map_out, const PtParams p) {
  typedef PtCfg<GATE> Cfg;
  constexpr int NH = GATE ? 2 : 1;
  constexpr int PITCH = Cfg::PITCH, OW = Cfg::OW, NPX = Cfg::NPX, NG = Cfg::NG, GP = Cfg::GP, GW = 2 * GP;
  constexpr int PT_IH = Cfg::IH, PT_OH = Cfg::OH;
  constexpr uint32_t XCHUNK = Cfg::XCHUNK, STAGE = Cfg::STAGE, WSTAGE = Cfg::WSTAGE;
  constexpr uint32_t W1CHUNK = PT_MB * 128;          // 128 rows x 64 channels of one chunk(2) half
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // smem (1024-aligned): W1 [kc][NH] chunks | x chunks | staging[2] | barriers
  // W1 blocks are double buffered and the x tile too where it fits: with single buffers the per-item chain "GEMM retires ->
  // W1 TMA -> GEMM" and the per-tile chain "GEMM retires -> x TMA -> rescale -> GEMM" ran back to back and the producer side
  // alone took 55-70 % of the kernel time (KDLAE_PT_DBG=2 probe).
  const uint32_t w1_bytes = p.kc * NH * W1CHUNK, x_bytes = p.kc * XCHUNK;
  const uint32_t w1_base = sbase;
  const uint32_t x_base = w1_base + 2 * w1_bytes;
  const uint32_t stage_base = x_base + p.xbufs * x_bytes;
  const uint32_t bar_base = stage_base + 2 * STAGE;
  auto w_full = [&](int b) { return bar_base + 8u * b; };
  auto w_empty = [&](int b) { return bar_base + 16 + 8u * b; };
  auto x_full = [&](int b) { return bar_base + 112 + 8u * b; };
  auto x_empty = [&](int b) { return bar_base + 128 + 8u * b; };
  // (Handing D[b] over in two-row chunks - the GEMM of item n+2 refilling rows the depthwise warps have moved past - was measured
  // 8-35 % SLOWER: four N = 64 MMAs re-read the A operand four times and every chunk costs each warp a tcgen05 fence pair.)
  auto d_full = [&](int b) { return bar_base + 32 + 8u * b; };
  auto d_empty = [&](int b) { return bar_base + 48 + 8u * b; };
  auto x_scaled = [&](int b) { return bar_base + 64 + 8u * b; };
  const uint32_t tmem_slot = bar_base + 96;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* x_gen = smem_raw + (x_base - smem_u32(smem_raw));
  uint8_t* stage_gen = smem_raw + (stage_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hp = p.Nt / 2;
  const int ksteps = (p.C + 15) / 16;
  const int ntiles = ((long)blockIdx.x < p.ntiles_all) ? (int)((p.ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const int ncb = p.ncb;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_w1); prefetch_tmap(&map_out);
    for (int b = 0; b < 2; ++b) {
      mbar_init(w_full(b), 1); mbar_init(w_empty(b), 1); mbar_init(x_full(b), 1); mbar_init(x_empty(b), 1); mbar_init(x_scaled(b), 2);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(d_full(b), 1); mbar_init(d_empty(b), Cfg::NDW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
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
    x0 = txi * OW; y0 = tyi * PT_OH;
  };

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, elected issuer: see elect_one()) =====================
    {
      uint32_t n = 0;
      for (int i = 0; i < ntiles; ++i) {
        int img, y0, x0;
        if (i + 1 < ntiles) {   // the x tile is single buffered: keep the next one warm in L2
          tile_xy(i + 1, img, y0, x0);
          if (elect_one())
            for (int k = 0; k < p.kc; ++k) tma_prefetch_4d(&map_x, k * 64, x0 - 1, y0 - 1, img);
          __syncwarp();
        }
        const int xb = (p.xbufs == 2) ? (i & 1) : 0;
        const uint32_t xph = (p.xbufs == 2) ? ((i >> 1) & 1) : (i & 1);
        mbar_wait_relaxed(x_empty(xb), xph ^ 1);     // last GEMM that read this x buffer has retired
        tile_xy(i, img, y0, x0);
        if (elect_one()) {
          mbar_expect_tx(x_full(xb), p.kc * Cfg::XLOAD);
          for (int k = 0; k < p.kc; ++k) tma_load_4d(x_base + xb * x_bytes + k * XCHUNK, &map_x, x_full(xb), k * 64, x0 - 1, y0 - 1, img);
        }
        __syncwarp();
        for (int cb = 0; cb < ncb; ++cb, ++n) {
          const int wb = n & 1;
          mbar_wait_relaxed(w_empty(wb), ((n >> 1) & 1) ^ 1);      // GEMM n-2 retired: this W1 buffer is free
          // W1 rows arrive per TMEM lane quarter (32 rows): a partial last block only loads its valid quarters, rotated by the
          // tile index so that the extra work lands on a different scheduler every tile (see the depthwise warps)
          const int nvalid = min(PT_MB, p.Cout - cb * PT_MB);
          const int nq = (nvalid + 31) >> 5, rot = (nq < 4) ? (i & 3) : 0;
          if (elect_one()) {
            mbar_expect_tx(w_full(wb), p.kc * NH * nq * 4096);
            for (int k = 0; k < p.kc; ++k)
              for (int h = 0; h < NH; ++h)
                for (int lq = 0; lq < nq; ++lq)
                  tma_load_3d(w1_base + wb * w1_bytes + (k * NH + h) * W1CHUNK + ((lq + rot) & 3) * 4096, &map_w1, w_full(wb), k * 64,
                              (GATE ? h * hp : 0) + cb * PT_MB + lq * 32, 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: T^T = W1 . X^T (A = W1 rows, B = pixels); warp-uniform loop, elected issuer ============
    {
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lo_tag = 1u << 16;
      uint32_t n = 0;
      for (int i = 0; i < ntiles; ++i) {
        for (int cb = 0; cb < ncb; ++cb, ++n) {
          const int b = n & 1;
          mbar_wait_relaxed(d_empty(b), ((n >> 1) & 1) ^ 1);   // depthwise warps have drained D[b] (item n-2): on the critical path
          const int wb = n & 1;
          const int xb = (p.xbufs == 2) ? (i & 1) : 0;
          mbar_wait_relaxed(w_full(wb), (n >> 1) & 1);
          if (cb == 0) mbar_wait_relaxed(x_scaled(xb), (p.xbufs == 2) ? ((i >> 1) & 1) : (i & 1));   // x tile landed and rescaled by rstd
          tc_fence_after();
          if (elect_one()) {
            for (int h = 0; h < NH; ++h) {
              for (int ks = 0; ks < ksteps; ++ks) {
                const int k = ks >> 2, kk = ks & 3;
                const uint32_t a_lo = (((w1_base + wb * w1_bytes + (k * NH + h) * W1CHUNK + kk * 32) & 0x3FFFF) >> 4) | lo_tag;
                const uint32_t b_lo = (((x_base + xb * x_bytes + k * XCHUNK + kk * 32) & 0x3FFFF) >> 4) | lo_tag;
                umma_bf16_lohi(tmem_base + b * 256 + h * NPX, a_lo, b_lo, desc_hi, idesc, ks != 0 ? 1u : 0u);
              }
            }
            umma_commit(w_empty(wb));                       // W1 block consumed
            if (cb == ncb - 1) umma_commit(x_empty(xb));    // x tile consumed
            umma_commit(d_full(b));                     // accumulators ready
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== x-tile rescale: x^[p][:] = bf16(x[p][:] * rstd[p]) in place (BiasFree LayerNorm folded) ==========
    // A pixel row is 128 contiguous bytes per 64-channel chunk (the 128B swizzle only permutes the 16-byte pieces inside it), so
    // thread t owns the physical 16-byte piece (t & 7) of pixels (t >> 3) + 8 m.
    const int t64 = (warp - 2) * 32 + lane;
    const int pc = t64 & 7, p0 = t64 >> 3;
    for (int i = 0; i < ntiles; ++i) {
      int img, y0, x0;
      tile_xy(i, img, y0, x0);
      float rsv[NPX / 8];
#pragma unroll
      for (int m = 0; m < NPX / 8; ++m) {
        const int pix = p0 + 8 * m;
        const int y = y0 - 1 + pix / PITCH, x = x0 - 1 + pix % PITCH;
        const bool inimg = y >= 0 && y < p.H && x >= 0 && x < p.W && pix < Cfg::NPX_REAL;
        rsv[m] = inimg ? __ldg(p.rstd + ((long)img * p.H + y) * p.W + x) : 0.f;
      }
      const int xb = (p.xbufs == 2) ? (i & 1) : 0;
      mbar_wait_relaxed(x_full(xb), (p.xbufs == 2) ? ((i >> 1) & 1) : (i & 1));
      for (int k = 0; k < p.kc; ++k) {
        const int vch = min(8, (p.C - k * 64 + 7) >> 3);          // 16-byte pieces of this chunk that hold channels
#pragma unroll
        for (int m = 0; m < NPX / 8; ++m) {
          const int pix = p0 + 8 * m;
          if ((pc ^ (pix & 7)) < vch) {
            uint4* q = reinterpret_cast<uint4*>(x_gen + xb * x_bytes + k * XCHUNK + pix * 128 + pc * 16);
            uint4 v = *q;
            const u64 r2 = splat2(rsv[m]);
            uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = as_float2(fmul2(pack2u(w[e] << 16, w[e] & 0xffff0000u), r2));
              __nv_bfloat162 h = __floats2bfloat162_rn(f.x, f.y);
              w[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            *q = v;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA's operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(x_scaled(xb));
    }
  } else if (warp >= 4) {
    // ===================== depthwise warps =====================
    const int dwp = warp - 4;
    const int quarter = warp & 3;                 // TMEM lane quarter (== dwp & 3)
    const int g = dwp >> 2;                       // column group
    const int s0 = g * GW;                        // first input column (tile coordinates) of this group
    constexpr int gw = GW;                        // output columns of every group
    uint32_t n = 0;
    uint32_t nstored = 0;                         // items this warp has stored: picks its staging block (see below)
    for (int i = 0; i < ntiles; ++i) {
      int img, y0, x0;
      tile_xy(i, img, y0, x0);
      for (int cb = 0; cb < ncb; ++cb, ++n) {
        const int b = n & 1;
        // logical quarter of this warp in the item's channel block (partial blocks are rotated by the tile index); warps whose
        // quarter holds no channel skip the item, so the depthwise work is proportional to the valid channels
        const int nvalid = min(PT_MB, p.Cout - cb * PT_MB);
        const int nq = (nvalid + 31) >> 5, rot = (nq < 4) ? (i & 3) : 0;
        const int lq = (quarter - rot) & 3;
        if (lq >= nq) {
          mbar_wait_relaxed(d_full(b), (n >> 1) & 1);     // keeps the d_empty phase accounting in step
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty(b));
          continue;
        }
        if (p.dbg & 2) {                                     // probe: producer / GEMM pipeline alone
          mbar_wait_relaxed(d_full(b), (n >> 1) & 1);
          tc_fence_after();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty(b));
          continue;
        }
        const int ch = cb * PT_MB + lq * 32 + lane;
        const bool ch_ok = ch < p.Cout;
        u64 w[NH][9];
#pragma unroll
        for (int h = 0; h < NH; ++h)
#pragma unroll
          for (int t = 0; t < 9; ++t) w[h][t] = splat2(ch_ok ? __ldg(p.w9c + (long)t * p.Nt + h * hp + ch) : 0.f);
        mbar_wait_relaxed(d_full(b), (n >> 1) & 1);
        tc_fence_after();
        // The staging block alternates with the items this warp STORES (not with the item index: a warp that skips the partial
        // channel block of every tile would otherwise reuse the same block for consecutive stores while `wait_group.read 1` still
        // allows the previous store to be reading it - a timing-dependent corruption seen on cold first launches).  Block sb was
        // handed to TMA two stores ago: wait until that store has read it.
        const uint32_t sb = nstored & 1;
        ++nstored;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        const uint32_t trow = tmem_base + b * 256 + ((uint32_t)(quarter * 32) << 16) + s0;
        uint8_t* st = stage_gen + sb * STAGE + dwp * WSTAGE + lane * 2;
        u64 acc[NH][3][GP];
        // Every accumulator value is pulled from TMEM ONCE (columns s0 .. s0+GW+1 of a row); the odd-aligned pairs of the middle
        // tap are formed with register moves (a second, one-column-shifted tcgen05.ld was tried first: same speed, 2.2x the TMEM
        // reads).  The loads are software pipelined: the row for step k+1 is requested before the FFMA2s of step k and waited
        // for after them, so the TMEM latency hides under the arithmetic instead of stalling the warp eight times per item.
        auto issue_row = [&](int step, uint32_t (&r)[GW + 2]) {
          const uint32_t ta = trow + (step % NH) * NPX + (step / NH) * PITCH;
          if (GATE) {
            tld8(ta, r);
            tld2(ta + 8, r[8], r[9]);
          } else {
            tld8(ta, r);
            tld4(ta + 8, r[8], r[9], r[10], r[11]);
          }
        };
        uint32_t rcur[GW + 2];
        issue_row(0, rcur);
        tld_wait(rcur);
#pragma unroll
        for (int ir = 0; ir < PT_IH; ++ir) {
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            const int step = ir * NH + h;
            uint32_t rnext[GW + 2];
            if (step + 1 < PT_IH * NH) issue_row(step + 1, rnext);
            u64 va[GP + 1], vb[GP];
#pragma unroll
            for (int j = 0; j < GP + 1; ++j) va[j] = pack2u(rcur[2 * j], rcur[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < GP; ++j) vb[j] = pack2u(rcur[2 * j + 1], rcur[2 * j + 2]);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int orow = ir - dy;
              if (orow >= 0 && orow < PT_OH) {
                const int a = orow % 3;
#pragma unroll
                for (int j = 0; j < GP; ++j) {
                  if (dy == 0) acc[h][a][j] = fmul2(va[j], w[h][0]);
                  else acc[h][a][j] = ffma2(va[j], w[h][dy * 3], acc[h][a][j]);
                  acc[h][a][j] = ffma2(vb[j], w[h][dy * 3 + 1], acc[h][a][j]);
                  acc[h][a][j] = ffma2(va[j + 1], w[h][dy * 3 + 2], acc[h][a][j]);
                }
              }
            }
            if (step + 1 < PT_IH * NH) {
              tld_wait(rnext);
#pragma unroll
              for (int j = 0; j < GW + 2; ++j) rcur[j] = rnext[j];
            }
          }
          if (ir == PT_IH - 1) {                  // all TMEM reads of this item done: the GEMM of item n+2 may overwrite D[b]
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d_empty(b));
          }
          if (ir >= 2) {
            const int orow = ir - 2, a = orow % 3;
#pragma unroll
            for (int j = 0; j < GP; ++j) {
              const float2 f = as_float2(GATE ? gelu_gate2(acc[0][a][j], acc[NH - 1][a][j]) : acc[0][a][j]);
              // output columns 2j, 2j+1 of this group
              *reinterpret_cast<__nv_bfloat16*>(st + (orow * gw + 2 * j) * 64) = __float2bfloat16_rn(f.x);
              *reinterpret_cast<__nv_bfloat16*>(st + (orow * gw + 2 * j + 1) * 64) = __float2bfloat16_rn(f.y);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the TMA store
        __syncwarp();
        if (lane == 0 && !(p.dbg & 1)) {
          const CUtensorMap* mo = &map_out;
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                       ::"l"(mo), "r"(stage_base + sb * STAGE + dwp * WSTAGE), "r"(cb * PT_MB + lq * 32), "r"(x0 + s0), "r"(y0), "r"(img)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int GATE>
int launch_pt(const bf16* x, long ldx, const float* rstd, const bf16* w1, int Nt, const float* w9c, bf16* out, long ldo, int nimg, int H,
              int W, int C, cudaStream_t s) {
  typedef PtCfg<GATE> Cfg;
  PtParams p;
  p.H = H; p.W = W; p.C = C; p.Nt = Nt; p.Cout = GATE ? Nt / 2 : Nt; p.nimg = nimg;
  p.kc = (C + 63) / 64;
  p.tiles_x = cdiv(W, Cfg::OW); p.tiles_y = cdiv(H, Cfg::OH); p.ncb = cdiv(p.Cout, PT_MB);
  p.ntiles_all = (long)nimg * p.tiles_x * p.tiles_y;
  KD_CHECK(p.ntiles_all < (1L << 24), "pwdw_t: too many tiles");
  p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.rstd = rstd; p.w9c = w9c;
  { const char* e = getenv("KDLAE_PT_DBG"); p.dbg = e ? atoi(e) : 0; }
  const int NH = GATE ? 2 : 1;
  const uint32_t fixed = 1024 + 2 * p.kc * NH * PT_MB * 128 + 2 * Cfg::STAGE + 256;
  p.xbufs = (fixed + 2 * p.kc * Cfg::XCHUNK <= 232448) ? 2 : 1;
  const uint32_t smem = fixed + p.xbufs * p.kc * Cfg::XCHUNK;
  KD_CHECK(smem <= 232448, "pwdw_t: shared memory budget exceeded (%u)", smem);
  static DeviceOnce once;
  bool first; int dev, g_pt_sms;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_pwdw_t<GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    device_mark(once, dev);
  }
  KD_TRY(device_sms(&g_pt_sms));
  CUtensorMap map_x, map_w1, map_out;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H};
    const cuuint32_t box[4] = {64, (cuuint32_t)Cfg::PITCH, (cuuint32_t)Cfg::IH, 1};
    KD_TRY(make_map(&map_x, x, 4, dims, str, box));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Nt, 1};
    const cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Nt};
    const cuuint32_t box[3] = {64, 32, 1};      // one TMEM lane quarter of W1 rows per load
    KD_TRY(make_map(&map_w1, w1, 3, dims, str, box));
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)ldo * 2 * W, (cuuint64_t)ldo * 2 * W * H};
    const cuuint32_t box[4] = {32, (cuuint32_t)(2 * Cfg::GP), (cuuint32_t)Cfg::OH, 1};
    KD_TRY(make_map(&map_out, out, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));
  }
  const double pix = (double)nimg * H * W;
  ProfScope prof(PC_PWDW, s, 2.0 * pix * Nt * C + 18.0 * pix * Nt, pix * (C + p.Cout) * 2.0 + 4.0 * pix + 2.0 * Nt * C);
  const int grid = (int)std::min<long>(p.ntiles_all, (long)g_pt_sms);
  k_pwdw_t<GATE><<<grid, Cfg::THREADS, smem, s>>>(map_x, map_w1, map_out, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool pwdw_t_eligible(int C, int Nt, int gate) {
  return C % 16 == 0 && C >= 16 && C <= 128 && Nt % 8 == 0 && (!gate || Nt % 16 == 0);
}

// x [nimg,H,W,C] (row stride ldx) --1x1 (w1: [Nt][C] bf16, LayerNorm gamma folded), * rstd--> t --dw3x3 (w9c fp32 [9][Nt])
// --> [gate] --> out (row stride ldo)
int pwdw_t(const bf16* x, long ldx, const float* rstd, const bf16* w1, int Nt, const float* w9c, bf16* out, long ldo, int nimg, int H,
           int W, int C, int gate, cudaStream_t s) {
  KD_CHECK(pwdw_t_eligible(C, Nt, gate), "pwdw_t: shape not eligible (C=%d Nt=%d)", C, Nt);
  KD_CHECK(!(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15) && !(reinterpret_cast<uintptr_t>(w1) & 15) &&
               ldx % 8 == 0 && ldo % 8 == 0,
           "pwdw_t: misaligned operands");
  return gate ? launch_pt<1>(x, ldx, rstd, w1, Nt, w9c, out, ldo, nimg, H, W, C, s)
              : launch_pt<0>(x, ldx, rstd, w1, Nt, w9c, out, ldo, nimg, H, W, C, s);
}

}  // namespace kd
