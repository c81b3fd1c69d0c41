// Fused  LayerNorm-folded 1x1 conv  ->  depthwise 3x3  (-> GELU gate)  on tcgen05, bf16 path.
//   qkv branch  (KDLAE_model.py:118-119):  qkv' = dw3x3(W_qkv . LN(x))
//   GDFN branch (KDLAE_model.py:95-104):   g    = gelu(dw(t)[:h]) * dw(t)[h:],  t = W_in . LN(x)
// The 3C / 2h wide intermediate t is the largest tensor of a TransformerBlock; here it only ever exists as a
// 6 x 32 pixel shared-memory tile.  Per CTA (bound to one 64-channel block of the output), per 4 x 30 pixel tile:
//   1. TMA: x tile with halo {C, 32 px, 6 rows} (zero fill outside the image = the conv's zero padding)
//   2. MMA1 (tcgen05, M=128 x2, N=64|128, K=C):  T = X . W1^T  -> TMEM        (halo rows recomputed: 1.5x of a tiny GEMM)
//   3. epilogue 1: tcgen05.ld, * rstd[pixel] (BiasFree LayerNorm folded: gamma in W1, rstd here), bf16, st.shared into
//      the 128B-swizzled K-major t tile with a row pitch of 32 pixels
//   4. MMA2: the depthwise conv as 9 shifted-descriptor MMAs (M=128, N=16, K=16) per 16-channel group against
//      diagonal weight blocks (see dwconv_tc.cu)                              -> TMEM
//   5. epilogue 2: (gate) gelu(x1)*x2, bf16, staged [4][30][64] -> one TMA store.
// Warp roles: 0 TMA producer, 1 MMA issuer (MMA1 + MMA2 of half 0), 2 MMA issuer (MMA2 of half 1, gate only),
// 3..10 epilogue.  Weights (W1 block, diagonal dw blocks) are resident in smem for the CTA's lifetime.
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int PD_TW = 32, PD_OW = 30, PD_OH = 4, PD_IH = 6, PD_CB = 64;
constexpr uint32_t PD_XCHUNK = PD_TW * PD_IH * 128;          // 24576: one 64-channel K chunk of the x tile / one t tile
constexpr uint32_t PD_DWB = 4 * 3 * 2048;                    // diagonal dw blocks per half
constexpr int PD_E1_WARPS = 4;                               // epilogue 1 (scale + convert): one warp per TMEM lane quarter
constexpr int PD_E2_WARPS = 16;                              // epilogue 2 (GELU gate, latency bound): 4 lane quarters x 4 channel quarters
constexpr int PD_THREADS = (3 + PD_E1_WARPS + PD_E2_WARPS) * 32;

struct PdParams {
  int H, W, C, Nt, Cout, nimg;     // Nt = rows of W1 (3C or 2hp); Cout = output channels (3C or hp)
  int kc;                          // 64-wide K chunks of C (1 or 2)
  int tiles_x, tiles_y, cblocks;
  long tiles_per_cb;
  float inv_tiles_x, inv_tiles_y;
  const float* rstd;               // [nimg*H*W]
};

__device__ __forceinline__ float gelu_as2(float x) {   // exact GELU via A&S 7.1.26 erf (see dwconv_tc.cu)
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float y = fmaf(1.061405429f, t, -1.453152027f);
  y = fmaf(y, t, 1.421413741f);
  y = fmaf(y, t, -0.284496736f);
  y = fmaf(y, t, 0.254829592f);
  y *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
  const float erf_x = copysignf(fmaf(-y, e, 1.0f), x);
  const float hx = 0.5f * x;
  return fmaf(hx, erf_x, hx);
}

// Two tiles are in flight per CTA: while epilogue group 2 gates / stores tile i, epilogue group 1 already converts the
// 1x1 result of tile i+1 and the tensor pipe runs MMA1(i+2) / MMA2(i+1).  t tiles and the depthwise accumulators are
// double buffered; the consumed t tile doubles as the output staging buffer of its own tile.
template <int GATE>
__global__ void __launch_bounds__(PD_THREADS, 1)
k_pwdw_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
          const __grid_constant__ CUtensorMap map_out, const uint8_t* __restrict__ wtc, const PdParams p) {
  constexpr int NH = GATE ? 2 : 1;
  constexpr int N1 = 64 * NH;                       // MMA1 N: t channels of this block (both halves for the gate)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // smem (all 1024-aligned): W1 [kc][NH][64 rows][128 B] | dw blocks | x chunks | t[2][NH] tiles (+1 KB slack) | barriers
  const uint32_t w1_base = sbase;
  const uint32_t dwb_base = w1_base + p.kc * NH * 8192;
  const uint32_t x_base = dwb_base + NH * PD_DWB;
  const uint32_t t_base = x_base + p.kc * PD_XCHUNK;
  const uint32_t bar_base = t_base + 2 * NH * PD_XCHUNK + 1024;
  const uint32_t w_bar = bar_base, x_full = bar_base + 8, x_empty = bar_base + 16, d1_full = bar_base + 24, d1_empty = bar_base + 32;
  auto t_ready = [&](int b) { return bar_base + 40 + 8u * b; };
  auto t_free = [&](int b) { return bar_base + 56 + 8u * b; };
  auto d2_full = [&](int b) { return bar_base + 72 + 8u * b; };
  auto d2_empty = [&](int b) { return bar_base + 88 + 8u * b; };
  const uint32_t tmem_slot = bar_base + 104;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* t_gen = smem_raw + (t_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = blockIdx.x % p.cblocks;
  const int cta_in_cb = blockIdx.x / p.cblocks;
  const int ctas_in_cb = (gridDim.x - cb + p.cblocks - 1) / p.cblocks;
  const int hp = p.Nt / 2;
  const int ch_valid = min(PD_CB, p.Cout - cb * PD_CB);
  const int ngroups = (ch_valid + 15) / 16;
  const int ksteps = (p.C + 15) / 16;               // K = 16 MMA steps over the C input channels
  const int ntiles = (cta_in_cb < p.tiles_per_cb) ? (int)((p.tiles_per_cb - cta_in_cb + ctas_in_cb - 1) / ctas_in_cb) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_x); prefetch_tmap(&map_w1); prefetch_tmap(&map_out);
    mbar_init(w_bar, 1); mbar_init(x_full, 1); mbar_init(x_empty, 1); mbar_init(d1_full, 1); mbar_init(d1_empty, PD_E1_WARPS);
    for (int b = 0; b < 2; ++b) { mbar_init(t_ready(b), 1); mbar_init(t_free(b), 1); mbar_init(d2_full(b), NH); mbar_init(d2_empty(b), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(GATE ? 512 : 256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t d1_col = 0;                                       // D1: 2 M-tiles x N1 columns
  auto d2_col = [&](int b) { return (uint32_t)(2 * N1 + b * NH * PD_CB); };   // D2[b]: NH x 64 columns

  auto tile_xy = [&](int i, int& img, int& y0, int& x0) {          // i-th tile of this CTA
    const int t = cta_in_cb + i * ctas_in_cb;
    const int rowt = fast_div(t, p.tiles_x, p.inv_tiles_x);
    const int txi = t - rowt * p.tiles_x;
    img = fast_div(rowt, p.tiles_y, p.inv_tiles_y);
    const int tyi = rowt - img * p.tiles_y;
    x0 = txi * PD_OW; y0 = tyi * PD_OH;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_bar, p.kc * NH * 8192 + NH * PD_DWB);
      for (int k = 0; k < p.kc; ++k)
        for (int h = 0; h < NH; ++h)
          tma_load_3d(w1_base + (k * NH + h) * 8192, &map_w1, w_bar, k * 64, (GATE ? h * hp : 0) + cb * PD_CB, 0);
      for (int h = 0; h < NH; ++h) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dwb_base + h * PD_DWB), "l"(wtc + ((size_t)h * p.cblocks + cb) * PD_DWB), "r"(PD_DWB), "r"(w_bar) : "memory");
      }
      for (int i = 0; i < ntiles; ++i) {
        int img, y0, x0;
        if (i + 1 < ntiles) {   // the x tile is single buffered: keep the next one warm in L2
          tile_xy(i + 1, img, y0, x0);
          for (int k = 0; k < p.kc; ++k) tma_prefetch_4d(&map_x, k * 64, x0 - 1, y0 - 1, img);
        }
        mbar_wait_relaxed(x_empty, (i & 1) ^ 1);
        tile_xy(i, img, y0, x0);
        mbar_expect_tx(x_full, p.kc * PD_XCHUNK);
        for (int k = 0; k < p.kc; ++k) tma_load_4d(x_base + k * PD_XCHUNK, &map_x, x_full, k * 64, x0 - 1, y0 - 1, img);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers =====================
    if (lane == 0 && (warp == 1 || GATE)) {
      const int h = warp - 1;
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lo_tag = 1u << 16;
      mbar_wait(w_bar, 0);
      auto issue_mma1 = [&](int i) {                // T = X . W1^T for both M tiles (pixels 0..127, 128..255)
        mbar_wait(x_full, i & 1);
        tc_fence_after();
        for (int mt = 0; mt < 2; ++mt) {
          for (int ks = 0; ks < ksteps; ++ks) {
            const int k = ks >> 2, kk = ks & 3;
            const uint32_t a_lo = (((x_base + k * PD_XCHUNK + mt * 16384 + kk * 32) & 0x3FFFF) >> 4) | lo_tag;
            const uint32_t b_lo = (((w1_base + k * NH * 8192 + kk * 32) & 0x3FFFF) >> 4) | lo_tag;
            umma_bf16_lohi(tmem_base + d1_col + mt * N1, a_lo, b_lo, desc_hi, idesc1, ks != 0 ? 1u : 0u);
          }
        }
        umma_commit(x_empty);      // x tile consumed
        umma_commit(d1_full);      // T accumulators ready for epilogue 1
      };
      const uint32_t b_lo0 = (((dwb_base + h * PD_DWB) & 0x3FFFF) >> 4) | lo_tag;
      if (h == 0 && ntiles > 0) issue_mma1(0);
      for (int i = 0; i < ntiles; ++i) {
        const int b = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        if (h == 0 && i + 1 < ntiles) {             // next tile's 1x1 as soon as epilogue 1 has drained D1
          mbar_wait(d1_empty, i & 1);
          issue_mma1(i + 1);
        }
        mbar_wait(t_ready(b), ph);                  // t tile of tile i written
        mbar_wait(d2_empty(b), ph ^ 1);             // D2[b] drained by epilogue 2 of tile i-2
        tc_fence_after();
        const uint32_t t_lo0 = (((t_base + (b * NH + h) * PD_XCHUNK) & 0x3FFFF) >> 4) | lo_tag;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g < ngroups) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3, dx = tap % 3;
              const uint32_t a_lo = t_lo0 + (uint32_t)((dy * PD_TW + dx) * 8 + g * 2);
              const uint32_t b_lo = b_lo0 + (uint32_t)((g * 6144 + (tap >> 2) * 2048 + (tap & 3) * 32) >> 4);
              umma_bf16_lohi(tmem_base + d2_col(b) + h * PD_CB + g * 16, a_lo, b_lo, desc_hi, idesc2, tap != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(d2_full(b));
      }
    }
  } else if (warp < 3 + PD_E1_WARPS) {
    // ===================== epilogue group 1: T (fp32, TMEM) * rstd -> bf16 t tile =====================
    const int ew = warp - 3;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;        // TMEM lane = row of the M tile
    for (int i = 0; i < ntiles; ++i) {
      const int b = i & 1;
      int img, y0, x0;
      tile_xy(i, img, y0, x0);
      float rs[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int pix = mt * 128 + r;           // pixel of the 6 x 32 halo tile
        const int y = y0 - 1 + pix / PD_TW, x = x0 - 1 + pix % PD_TW;
        const bool inimg = pix < PD_TW * PD_IH && y >= 0 && y < p.H && x >= 0 && x < p.W;
        rs[mt] = inimg ? __ldg(p.rstd + ((long)img * p.H + y) * p.W + x) : 0.f;   // 0 outside: conv zero padding of t
      }
      mbar_wait_relaxed(t_free(b), ((i >> 1) & 1) ^ 1);   // t[b] no longer read by MMA2(i-2) / its output store
      mbar_wait_relaxed(d1_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const bool live = !(mt == 1 && quarter >= 2);      // pixels 192..255 do not exist (warp-uniform)
        const int pix = mt * 128 + r;
        const uint32_t t_row = tmem_base + d1_col + mt * N1 + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // two runs of 32 channels per chunk(2) half
          uint32_t v[NH * 2][16];
          if (live) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              tmem_ld16_issue(t_row + h * 64 + half * 32, v[h * 2]);
              tmem_ld16_issue(t_row + h * 64 + half * 32 + 16, v[h * 2 + 1]);
            }
#pragma unroll
            for (int u = 0; u < NH * 2; ++u) tmem_ld16_wait(v[u]);
          }
          if (mt == 1 && half == 1) {             // all TMEM reads of this tile done: MMA1 of the next tile may overwrite D1
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_empty);
          }
          if (live) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              uint8_t* trow = t_gen + (b * NH + h) * PD_XCHUNK + pix * 128;
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t (&vv)[16] = v[h * 2 + (c4 >> 1)];
                const int o = (c4 & 1) * 8;
                uint4 w4;
                w4.x = pack_bf16x2(__uint_as_float(vv[o + 0]) * rs[mt], __uint_as_float(vv[o + 1]) * rs[mt]);
                w4.y = pack_bf16x2(__uint_as_float(vv[o + 2]) * rs[mt], __uint_as_float(vv[o + 3]) * rs[mt]);
                w4.z = pack_bf16x2(__uint_as_float(vv[o + 4]) * rs[mt], __uint_as_float(vv[o + 5]) * rs[mt]);
                w4.w = pack_bf16x2(__uint_as_float(vv[o + 6]) * rs[mt], __uint_as_float(vv[o + 7]) * rs[mt]);
                const int chunk = half * 4 + c4;
                *reinterpret_cast<uint4*>(trow + ((chunk ^ (pix & 7)) << 4)) = w4;
              }
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, %0;" ::"n"(PD_E1_WARPS * 32) : "memory");          // t tile complete
      if (ew == 0 && lane == 0) mbar_arrive(t_ready(b));
    }
  } else {
    // ===================== epilogue group 2: depthwise result -> (gate) -> staged in the consumed t tile -> TMA store =====
    const int ew = warp - 3 - PD_E1_WARPS;
    const int quarter = warp & 3;
    const int cq = ew >> 2;                   // 16 of the 64 channels
    const int r = quarter * 32 + lane;
    const int oy_l = r / PD_TW, ox_l = r % PD_TW;
    const int opix = oy_l * PD_OW + ox_l;
    const bool in_box = ox_l < PD_OW;
    for (int i = 0; i < ntiles; ++i) {
      const int b = i & 1;
      mbar_wait_relaxed(d2_full(b), (i >> 1) & 1);    // all MMA2(i) retired: D2[b] valid, t[b] no longer read
      tc_fence_after();
      const uint32_t t_row = tmem_base + d2_col(b) + ((uint32_t)(quarter * 32) << 16) + cq * 16;
      uint8_t* srow = t_gen + (b * NH) * PD_XCHUNK + opix * 128;     // staging = half-0 t tile of this buffer
      if (cq * 16 < ch_valid) {                       // warp-uniform
        uint32_t a[16], bb[16];
        tmem_ld16_issue(t_row, a);
        if (GATE) tmem_ld16_issue(t_row + PD_CB, bb);
        tmem_ld16_wait(a);
        if (GATE) tmem_ld16_wait(bb);
        if (in_box) {
#pragma unroll
          for (int v8 = 0; v8 < 2; ++v8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float x1 = __uint_as_float(a[v8 * 8 + e]);
              f[e] = GATE ? gelu_fast(x1) * __uint_as_float(bb[v8 * 8 + e]) : x1;
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
            const int chunk = cq * 2 + v8;
            *reinterpret_cast<uint4*>(srow + ((chunk ^ (opix & 7)) << 4)) = o;
          }
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 2, %0;" ::"n"(PD_E2_WARPS * 32) : "memory");   // D2[b] drained by the whole group, output tile staged
      if (ew == 0 && lane == 0) {
        mbar_arrive(d2_empty(b));
        int img, y0, x0;
        tile_xy(i, img, y0, x0);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(&map_out), "r"(t_base + (b * NH) * PD_XCHUNK), "r"(cb * PD_CB), "r"(x0), "r"(y0), "r"(img) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(t_free(b));                        // t[b] may be rewritten by epilogue 1 of tile i+2
      }
    }
    if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(GATE ? 512 : 256));
  }
}

int g_pd_sms = 0;

}  // namespace

bool pwdw_tc_eligible(int C, int Nt, int gate) {
  return C % 16 == 0 && C >= 16 && C <= 128 && Nt % 8 == 0 && (!gate || Nt % 16 == 0);
}

// x [nimg,H,W,C] (row stride ldx) --1x1 (w1: [Nt][C] bf16, LayerNorm gamma folded), * rstd--> t --dw3x3 (wtc)--> [gate] --> out
int pwdw_tc(const bf16* x, long ldx, const float* rstd, const bf16* w1, int Nt, const void* wtc, bf16* out, long ldo, int nimg,
            int H, int W, int C, int gate, cudaStream_t s) {
  KD_CHECK(pwdw_tc_eligible(C, Nt, gate), "pwdw_tc: shape not eligible (C=%d Nt=%d)", C, Nt);
  KD_CHECK(!(reinterpret_cast<uintptr_t>(x) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15) && !(reinterpret_cast<uintptr_t>(w1) & 15) &&
               !(reinterpret_cast<uintptr_t>(wtc) & 15) && ldx % 8 == 0 && ldo % 8 == 0,
           "pwdw_tc: misaligned operands");
  PdParams p;
  p.H = H; p.W = W; p.C = C; p.Nt = Nt; p.Cout = gate ? Nt / 2 : Nt; p.nimg = nimg;
  p.kc = (C + 63) / 64;
  p.tiles_x = cdiv(W, PD_OW); p.tiles_y = cdiv(H, PD_OH); p.cblocks = cdiv(p.Cout, PD_CB);
  p.tiles_per_cb = (long)nimg * p.tiles_x * p.tiles_y;
  KD_CHECK(p.tiles_per_cb < (1L << 24), "pwdw_tc: too many tiles");
  p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.rstd = rstd;
  const int NH = gate ? 2 : 1;
  const uint32_t smem = 1024 + p.kc * NH * 8192 + NH * PD_DWB + p.kc * PD_XCHUNK + 2 * NH * PD_XCHUNK + 1024 + 128;
  static bool attr = false;
  if (!attr) {
    int dev = 0;
    KD_CUDA(cudaGetDevice(&dev));
    KD_CUDA(cudaDeviceGetAttribute(&g_pd_sms, cudaDevAttrMultiProcessorCount, dev));
    g_pd_sms = sm_limit(g_pd_sms);
    KD_CUDA(cudaFuncSetAttribute(k_pwdw_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    KD_CUDA(cudaFuncSetAttribute(k_pwdw_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  KD_CHECK(smem <= 232448, "pwdw_tc: shared memory budget exceeded (%u)", smem);
  CUtensorMap map_x, map_w1, map_out;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H};
    const cuuint32_t box[4] = {64, PD_TW, PD_IH, 1};
    KD_TRY(make_map(&map_x, x, 4, dims, str, box));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Nt, 1};
    const cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * Nt};
    const cuuint32_t box[3] = {64, 64, 1};
    KD_TRY(make_map(&map_w1, w1, 3, dims, str, box));
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)ldo * 2 * W, (cuuint64_t)ldo * 2 * W * H};
    const cuuint32_t box[4] = {PD_CB, PD_OW, PD_OH, 1};
    KD_TRY(make_map(&map_out, out, 4, dims, str, box));
  }
  // one launch does the work of conv_gemm (1x1) + dwconv3x3: report it under its own class
  const double pix = (double)nimg * H * W;
  ProfScope prof(PC_PWDW, s, 2.0 * pix * Nt * C + 18.0 * pix * Nt, pix * (C + p.Cout) * 2.0 + 4.0 * pix + 2.0 * Nt * C);
  const int blocks_per_sm = 1;
  int grid = (int)std::min<long>((long)p.cblocks * p.tiles_per_cb, (long)g_pd_sms * blocks_per_sm);
  if (grid < p.cblocks) grid = p.cblocks;
  if (gate) k_pwdw_tc<1><<<grid, PD_THREADS, smem, s>>>(map_x, map_w1, map_out, reinterpret_cast<const uint8_t*>(wtc), p);
  else k_pwdw_tc<0><<<grid, PD_THREADS, smem, s>>>(map_x, map_w1, map_out, reinterpret_cast<const uint8_t*>(wtc), p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
