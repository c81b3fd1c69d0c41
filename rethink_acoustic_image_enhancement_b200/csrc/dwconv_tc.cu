// Depthwise 3x3 (+ GELU gate) on the tensor cores (KDLAE_model.py:97,103-104,119), bf16 path.
//
// The CUDA-core version is FP32-issue bound (18 FMA + unpack per 2 bytes) while tcgen05 idles, so the conv is recast
// as an implicit GEMM with block-diagonal weights:  for a group of 16 channels
//      D[128 pixels x 16 ch] = sum_{tap=0..8}  A_tap[128 pixels x 16 ch] * diag(w[tap][16 ch])
// i.e. nine tcgen05.mma (M=128, N=16, K=16) per group.  15/16 of the MACs multiply zeros, which is free: the tensor
// pipe has > 10x headroom here (8 cycles per MMA).
//   * A: one TMA box {64 ch, 32 px, 6 rows} per tile lands as a 128B-swizzled K-major tile with a row pitch of
//     exactly 32 pixels, so "the 128 pixels shifted by (dy, dx)" is the same tile read from a start address
//     advanced by (dy*32 + dx) * 128 bytes (the swizzle XOR is taken from the absolute smem address, so the shifted
//     descriptor reads exactly what TMA wrote; base_offset stays 0), and the channel group is a 32-byte K advance.  Zero padding = TMA out-of-bounds fill.  A tile yields 4 rows x 30 pixels.
//   * B: per (group, tap) a 16x16 diagonal bf16 matrix, pre-swizzled at weight-pack time, resident in smem.
//   * D: 64 (gate: 128) TMEM columns, double buffered; epilogue warps tcgen05.ld, apply gelu(x1)*x2 for the gate
//     variant, and write 128-byte channel runs.
// Persistent CTAs, each bound to one 64-channel block; warp 0 TMA producer, warp 1 (+ warp 2 for the gate's
// second half) MMA issuers, 8 epilogue warps.
#include <algorithm>
#include <cstdlib>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int DT_TW = 32, DT_OW = 30, DT_OH = 4, DT_IH = 6, DT_CB = 64;
constexpr uint32_t DT_TILE_BYTES = DT_TW * DT_IH * DT_CB * 2;   // 24576
constexpr uint32_t DT_TILE_SLOT = DT_TILE_BYTES + 1024;         // + slack: shifted reads run 2 pixels past the tile
constexpr uint32_t DT_B_BYTES = 4 * 3 * 2048;                   // per 64-channel block and half: 4 groups x 3 atoms
constexpr int DT_EPI_WARPS = 8;
// warp 0: TMA producer; warps 1..NI: MMA issuers (2 per chunk(2) half, two 16-channel groups each); then 8 epilogue warps
template <int GATE> struct DtCfg {
  static constexpr int NH = GATE ? 2 : 1, NI = 2 * NH, EPI0 = 1 + NI, THREADS = (1 + NI + DT_EPI_WARPS) * 32;
  static constexpr int STAGES = GATE ? 3 : 6;      // smem tile ring (50 KB / 25 KB per stage)
  static constexpr int NACC = 4;                   // TMEM accumulator ring: 4 x 128 / 4 x 64 columns
  static constexpr int TMEM_COLS = GATE ? 512 : 256;
  static constexpr uint32_t SMEM = 1024 + NH * DT_B_BYTES + STAGES * NH * DT_TILE_SLOT + 256;
};
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16x(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// exact-GELU to ~3e-7: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), branch free, 2 MUFU (rcp, ex2)
__device__ __forceinline__ float gelu_as(float x) {
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

struct DtParams {
  int H, W, C, Cout, nimg;
  int tiles_x, tiles_y, cblocks;
  long tiles_per_cb;        // nimg * tiles_y * tiles_x
  int base_offset_mode;     // 0 (default, verified on B200): the 128B swizzle is a function of the absolute smem address, so a
                            // row-shifted start needs base_offset = 0; 1 sets base_offset = dx (kept as a bring-up switch)
  long ldo;
  float inv_tiles_x, inv_tiles_y;
};

template <int GATE>
__global__ void __launch_bounds__(DtCfg<GATE>::THREADS, 1)
k_dwconv_tc(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap map_out, const uint8_t* __restrict__ wtc,
            const DtParams p) {
  constexpr int NH = DtCfg<GATE>::NH, NI = DtCfg<GATE>::NI, EPI0 = DtCfg<GATE>::EPI0;
  constexpr int DT_STAGES = DtCfg<GATE>::STAGES, NACC = DtCfg<GATE>::NACC;
  constexpr uint32_t STAGE = NH * DT_TILE_SLOT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = sbase;                                  // resident weights: NH * 24 KB
  const uint32_t a_base = b_base + NH * DT_B_BYTES;               // DT_STAGES tile slots
  const uint32_t bar_base = a_base + DT_STAGES * STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (DT_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * DT_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * DT_STAGES + NACC + a); };
  const uint32_t wbar = bar_base + 8u * (2 * DT_STAGES + 2 * NACC);
  const uint32_t tmem_slot = bar_base + 8u * (2 * DT_STAGES + 2 * NACC + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // this CTA's channel block and its share of the pixel tiles
  const int cb = blockIdx.x % p.cblocks;
  const int cta_in_cb = blockIdx.x / p.cblocks;
  const int ctas_in_cb = (gridDim.x - cb + p.cblocks - 1) / p.cblocks;
  const int hp = p.C / 2;
  const int ch_valid = min(DT_CB, p.Cout - cb * DT_CB);           // valid channels of this block
  const int ngroups = (ch_valid + 15) / 16;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map);
    prefetch_tmap(&map_out);
    for (int s = 0; s < DT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull_bar(a), NI); mbar_init(tempty_bar(a), DT_EPI_WARPS); }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(DtCfg<GATE>::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  auto tile_xy = [&](long t64, int& img, int& y0, int& x0) {
    const int t = (int)t64;
    const int rowt = fast_div(t, p.tiles_x, p.inv_tiles_x);
    const int txi = t - rowt * p.tiles_x;
    img = fast_div(rowt, p.tiles_y, p.inv_tiles_y);
    const int tyi = rowt - img * p.tiles_y;
    x0 = txi * DT_OW; y0 = tyi * DT_OH;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // resident diagonal-weight blocks of this channel block (pre-swizzled by pack_dw_tc)
      mbar_expect_tx(wbar, NH * DT_B_BYTES);
      for (int h = 0; h < NH; ++h)
        bulk_copy_g2s(b_base + h * DT_B_BYTES, wtc + ((size_t)h * p.cblocks + cb) * DT_B_BYTES, DT_B_BYTES, wbar);
      uint32_t it = 0;
      for (long t = cta_in_cb; t < p.tiles_per_cb; t += ctas_in_cb, ++it) {
        const int s = it % DT_STAGES;
        mbar_wait_relaxed(empty_bar(s), ((it / DT_STAGES) & 1) ^ 1);
        int img, y0, x0;
        tile_xy(t, img, y0, x0);
        mbar_expect_tx(full_bar(s), NH * DT_TILE_BYTES);
        const uint32_t dst = a_base + s * STAGE;
        tma_load_4d(dst, &map, full_bar(s), cb * DT_CB, x0 - 1, y0 - 1, img);
        if (GATE) tma_load_4d(dst + DT_TILE_SLOT, &map, full_bar(s), hp + cb * DT_CB, x0 - 1, y0 - 1, img);
      }
    }
  } else if (warp <= NI) {
    // ===================== MMA issuers: (half h, group pair gp) each =====================
    if (lane == 0) {
      const int h = (warp - 1) >> 1, gp = (warp - 1) & 1;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
      mbar_wait(wbar, 0);
      const uint32_t b_lo0 = (((b_base + h * DT_B_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
      uint32_t it = 0;
      for (long t = cta_in_cb; t < p.tiles_per_cb; t += ctas_in_cb, ++it) {
        const int s = it % DT_STAGES;
        const uint32_t acc = it % NACC, aph = (it / NACC) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        mbar_wait(full_bar(s), (it / DT_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_lo0 = (((a_base + s * STAGE + h * DT_TILE_SLOT) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t d_base = tmem_base + acc * (NH * DT_CB) + h * DT_CB;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int g = gp * 2 + gg;
          if (g < ngroups) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int dy = tap / 3, dx = tap % 3;
              // A: tile shifted by (dy, dx) pixels (128 B each) + 32 B per channel group; B: diagonal block of (g, tap)
              const uint32_t a_lo = a_lo0 + (uint32_t)((dy * DT_TW + dx) * 8 + g * 2);
              const uint32_t b_lo = b_lo0 + (uint32_t)((g * 6144 + (tap >> 2) * 2048 + (tap & 3) * 32) >> 4);
              umma_bf16_lohi(d_base + g * 16, a_lo, b_lo, desc_hi, idesc, tap != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(tfull_bar(acc));     // accumulator ready; the smem slot is released by the epilogue after its store
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - EPI0;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int chalf = ew >> 2;               // channels [chalf*32, +32) of the 64-channel block
    const int r = quarter * 32 + lane;       // accumulator row = pixel of the tile
    const int ty_l = r / DT_TW, tx_l = r % DT_TW;
    // Output staging: the tile's own (already consumed) input slot, packed [4 rows][30 px][64 ch] with the 128B swizzle,
    // leaves through one TMA store box {64 ch, 30 px, 4 rows}: full-line writes, clipped at the image / channel edge.
    const int opix = ty_l * DT_OW + tx_l;
    const bool in_box = tx_l < DT_OW;
    uint8_t* sgen = smem_raw + (a_base - smem_u32(smem_raw));
    uint32_t it = 0;
    for (long t = cta_in_cb; t < p.tiles_per_cb; t += ctas_in_cb, ++it) {
      const uint32_t acc = it % NACC, aph = (it / NACC) & 1;
      const int s = it % DT_STAGES;
      mbar_wait_relaxed(tfull_bar(acc), aph);   // all MMAs of the tile retired: TMEM valid, smem slot no longer read
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * (NH * DT_CB) + ((uint32_t)(quarter * 32) << 16) + chalf * 32;
      const bool run0 = chalf * 32 < ch_valid, run1 = chalf * 32 + 16 < ch_valid;   // warp-uniform
      uint32_t a0[16], a1[16], b0[16], b1[16];
      if (run0) { tmem_ld16x(t_row, a0); if (GATE) tmem_ld16x(t_row + DT_CB, b0); }
      if (run1) { tmem_ld16x(t_row + 16, a1); if (GATE) tmem_ld16x(t_row + DT_CB + 16, b1); }
      uint8_t* srow = sgen + s * STAGE + opix * 128;
      auto emit = [&](uint32_t (&a)[16], uint32_t (&b)[16], int q) {
        tmem_wait16(a);
        if (GATE) tmem_wait16(b);
        if (in_box) {
#pragma unroll
          for (int v8 = 0; v8 < 2; ++v8) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float x1 = __uint_as_float(a[v8 * 8 + i]);
              f[i] = GATE ? gelu_fast(x1) * __uint_as_float(b[v8 * 8 + i]) : x1;
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
            const int chunk = chalf * 4 + q * 2 + v8;
            *reinterpret_cast<uint4*>(srow + ((chunk ^ (opix & 7)) << 4)) = o;
          }
        }
      };
      if (run0) emit(a0, b0, 0);
      if (run1) emit(a1, b1, 1);
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      asm volatile("bar.sync 1, %0;" ::"n"(DT_EPI_WARPS * 32) : "memory");   // whole tile staged
      if (ew == 0 && lane == 0) {
        int img, y0, x0;
        tile_xy(t, img, y0, x0);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(&map_out), "r"(a_base + s * STAGE), "r"(cb * DT_CB), "r"(x0), "r"(y0), "r"(img) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");       // smem read out -> slot free for the next load
        mbar_arrive(empty_bar(s));
      }
    }
    if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(DtCfg<GATE>::TMEM_COLS));
  }
}

// Diagonal weight blocks, pre-swizzled (128B swizzle, K-major): [half][cb][group 4][atom 3][16 rows][128 B]
__global__ void k_pack_dw_tc(const float* __restrict__ w9c, int C, int Cout, int gate, int cblocks, uint8_t* __restrict__ dst, long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;   // one bf16 element each
  if (idx >= total) return;
  long e = idx;
  const int el = (int)(e % 8); e /= 8;      // element in 16-byte chunk
  const int j = (int)(e % 8); e /= 8;       // logical chunk in the 128-byte row
  const int n = (int)(e % 16); e /= 16;     // row = output channel in group
  const int a = (int)(e % 3); e /= 3;       // atom (taps 4a .. 4a+3)
  const int g = (int)(e % 4); e /= 4;
  const int cb = (int)(e % cblocks);
  const int h = (int)(e / cblocks);
  const int k = j * 8 + el;
  const int tap = a * 4 + k / 16, cin = k % 16;
  const int chl = cb * DT_CB + g * 16 + n;                        // channel inside the half
  float v = 0.f;
  if (tap < 9 && cin == n && chl < Cout) v = w9c[(long)tap * C + (gate ? h * (C / 2) : 0) + chl];
  const long blk = ((((long)h * cblocks + cb) * 4 + g) * 3 + a) * 2048;
  const long off = blk + n * 128 + ((j ^ (n & 7)) << 4) + el * 2;
  *reinterpret_cast<bf16*>(dst + off) = __float2bfloat16_rn(v);
}

int g_dt_sms = 0;
int g_dt_base_offset_mode = 0;

}  // namespace

size_t dwconv_tc_weight_bytes(int C, int gate) {
  const int Cout = gate ? C / 2 : C;
  return (size_t)(gate ? 2 : 1) * cdiv(Cout, DT_CB) * DT_B_BYTES;
}

int pack_dw_tc(const float* w9c, int C, int gate, void* dst, cudaStream_t s) {
  const int Cout = gate ? C / 2 : C;
  const int cblocks = cdiv(Cout, DT_CB);
  const long total = (long)dwconv_tc_weight_bytes(C, gate) / 2;
  k_pack_dw_tc<<<cdiv(total, 256), 256, 0, s>>>(w9c, C, Cout, gate, cblocks, reinterpret_cast<uint8_t*>(dst), total);
  KD_LAUNCH_CHECK();
  return 0;
}

// returns -1 when the shape is not eligible
int dwconv3x3_tc(const bf16* x, long ldx, bf16* out, long ldo, const void* wtc, int nimg, int H, int W, int C, int gate,
                 cudaStream_t s) {
  if (wtc == nullptr || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || ldx % 8 || ldo % 8 ||
      C % (gate ? 16 : 8) || (reinterpret_cast<uintptr_t>(wtc) & 15))
    return -1;
  const uint32_t smem0 = DtCfg<0>::SMEM, smem1 = DtCfg<1>::SMEM;
  if (g_dt_sms == 0) {
    int dev = 0;
    KD_CUDA(cudaGetDevice(&dev));
    KD_CUDA(cudaDeviceGetAttribute(&g_dt_sms, cudaDevAttrMultiProcessorCount, dev));
    g_dt_sms = sm_limit(g_dt_sms);
    KD_CUDA(cudaFuncSetAttribute(k_dwconv_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem0));
    KD_CUDA(cudaFuncSetAttribute(k_dwconv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    const char* e = getenv("KDLAE_DWTC_BASE_OFFSET");
    if (e) g_dt_base_offset_mode = atoi(e);
  }
  DtParams p;
  p.H = H; p.W = W; p.C = C; p.Cout = gate ? C / 2 : C; p.nimg = nimg;
  p.tiles_x = cdiv(W, DT_OW); p.tiles_y = cdiv(H, DT_OH); p.cblocks = cdiv(p.Cout, DT_CB);
  p.tiles_per_cb = (long)nimg * p.tiles_x * p.tiles_y;
  p.base_offset_mode = g_dt_base_offset_mode;
  p.ldo = ldo;
  p.inv_tiles_x = 1.0f / (float)p.tiles_x;
  p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  if (p.tiles_per_cb >= (1L << 24)) return -1;   // fast_div range
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
  const cuuint64_t str[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H};
  const cuuint32_t box[4] = {DT_CB, DT_TW, DT_IH, 1};
  KD_TRY(make_map(&map, x, 4, dims, str, box));
  CUtensorMap map_out;
  {
    const cuuint64_t od[4] = {(cuuint64_t)p.Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
    const cuuint64_t os[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)ldo * 2 * W, (cuuint64_t)ldo * 2 * W * H};
    const cuuint32_t ob[4] = {DT_CB, DT_OW, DT_OH, 1};
    KD_TRY(make_map(&map_out, out, 4, od, os, ob));
  }
  const long want = (long)p.cblocks * p.tiles_per_cb;
  int grid = (int)std::min<long>(want, (long)g_dt_sms);
  if (grid < p.cblocks) grid = p.cblocks;   // every channel block needs at least one CTA
  ProfScope prof(PC_DWCONV, s, 18.0 * nimg * H * W * C, (double)nimg * H * W * (C + p.Cout) * 2.0 + 36.0 * C);
  if (gate) k_dwconv_tc<1><<<grid, DtCfg<1>::THREADS, smem1, s>>>(map, map_out, reinterpret_cast<const uint8_t*>(wtc), p);
  else k_dwconv_tc<0><<<grid, DtCfg<0>::THREADS, smem0, s>>>(map, map_out, reinterpret_cast<const uint8_t*>(wtc), p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
