// MDTA Gram / norm reduction on tcgen05 (KDLAE_model.py:134-137), bf16 path.
//
// For one (image, head, pixel split) the CTA accumulates over pixels p (the GEMM K axis):
//     G[i][j] = sum_p q[p][i] k[p][j],   Nq = q^T q,   Nk = k^T k   (only the diagonals of Nq, Nk are used)
// q and k live in the depthwise-conv output [pixel][3C] (channel-contiguous), i.e. both operands are
// MN-major: a TMA box {64 channels, 64 pixels} lands as a 128B-swizzled MN-major UMMA atom stack with
// no transpose.  M = 128 rows (channels of the head; rows >= ch are ignored), N = ch, K = 16 pixels per
// tcgen05.mma, fp32 accumulation in TMEM across the whole split.  The kernel is a pure stream over
// q,k (2C bf16 per pixel): HBM-bound.  Output: per-split partials [ch*ch + 2ch] reduced by mdta_fold.
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int GR_PIX = 64;                       // pixels per stage (4 MMA k-steps)
constexpr uint32_t GR_ATOM = GR_PIX * 128;       // 64 pixels x 64 channels bf16 = 8 KB
// NATOMS = 64-channel atoms per operand (1: head width <= 64, the KDLAE case; 2: up to 128).
// NATOMS = 1: 6 stages of 16 KB and 256 TMEM columns, so two CTAs share an SM (one CTA's ramp-up / epilogue hides under the
// other's stream); NATOMS = 2: 4 stages of 32 KB, 512 columns, one CTA per SM.
template <int NATOMS> struct GrCfg {
  static constexpr int STAGES = NATOMS == 1 ? 6 : 4;
  static constexpr uint32_t STAGE_BYTES = 2 * NATOMS * GR_ATOM;    // q atoms, then k atoms
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + 1024 + 128;
  static constexpr int ACC = NATOMS * 64;                          // TMEM columns per accumulator (G, Nq, Nk)
  static constexpr int TMEM_COLS = NATOMS == 1 ? 256 : 512;
};

// MN-major SWIZZLE_128B descriptor: 64-element (128 B) MN chunks, 8 K-rows per 1024 B atom,
// SBO = stride between 8-row K groups, LBO = stride between 64-wide MN chunks.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {   // D=f32, A=B=bf16, A and B MN-major, M=128
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int NATOMS>
__global__ void __launch_bounds__(128, NATOMS == 1 ? 2 : 1)
k_mdta_gram_tc(const __grid_constant__ CUtensorMap map, int HW, int C, int heads, int splits, int per, float* __restrict__ part) {
  constexpr int GR_STAGES = GrCfg<NATOMS>::STAGES, ACC = GrCfg<NATOMS>::ACC;
  constexpr uint32_t GR_STAGE_BYTES = GrCfg<NATOMS>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = sbase + GR_STAGES * GR_STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GR_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * GR_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * GR_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int ch = C / heads;
  const int split = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int natoms = NATOMS;
  const int p_begin = split * per;
  const int p_end = min(HW, p_begin + per);
  const int nchunks = (p_end > p_begin) ? (p_end - p_begin + GR_PIX - 1) / GR_PIX : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map);
    for (int s = 0; s < GR_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(GrCfg<NATOMS>::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ---- TMA producer: q and k channel atoms of 64 pixels per stage (warp-uniform loop, elected issuer: see elect_one()) ----
    for (int i = 0; i < nchunks; ++i) {
      const int s = i % GR_STAGES;
      mbar_wait(empty_bar(s), ((i / GR_STAGES) & 1) ^ 1);
      const uint32_t dst = sbase + s * GR_STAGE_BYTES;
      const int p = p_begin + i * GR_PIX;
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), 2 * natoms * GR_ATOM);
        tma_load_4d(dst, &map, full_bar(s), 0, head, p, img);
        tma_load_4d(dst + natoms * GR_ATOM, &map, full_bar(s), 0, heads + head, p, img);
        if (natoms == 2) {
          tma_load_4d(dst + GR_ATOM, &map, full_bar(s), 64, head, p, img);
          tma_load_4d(dst + 3 * GR_ATOM, &map, full_bar(s), 64, heads + head, p, img);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer (warp-uniform loop, elected issuer) ----
    const uint32_t idesc = make_idesc_mn(ch);
    const uint32_t lbo = (natoms == 2) ? GR_ATOM : 0u;
    for (int i = 0; i < nchunks; ++i) {
      const int s = i % GR_STAGES;
      mbar_wait(full_bar(s), (i / GR_STAGES) & 1);
      tc_fence_after();
      const uint32_t qa = sbase + s * GR_STAGE_BYTES, ka = qa + natoms * GR_ATOM;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < GR_PIX / 16; ++k) {
          const uint64_t dq = make_desc_mn(qa + k * 2048, lbo), dk = make_desc_mn(ka + k * 2048, lbo);
          const uint32_t accum = (i | k) != 0 ? 1u : 0u;
          umma_bf16(tmem_base + 0, dq, dk, idesc, accum);     // G  = q^T k
          umma_bf16(tmem_base + ACC, dq, dq, idesc, accum);       // Nq = q^T q
          umma_bf16(tmem_base + 2 * ACC, dk, dk, idesc, accum);   // Nk = k^T k
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  }
  __syncwarp();

  // ---- epilogue: all 4 warps; thread = accumulator row i (q channel) ----
  float* dst = part + (((long)img * heads + head) * splits + split) * (long)(ch * ch + 2 * ch);
  const int i = warp * 32 + lane;
  if (nchunks > 0) {
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    float nq = 0.f, nk = 0.f;
    for (int c0 = 0; c0 < ch; c0 += 16) {
      uint32_t g[16], a[16], b[16];
      tmem_ld16(t_row + c0, g);
      tmem_ld16(t_row + ACC + c0, a);
      tmem_ld16(t_row + 2 * ACC + c0, b);
      if (i < ch) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j < ch) dst[(long)i * ch + c0 + j] = __uint_as_float(g[j]);
          if (c0 + j == i) { nq = __uint_as_float(a[j]); nk = __uint_as_float(b[j]); }
        }
      }
    }
    if (i < ch) { dst[ch * ch + i] = nq; dst[ch * ch + ch + i] = nk; }
  } else {
    for (int e = threadIdx.x; e < ch * ch + 2 * ch; e += 128) dst[e] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(GrCfg<NATOMS>::TMEM_COLS));
  }
}

}  // namespace

// returns -1 when the shape is not eligible (caller falls back to the CUDA-core kernel)
int mdta_gram_tc(const bf16* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s) {
  const int ch = C / heads;
  if (ch % 16 || ch > 128 || ch < 16 || ld % 8 || (reinterpret_cast<uintptr_t>(qk) & 15) || C % 8) return -1;
  static DeviceOnce once;
  bool first; int dev;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_mdta_gram_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GrCfg<1>::SMEM));
    KD_CUDA(cudaFuncSetAttribute(k_mdta_gram_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GrCfg<2>::SMEM));
    device_mark(once, dev);
  }
  CUtensorMap map;
  // {channel within a head, head slot (q heads then k heads), pixel, image}: the 64-channel box of a 48-channel head is clipped at
  // the head boundary (zero fill, nothing fetched) instead of pulling the neighbouring head's / v's channels through HBM
  const cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)(2 * heads), (cuuint64_t)HW, (cuuint64_t)nimg};
  const cuuint64_t str[3] = {(cuuint64_t)ch * 2, (cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * HW};
  const cuuint32_t box[4] = {64, 1, GR_PIX, 1};
  KD_TRY(make_map(&map, qk, 4, dims, str, box));
  int per = (HW + splits - 1) / splits;
  per = (per + GR_PIX - 1) / GR_PIX * GR_PIX;   // split boundaries on 64-pixel chunks; the tail is TMA zero fill
  ProfScope prof(PC_MDTA_GRAM, s, 2.0 * nimg * HW * C * ch + 4.0 * nimg * HW * C,
                 (double)nimg * HW * 2 * C * 2.0 + 4.0 * nimg * heads * splits * (ch * ch + 2 * ch));
  if (ch <= 64) k_mdta_gram_tc<1><<<dim3(splits, heads, nimg), 128, GrCfg<1>::SMEM, s>>>(map, HW, C, heads, splits, per, part);
  else k_mdta_gram_tc<2><<<dim3(splits, heads, nimg), 128, GrCfg<2>::SMEM, s>>>(map, HW, C, heads, splits, per, part);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
