// TF32 tensor-core GEMM for the dense convolutions of the training step (fp32 storage, tcgen05 kind::tf32, fp32 accumulation):
//   out[p][n] = sum_k A[p][k] * W[n][k] (+ res[p][n]),   A [P][K] fp32 (row stride ld0), W [N][K] fp32 (per-image groups optional)
// and, for the dense 3x3 convs, the same sum over nine taps: the A tile of a tap is the 16 x 8 pixel patch shifted by the tap
// (TMA box coordinates; zero padding = out-of-bounds fill; dilation = a longer shift), W [N][tap][K].
// Opt-in (training.set_matmul_precision("tf32") / KDLAE_TRAIN_TF32=1): PyTorch itself runs the reference's nn.Conv2d layers in
// TF32 on an Ampere-or-newer GPU by default (torch.backends.cudnn.allow_tf32), the fp32 CUDA-core path stays the default here
// because the gradient-parity tests hold it to 1e-5 of float64 autograd.
// Operands arrive by TMA as 128-byte-swizzled K-major tiles of 32 floats per row (one swizzle atom), four K = 8 MMAs per tile;
// persistent CTAs, warp 0 producer, warp 1 MMA issuer + TMEM allocator, warps 2-5 epilogue (one TMEM lane quarter each),
// two 256-column accumulators so the epilogue of a tile overlaps the MMAs of the next.  K and N tails are TMA zero fill.
// The wgrad kernel below runs over MN-major operands (the reduction is over pixels): 32-bit MN-major tiles need the 128-byte swizzle
// with 32-byte atoms (TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, descriptor layout type 1, K groups of 4 rows) - with the ordinary
// 128B swizzle that gram_tc.cu uses for bf16 a kind::tf32 MMA returns zeros.
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int TF_BM = 128, TF_BK = 32, TF_NC_MAX = 256, TF_STAGES = 4;
constexpr uint32_t TF_A_BYTES = TF_BM * 128, TF_B_BYTES = TF_NC_MAX * 128, TF_STAGE = TF_A_BYTES + TF_B_BYTES;
constexpr uint32_t TF_SMEM = TF_STAGES * TF_STAGE + 1024 + 256;
constexpr int TF_THREADS = 6 * 32;

struct TfParams {
  int K, N, nc, n_chunks, kchunks;
  long rows_per_group;
  int tiles_per_group;
  long items;
  float inv_n_chunks, inv_tiles_per_group;
  // dense 3x3 conv (dilation dil, zero padding = TMA out-of-bounds fill): tiles are 16 x 8 pixel patches, 9 taps x kchunks K steps
  int spatial, taps, dil, H, W, tiles_x, tiles_y;
  float inv_tiles_x, inv_tiles_y;
  const float* res; long res_ld;
  float* out; long out_ld; int out_coff;
};

__device__ __forceinline__ void umma_tf32_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accum)
      : "memory");
}

__global__ void __launch_bounds__(TF_THREADS, 1)
k_gemm_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const TfParams p) {
  constexpr int TW = 16, TH = 8;          // spatial tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TF_STAGES * TF_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TF_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TF_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TF_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TF_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a); prefetch_tmap(&map_w);
    for (int s = 0; s < TF_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
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

  // linear mode: g = weight group, r0 = first row in the group.  spatial mode: g = image, r0 = (ty0 << 16) | tx0
  auto coord = [&](long item64, int& g, int& r0, int& nchunk) {
    const int item = (int)item64;
    const int mt = fast_div(item, p.n_chunks, p.inv_n_chunks);
    nchunk = item - mt * p.n_chunks;
    if (p.spatial) {
      const int rowt = fast_div(mt, p.tiles_x, p.inv_tiles_x);
      const int txi = mt - rowt * p.tiles_x;
      g = fast_div(rowt, p.tiles_y, p.inv_tiles_y);
      r0 = (((rowt - g * p.tiles_y) * TH) << 16) | (txi * TW);
    } else {
      g = fast_div(mt, p.tiles_per_group, p.inv_tiles_per_group);
      r0 = (mt - g * p.tiles_per_group) * TF_BM;
    }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0; uint32_t ph = 0;
    const uint32_t stage_tx = TF_A_BYTES + (uint32_t)p.nc * 128;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
      int g, r0, nchunk;
      coord(item, g, r0, nchunk);
      for (int tap = 0; tap < p.taps; ++tap) {
        const int dx = (tap % 3 - 1) * p.dil, dy = (tap / 3 - 1) * p.dil;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait_relaxed(empty_bar(s), ph ^ 1);
          if (elect_one()) {
            const uint32_t a_dst = smem_base + s * TF_STAGE;
            mbar_expect_tx(full_bar(s), stage_tx);
            if (p.spatial) {
              tma_load_4d(a_dst, &map_a, full_bar(s), kc * TF_BK, (r0 & 0xFFFF) + dx, (r0 >> 16) + dy, g);
              tma_load_3d(a_dst + TF_A_BYTES, &map_w, full_bar(s), kc * TF_BK, nchunk * p.nc, tap);
            } else {
              tma_load_3d(a_dst, &map_a, full_bar(s), kc * TF_BK, r0, g);
              tma_load_3d(a_dst + TF_A_BYTES, &map_w, full_bar(s), kc * TF_BK, nchunk * p.nc, g);
            }
          }
          __syncwarp();
          if (++s == TF_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    // instruction descriptor for kind::tf32: D = f32, A = B = tf32, both K-major, M = 128, N = nc
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, version 1, SWIZZLE_128B
    const uint32_t lo_tag = 1u << 16;
    int s = 0; uint32_t ph = 0, it = 0;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TF_NC_MAX;
      for (int kt = 0; kt < p.taps * p.kchunks; ++kt) {
        const int kc = p.taps == 1 ? kt : kt % p.kchunks;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_lo = (((smem_base + s * TF_STAGE) & 0x3FFFF) >> 4) | lo_tag;
        const uint32_t b_lo = (((smem_base + s * TF_STAGE + TF_A_BYTES) & 0x3FFFF) >> 4) | lo_tag;
        const int rem = p.K - kc * TF_BK;
        const int ksn = rem >= TF_BK ? 4 : (rem + 7) >> 3;      // K = 8 steps that hold real channels (the rest is zero fill)
        const uint32_t first = kt != 0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < ksn) umma_tf32_lohi(d_tmem, a_lo + k * 2, b_lo + k * 2, desc_hi, idesc, k != 0 ? 1u : first);
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == TF_STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(tfull_bar(acc));
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warp w may access TMEM lanes [32 (w % 4), 32 (w % 4) + 32) =====================
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    uint32_t it = 0;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      int g, r0, nchunk;
      coord(item, g, r0, nchunk);
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      bool valid;
      long prow;
      if (p.spatial) {
        const int x = (r0 & 0xFFFF) + (r % TW), y = (r0 >> 16) + (r / TW);
        valid = x < p.W && y < p.H;
        prow = ((long)g * p.H + y) * p.W + x;
      } else {
        const long rr = (long)r0 + r;
        valid = rr < p.rows_per_group;
        prow = (long)g * p.rows_per_group + rr;
      }
      const uint32_t t_row = tmem_base + acc * TF_NC_MAX + ((uint32_t)(quarter * 32) << 16);
      const int nbase = nchunk * p.nc;
      mbar_wait_relaxed(tfull_bar(acc), aph);
      tc_fence_after();
      for (int c0 = 0; c0 < p.nc; c0 += 16) {
        uint32_t v[16];
        tmem_ld16_issue(t_row + c0, v);
        tmem_ld16_wait(v);
        if (valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int n = nbase + c0 + q * 4;
            if (n < p.N) {                                  // N % 4 == 0: a float4 is entirely inside or outside
              float4 o = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                                     __uint_as_float(v[q * 4 + 3]));
              if (p.res) {
                const float4 rv = __ldg(reinterpret_cast<const float4*>(p.res + prow * p.res_ld + n));
                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
              }
              *reinterpret_cast<float4*>(p.out + prow * p.out_ld + p.out_coff + n) = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

int make_map_f32(CUtensorMap* m, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                 CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int rank = 3) {
  EncodeTiledFn fn = get_encode_fn();
  KD_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KD_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------------------
// wgrad of a 1x1 conv on the same tensor-core path:  part[split][n][k] = sum_{p in split} A[p][n] * B[p][k]
// (A = dY [P][N], B = X [P][K], both row-major in the pixel).  The reduction runs over pixels, so both operands are MN-major:
// a TMA box {32 floats, 32 pixels} lands as rows of 128 bytes whose 32-byte pieces are permuted by (row mod 4) - the MN-major
// canonical layout for 32-bit operands; M = 128 rows of n per CTA = four such chunks (LBO = chunk stride), N = up to 256 columns
// of k = eight, K = 8 pixels per tcgen05.mma = two 4-row groups (SBO = 512 B), fp32 accumulation in TMEM over the CTA's whole
// pixel split; the partials are reduced in fixed order by the caller.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int WT_PIX = 32, WT_STAGES = 4;
constexpr uint32_t WT_CHUNK = WT_PIX * 128;                       // 32 pixels x 32 floats
constexpr uint32_t WT_STAGE = (4 + 8) * WT_CHUNK;                 // A chunks, then up to 8 B chunks
constexpr uint32_t WT_SMEM = WT_STAGES * WT_STAGE + 1024 + 128;

struct WtParams {
  int N, K, kc, nbch, k_tiles;
  long P;
  int per;               // pixels per split, a multiple of WT_PIX
  float* part;
};

__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // stride between 32-float MN chunks
  d |= (uint64_t)(512 >> 4) << 32;                        // stride between 4-pixel K groups
  d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
  d |= (uint64_t)1 << 61;                                 // layout type 1: SWIZZLE_128B with 32-byte atoms
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
k_wgrad_tf32(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const WtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = sbase + WT_STAGES * WT_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WT_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * WT_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * WT_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x / p.k_tiles, kt = blockIdx.x % p.k_tiles, split = blockIdx.y;
  const long p_begin = (long)split * p.per, p_end = min(p.P, p_begin + p.per);
  const int nsteps = p_end > p_begin ? (int)((p_end - p_begin + WT_PIX - 1) / WT_PIX) : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a); prefetch_tmap(&map_b);
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    for (int i = 0; i < nsteps; ++i) {
      const int s = i % WT_STAGES;
      mbar_wait_relaxed(empty_bar(s), ((i / WT_STAGES) & 1) ^ 1);
      const uint32_t dst = sbase + s * WT_STAGE;
      const int px = (int)(p_begin + (long)i * WT_PIX);
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), (uint32_t)(4 + p.nbch) * WT_CHUNK);
        for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * WT_CHUNK, &map_a, full_bar(s), nt * 128 + c * 32, px, 0);
        for (int c = 0; c < p.nbch; ++c) tma_load_3d(dst + (4 + c) * WT_CHUNK, &map_b, full_bar(s), kt * 256 + c * 32, px, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // D = f32, A = B = tf32, both MN-major, M = 128, N = kc
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.kc >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    for (int i = 0; i < nsteps; ++i) {
      const int s = i % WT_STAGES;
      mbar_wait(full_bar(s), (i / WT_STAGES) & 1);
      tc_fence_after();
      const uint32_t aa = sbase + s * WT_STAGE, ba = aa + 4 * WT_CHUNK;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < WT_PIX / 8; ++k)
          umma_tf32(tmem_base, make_desc_mn32(aa + k * 1024, WT_CHUNK), make_desc_mn32(ba + k * 1024, WT_CHUNK), idesc, (i | k) != 0 ? 1u : 0u);
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  }
  __syncwarp();

  // epilogue: all four warps, thread = accumulator row n
  const int n = nt * 128 + warp * 32 + lane;
  float* dst = p.part + ((long)split * p.N + n) * p.K + kt * 256;
  if (nsteps > 0) {
    mbar_wait_relaxed(done_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.kc; c0 += 16) {
      uint32_t v[16];
      tmem_ld16_issue(t_row + c0, v);
      tmem_ld16_wait(v);
      if (n < p.N) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (kt * 256 + c0 + j < p.K) dst[c0 + j] = __uint_as_float(v[j]);
      }
    }
  } else if (n < p.N) {
    for (int j = 0; j < p.kc; ++j)
      if (kt * 256 + j < p.K) dst[j] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------------
// MDTA Gram reductions of the training step (KDLAE_model.py:134-137 and their backward) on the same MN-major path:
// for every (image, head, pixel split)   G = q^T k,  Nq = q^T q,  Nk = k^T k   (only the diagonals of Nq, Nk are used)
// with q at channel head * ch and k at channel C + head * ch of a [pixel][ld] fp32 tensor - the layout, grid and partial-sum
// format of k_mdta_gram<float> (glue.cu) / gram_tc.cu, so the fixed-order reduction that follows is unchanged.
// ---------------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int GT_STAGES = 4;
constexpr uint32_t GT_STAGE = 8 * WT_CHUNK;              // q: 4 chunks of 32 channels (M = 128 rows), k: 4 chunks

__global__ void __launch_bounds__(128, 1)
k_gram_tf32(const __grid_constant__ CUtensorMap map, int HW, int C, int heads, int splits, int per, float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = sbase + GT_STAGES * GT_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GT_STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * GT_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * GT_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = C / heads, kc = (ch + 15) / 16 * 16;
  const int split = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int p_begin = split * per, p_end = min(HW, p_begin + per);
  const int nsteps = p_end > p_begin ? (p_end - p_begin + WT_PIX - 1) / WT_PIX : 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map);
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
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

  if (warp == 0) {
    for (int i = 0; i < nsteps; ++i) {
      const int s = i % GT_STAGES;
      mbar_wait_relaxed(empty_bar(s), ((i / GT_STAGES) & 1) ^ 1);
      const uint32_t dst = sbase + s * GT_STAGE;
      const int px = p_begin + i * WT_PIX;
      if (elect_one()) {
        mbar_expect_tx(full_bar(s), GT_STAGE);
        for (int c = 0; c < 4; ++c) tma_load_3d(dst + c * WT_CHUNK, &map, full_bar(s), head * ch + c * 32, px, img);
        for (int c = 0; c < 4; ++c) tma_load_3d(dst + (4 + c) * WT_CHUNK, &map, full_bar(s), C + head * ch + c * 32, px, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(kc >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    for (int i = 0; i < nsteps; ++i) {
      const int s = i % GT_STAGES;
      mbar_wait(full_bar(s), (i / GT_STAGES) & 1);
      tc_fence_after();
      const uint32_t qa = sbase + s * GT_STAGE, ka = qa + 4 * WT_CHUNK;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < WT_PIX / 8; ++k) {
          const uint64_t dq = make_desc_mn32(qa + k * 1024, WT_CHUNK), dk = make_desc_mn32(ka + k * 1024, WT_CHUNK);
          const uint32_t accum = (i | k) != 0 ? 1u : 0u;
          umma_tf32(tmem_base + 0, dq, dk, idesc, accum);       // G  = q^T k
          umma_tf32(tmem_base + 128, dq, dq, idesc, accum);     // Nq = q^T q
          umma_tf32(tmem_base + 256, dk, dk, idesc, accum);     // Nk = k^T k
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  }
  __syncwarp();

  float* dst = part + (((long)img * heads + head) * splits + split) * (long)(ch * ch + 2 * ch);
  const int i = warp * 32 + lane;
  if (nsteps > 0) {
    mbar_wait_relaxed(done_bar, 0);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    float nq = 0.f, nk = 0.f;
    for (int c0 = 0; c0 < kc; c0 += 16) {
      uint32_t g[16], a[16], b[16];
      tmem_ld16_issue(t_row + c0, g); tmem_ld16_wait(g);
      tmem_ld16_issue(t_row + 128 + c0, a); tmem_ld16_wait(a);
      tmem_ld16_issue(t_row + 256 + c0, b); tmem_ld16_wait(b);
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace

bool gram_tf32_eligible(const float* qk, long ld, int C, int heads) {
  const int ch = heads > 0 ? C / heads : 0;
  return heads > 0 && C % heads == 0 && ch % 4 == 0 && ch >= 8 && ch <= 96 && ld % 4 == 0 && !(reinterpret_cast<uintptr_t>(qk) & 15);
}

// same contract as mdta_gram<float> (glue.cu): part[((img * heads + head) * splits + split)][ch*ch + 2ch]
int gram_tf32(const float* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s) {
  KD_CHECK(gram_tf32_eligible(qk, ld, C, heads), "gram_tf32: shape not eligible");
  static DeviceOnce once;
  bool first; int dev;
  KD_TRY(device_first_use(once, &first, &dev));
  const uint32_t smem = GT_STAGES * GT_STAGE + 1024 + 128;
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_gram_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    device_mark(once, dev);
  }
  CUtensorMap map;
  const cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)HW, (cuuint64_t)nimg};
  const cuuint64_t str[2] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * HW};
  const cuuint32_t box[3] = {32, WT_PIX, 1};
  KD_TRY(make_map_f32(&map, qk, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  int per = (HW + splits - 1) / splits;
  per = (per + WT_PIX - 1) / WT_PIX * WT_PIX;           // split boundaries on 32-pixel steps; the tail is TMA zero fill
  const int ch = C / heads;
  ProfScope prof(PC_MDTA_GRAM, s, 2.0 * nimg * HW * C * ch + 4.0 * nimg * HW * C,
                 (double)nimg * HW * 2 * C * 4.0 + 4.0 * nimg * heads * splits * (ch * ch + 2 * ch));
  k_gram_tf32<<<dim3(splits, heads, nimg), 128, smem, s>>>(map, HW, C, heads, splits, per, part);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

bool wgrad_tf32_eligible(const float* A, long lda, int N, const float* B, long ldb, int K) {
  return N % 4 == 0 && K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && !(reinterpret_cast<uintptr_t>(A) & 15) &&
         !(reinterpret_cast<uintptr_t>(B) & 15);
}

// part[split][n][k] over pixel ranges that are multiples of 32 pixels; *splits_out = the number of splits written (<= splits)
int wgrad_tf32(const float* A, long lda, int N, const float* B, long ldb, int K, long P, float* part, int splits, int* splits_out,
               cudaStream_t s) {
  KD_CHECK(wgrad_tf32_eligible(A, lda, N, B, ldb, K), "wgrad_tf32: shape not eligible");
  static DeviceOnce once;
  bool first; int dev;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_wgrad_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
    device_mark(once, dev);
  }
  WtParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.K = K; p.P = P; p.part = part;
  p.k_tiles = (K + 255) / 256;
  p.kc = p.k_tiles == 1 ? (K + 15) / 16 * 16 : 256;
  p.nbch = (p.kc + 31) / 32;
  long per = (P + splits - 1) / splits;
  per = (per + WT_PIX - 1) / WT_PIX * WT_PIX;
  p.per = (int)per;
  const int used = (int)((P + per - 1) / per);
  *splits_out = used;
  CUtensorMap ma, mb;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)P, 1};
    const cuuint64_t str[2] = {(cuuint64_t)lda * 4, (cuuint64_t)lda * 4 * P};
    const cuuint32_t box[3] = {32, WT_PIX, 1};
    KD_TRY(make_map_f32(&ma, A, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)P, 1};
    const cuuint64_t str[2] = {(cuuint64_t)ldb * 4, (cuuint64_t)ldb * 4 * P};
    const cuuint32_t box[3] = {32, WT_PIX, 1};
    KD_TRY(make_map_f32(&mb, B, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
  }
  const int n_tiles = (N + 127) / 128;
  ProfScope prof(PC_GEMM_TC, s, 2.0 * P * N * K, 4.0 * (double)P * (N + K));
  k_wgrad_tf32<<<dim3(n_tiles * p.k_tiles, used), 128, WT_SMEM, s>>>(ma, mb, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

// 1x1, one source, identity addressing, no row scale / bias / ReLU / statistics: what the training step's GEMMs use
bool gemm_tf32_eligible(const ConvOp& op) {
  const Epilogue& e = op.epi;
  const bool k1 = op.kh == 1 && op.kw == 1;
  const bool k3 = op.kh == 3 && op.kw == 3 && op.groups == 1 && op.w_tap_ld % 4 == 0 && op.W < 65536 && op.H < 32768;
  return op.kd == 1 && (k1 || k3) && op.c1 == 0 && op.xpack_cin == 0 && e.mode == OUT_IDENTITY && e.row_scale == nullptr &&
         e.row_mu == nullptr && e.col_bias == nullptr && !e.relu && e.stat_rstd == nullptr && e.planar_out == nullptr &&
         op.c0 % 4 == 0 && op.ld0 % 4 == 0 && op.w_ld % 4 == 0 && op.w_group_stride % 4 == 0 && e.N % 4 == 0 && e.out_ld % 4 == 0 &&
         e.out_coff % 4 == 0 && (e.res == nullptr || e.res_ld % 4 == 0) && !(reinterpret_cast<uintptr_t>(op.a0) & 15) &&
         !(reinterpret_cast<uintptr_t>(op.w) & 15) && !(reinterpret_cast<uintptr_t>(e.out) & 15) &&
         !(reinterpret_cast<uintptr_t>(e.res) & 15);
}

int gemm_tf32(const ConvOp& op, cudaStream_t s) {
  KD_CHECK(gemm_tf32_eligible(op), "gemm_tf32: shape not eligible");
  static DeviceOnce once;
  bool first; int dev, sms;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_gemm_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));
    device_mark(once, dev);
  }
  KD_TRY(device_sms(&sms));
  const Epilogue& e = op.epi;
  const long rows = (long)op.nimg * op.H * op.W;
  KD_CHECK(op.groups >= 1 && rows % op.groups == 0, "gemm_tf32: rows %ld not divisible by groups %d", rows, op.groups);
  TfParams p;
  memset(&p, 0, sizeof(p));
  p.K = op.c0; p.N = e.N;
  const int n16 = (e.N + 15) / 16 * 16;
  if (n16 <= TF_NC_MAX) { p.nc = n16; p.n_chunks = 1; }
  else { p.n_chunks = (e.N + TF_NC_MAX - 1) / TF_NC_MAX; p.nc = ((e.N + p.n_chunks - 1) / p.n_chunks + 15) / 16 * 16; }
  p.kchunks = (op.c0 + TF_BK - 1) / TF_BK;
  p.rows_per_group = rows / op.groups;
  p.tiles_per_group = (int)cdiv(p.rows_per_group, TF_BM);
  p.items = (long)p.tiles_per_group * op.groups * p.n_chunks;
  p.spatial = op.kh == 3 ? 1 : 0;
  p.taps = op.kh * op.kw; p.dil = op.dil; p.H = op.H; p.W = op.W;
  if (p.spatial) {
    p.tiles_x = (int)cdiv(op.W, 16); p.tiles_y = (int)cdiv(op.H, 8);
    p.items = (long)op.nimg * p.tiles_x * p.tiles_y * p.n_chunks;
    p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  }
  KD_CHECK(p.items < (1L << 24), "gemm_tf32: too many tiles (%ld)", p.items);
  p.inv_n_chunks = 1.0f / (float)p.n_chunks; p.inv_tiles_per_group = 1.0f / (float)p.tiles_per_group;
  p.res = reinterpret_cast<const float*>(e.res); p.res_ld = e.res_ld;
  p.out = reinterpret_cast<float*>(e.out); p.out_ld = e.out_ld; p.out_coff = (int)e.out_coff;
  CUtensorMap ma, mw;
  if (p.spatial) {
    const cuuint64_t dims[4] = {(cuuint64_t)op.c0, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.nimg};
    const cuuint64_t str[3] = {(cuuint64_t)op.ld0 * 4, (cuuint64_t)op.ld0 * 4 * op.W, (cuuint64_t)op.ld0 * 4 * op.W * op.H};
    const cuuint32_t box[4] = {TF_BK, 16, 8, 1};
    KD_TRY(make_map_f32(&ma, op.a0, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, 4));
    // weights [n][tap][c] seen as {c, n, tap}: channels past C are zero fill, so a K chunk never runs into the next tap
    const cuuint64_t wdims[3] = {(cuuint64_t)op.c0, (cuuint64_t)e.N, (cuuint64_t)p.taps};
    const cuuint64_t wstr[2] = {(cuuint64_t)op.w_ld * 4, (cuuint64_t)op.w_tap_ld * 4};
    const cuuint32_t wbox[3] = {TF_BK, (cuuint32_t)p.nc, 1};
    KD_TRY(make_map_f32(&mw, op.w, wdims, wstr, wbox));
  } else {
    const cuuint64_t dims[3] = {(cuuint64_t)op.c0, (cuuint64_t)p.rows_per_group, (cuuint64_t)op.groups};
    const cuuint64_t str[2] = {(cuuint64_t)op.ld0 * 4, (cuuint64_t)op.ld0 * 4 * p.rows_per_group};
    const cuuint32_t box[3] = {TF_BK, TF_BM, 1};
    KD_TRY(make_map_f32(&ma, op.a0, dims, str, box));
    const cuuint64_t wdims[3] = {(cuuint64_t)op.c0, (cuuint64_t)e.N, (cuuint64_t)op.groups};
    const cuuint64_t wstr[2] = {(cuuint64_t)op.w_ld * 4, (cuuint64_t)(op.groups > 1 ? op.w_group_stride : (long)op.w_ld * e.N) * 4};
    const cuuint32_t wbox[3] = {TF_BK, (cuuint32_t)p.nc, 1};
    KD_TRY(make_map_f32(&mw, op.w, wdims, wstr, wbox));
  }
  const int grid = (int)std::min<long>(p.items, (long)sms);
  ProfScope prof(PC_GEMM_TC, s, 2.0 * rows * e.N * op.c0 * p.taps,
                 4.0 * ((double)rows * (op.c0 + e.N * (e.res ? 2 : 1)) + (double)op.groups * e.N * op.c0 * p.taps));
  k_gemm_tf32<<<grid, TF_THREADS, TF_SMEM, s>>>(ma, mw, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
