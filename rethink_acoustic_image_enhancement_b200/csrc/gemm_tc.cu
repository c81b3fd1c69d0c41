// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
//   out[pixel, n] = epilogue( sum_{tap, c} A[shift(pixel, tap), c] * W[n, tap, c] )
//
// * A (NHWC activations) is staged by TMA into 128B-swizzled K-major smem tiles of 128 pixels x 64
//   channels.  1x1 convs use a linear-row tensor map; 3x3 convs use a 4-D (C, W, H, image) map whose
//   box is a 16x8 pixel patch shifted by the tap offset - the conv's zero padding is TMA's
//   out-of-bounds fill, so there is no im2col buffer.  Channel tails (C % 64 != 0) and the channel
//   concat of two sources (skip connections) are also handled by TMA zero fill / two maps.
// * W tiles ([n-chunk <= 256] x 64, K-major) come from a 3-D map (k, n, group); group = image for
//   the per-image attention matrices (MDTA folded into project_out).
// * One elected thread issues tcgen05.mma (M=128, N=n-chunk, K=16) into one of two TMEM accumulator
//   buffers (2 x 256 columns); 8 epilogue warps drain the other buffer with tcgen05.ld and apply the
//   fused epilogue (LayerNorm row scale, bias, ReLU, residual, PixelShuffle/Unshuffle addressing).
// * Persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer + TMEM
//   allocator, warps 2..9 epilogue.  4-stage smem ring, mbarrier full/empty pipelines.
#include <cuda.h>
#include "ops.cuh"

namespace kd {

namespace {

constexpr int TC_BM = 128;         // pixels per tile (UMMA M)
constexpr int TC_BK = 64;          // K elements per stage (one 128B swizzle atom of bf16)
constexpr int TC_NC_MAX = 256;     // max N per accumulator (UMMA N)
constexpr int TC_STAGES = 4;
constexpr int TC_TW = 16, TC_TH = 8;   // spatial tile of the 3x3 path
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + TC_EPI_WARPS * 32;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr uint32_t TC_B_BYTES = TC_NC_MAX * TC_BK * 2;
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TcParams {
  int spatial;          // 0: linear rows (1x1), 1: 16x8 patches with 3x3 taps
  int taps, kw, dil;
  int kc0, kc1;         // 64-wide K chunks per tap from source 0 / 1
  int c0;               // weight column offset of source 1
  long w_tap_ld;
  int n_chunks, nc;     // N split
  long items;           // tiles_m * n_chunks
  // linear mode
  long rows_per_group;
  int tiles_per_group;
  // spatial mode
  int H, W, tiles_x, tiles_y;
  Epilogue epi;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 200000000u) {
      printf("kdlae gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
// 8-row x 128-byte swizzle atoms, atoms stacked every 1024 bytes along M/N (SBO), LBO unused.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // layout type: SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_gemm_tc(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
  // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.kc1 > 0) prefetch_tmap(&map_a1);
    prefetch_tmap(&map_w);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TC_EPI_WARPS); }
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

  const int kchunks = p.kc0 + p.kc1;
  const int kblocks = p.taps * kchunks;
  const uint32_t stage_tx = TC_A_BYTES + (uint32_t)p.nc * TC_BK * 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t kidx = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const long mt = item / p.n_chunks;
        const int nchunk = (int)(item % p.n_chunks);
        int g = 0, r0 = 0, tx0 = 0, ty0 = 0, img = 0;
        if (p.spatial) {
          const int txi = (int)(mt % p.tiles_x);
          const int tyi = (int)((mt / p.tiles_x) % p.tiles_y);
          img = (int)(mt / ((long)p.tiles_x * p.tiles_y));
          tx0 = txi * TC_TW; ty0 = tyi * TC_TH;
        } else {
          g = (int)(mt / p.tiles_per_group);
          r0 = (int)(mt % p.tiles_per_group) * TC_BM;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dx = (tap % p.kw - p.kw / 2) * p.dil, dy = (tap / p.kw - p.kw / 2) * p.dil;
          for (int kc = 0; kc < kchunks; ++kc, ++kidx) {
            const int s = kidx % TC_STAGES;
            const uint32_t ph = (kidx / TC_STAGES) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            const uint32_t a_dst = smem_base + s * TC_STAGE_BYTES;
            const uint32_t b_dst = a_dst + TC_A_BYTES;
            mbar_expect_tx(full_bar(s), stage_tx);
            const bool src1 = kc >= p.kc0;
            const CUtensorMap* ma = src1 ? &map_a1 : &map_a0;
            const int cc = (src1 ? kc - p.kc0 : kc) * TC_BK;
            if (p.spatial) tma_load_4d(a_dst, ma, full_bar(s), cc, tx0 + dx, ty0 + dy, img);
            else tma_load_3d(a_dst, ma, full_bar(s), cc, r0, g);
            const int wk = (int)(tap * p.w_tap_ld) + (src1 ? p.c0 : 0) + cc;
            tma_load_3d(b_dst, &map_w, full_bar(s), wk, nchunk * p.nc, g);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.nc);
      uint32_t kidx = 0, it = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_NC_MAX;
        for (int kb = 0; kb < kblocks; ++kb, ++kidx) {
          const int s = kidx % TC_STAGES;
          const uint32_t ph = (kidx / TC_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * TC_STAGE_BYTES;
          const uint32_t b_addr = a_addr + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t ad = make_smem_desc(a_addr + k * 32);
            const uint64_t bd = make_smem_desc(b_addr + k * 32);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));      // frees the smem stage when these MMAs retire
        }
        umma_commit(tfull_bar(acc));      // accumulator complete
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = ew >> 2;             // which half of the 32-column groups
    const int r = quarter * 32 + lane;    // accumulator row (pixel within the tile)
    uint32_t it = 0;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const long mt = item / p.n_chunks;
      const int nchunk = (int)(item % p.n_chunks);
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      long prow; int img = 0, y = 0, x = 0; bool valid;
      if (p.spatial) {
        const int txi = (int)(mt % p.tiles_x);
        const int tyi = (int)((mt / p.tiles_x) % p.tiles_y);
        img = (int)(mt / ((long)p.tiles_x * p.tiles_y));
        x = txi * TC_TW + (r % TC_TW); y = tyi * TC_TH + (r / TC_TW);
        valid = (x < p.W) && (y < p.H);
        prow = ((long)img * p.H + y) * p.W + x;
      } else {
        const int g = (int)(mt / p.tiles_per_group);
        const long rr = (long)(mt % p.tiles_per_group) * TC_BM + r;
        valid = rr < p.rows_per_group;
        prow = (long)g * p.rows_per_group + rr;
      }
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * TC_NC_MAX + ((uint32_t)(quarter * 32) << 16);
      const int ngroups = (p.nc + 31) / 32;
      for (int cgp = half; cgp < ngroups; cgp += 2) {
        uint32_t v[32];
        tmem_ld32(t_row + cgp * 32, v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = cgp * 32 + j * 8;
            if (col < p.nc) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[j * 8 + i]);
              epilogue_store8<bf16>(p.epi, prow, img, y, x, nchunk * p.nc + col, f);
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

// ---------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
             const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  KD_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KD_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu %llu)", (int)r, rank,
           (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0));
  return 0;
}

int pick_nc(int N, int* n_chunks) {
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= TC_NC_MAX) { *n_chunks = 1; return n16; }
  const int chunks = (n16 + TC_NC_MAX - 1) / TC_NC_MAX;
  int nc = ((n16 + chunks - 1) / chunks + 15) / 16 * 16;
  *n_chunks = (N + nc - 1) / nc;
  return nc;
}

int g_num_sms = 0;

}  // namespace

bool conv_gemm_tc_eligible(const ConvOp& op) {
  const bool k11 = (op.kd == 1 && op.kh == 1 && op.kw == 1);
  const bool k33 = (op.kd == 1 && op.kh == 3 && op.kw == 3);
  if (!k11 && !k33) return false;
  if (op.c0 % 8 || op.ld0 % 8 || (reinterpret_cast<uintptr_t>(op.a0) & 15)) return false;
  if (op.c1 > 0 && (op.c1 % 8 || op.ld1 % 8 || (reinterpret_cast<uintptr_t>(op.a1) & 15))) return false;
  if (op.w_ld % 8 || op.w_tap_ld % 8 || (reinterpret_cast<uintptr_t>(op.w) & 15) || op.w_group_stride % 8) return false;
  if (k33 && op.groups != 1) return false;
  if (k11 && op.epi.mode != OUT_IDENTITY) return false;
  if (op.epi.mode == OUT_PIXEL_SHUFFLE && (op.epi.cq % 8 != 0)) return false;
  return true;
}

int conv_gemm_tc(const ConvOp& op, cudaStream_t s) {
  KD_CHECK(conv_gemm_tc_eligible(op), "conv_gemm_tc: shape not eligible");
  if (g_num_sms == 0) {
    int dev = 0;
    KD_CUDA(cudaGetDevice(&dev));
    KD_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    KD_CUDA(cudaFuncSetAttribute(k_conv_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.spatial = (op.kh == 3) ? 1 : 0;
  p.taps = op.kh * op.kw;
  p.kw = op.kw;
  p.dil = op.dil;
  p.kc0 = (op.c0 + TC_BK - 1) / TC_BK;
  p.kc1 = (op.c1 + TC_BK - 1) / TC_BK;
  p.c0 = op.c0;
  p.w_tap_ld = op.w_tap_ld;
  p.nc = pick_nc(op.epi.N, &p.n_chunks);
  p.H = op.H; p.W = op.W;
  p.epi = op.epi;

  CUtensorMap ma0, ma1, mw;
  const long rows = (long)op.nimg * op.H * op.W;
  long tiles_m;
  if (p.spatial) {
    p.tiles_x = cdiv(op.W, TC_TW);
    p.tiles_y = cdiv(op.H, TC_TH);
    tiles_m = (long)op.nimg * p.tiles_x * p.tiles_y;
    const cuuint32_t box[4] = {TC_BK, TC_TW, TC_TH, 1};
    {
      const cuuint64_t dims[4] = {(cuuint64_t)op.c0, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.nimg};
      const cuuint64_t str[3] = {(cuuint64_t)op.ld0 * 2, (cuuint64_t)op.ld0 * 2 * op.W, (cuuint64_t)op.ld0 * 2 * op.W * op.H};
      KD_TRY(make_map(&ma0, op.a0, 4, dims, str, box));
    }
    if (op.c1 > 0) {
      const cuuint64_t dims[4] = {(cuuint64_t)op.c1, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.nimg};
      const cuuint64_t str[3] = {(cuuint64_t)op.ld1 * 2, (cuuint64_t)op.ld1 * 2 * op.W, (cuuint64_t)op.ld1 * 2 * op.W * op.H};
      KD_TRY(make_map(&ma1, op.a1, 4, dims, str, box));
    } else {
      ma1 = ma0;
    }
  } else {
    KD_CHECK(rows % op.groups == 0, "conv_gemm_tc: rows %ld not divisible by groups %d", rows, op.groups);
    p.rows_per_group = rows / op.groups;
    p.tiles_per_group = cdiv(p.rows_per_group, TC_BM);
    tiles_m = (long)p.tiles_per_group * op.groups;
    const cuuint32_t box[3] = {TC_BK, TC_BM, 1};
    {
      const cuuint64_t dims[3] = {(cuuint64_t)op.c0, (cuuint64_t)p.rows_per_group, (cuuint64_t)op.groups};
      const cuuint64_t str[2] = {(cuuint64_t)op.ld0 * 2, (cuuint64_t)op.ld0 * 2 * p.rows_per_group};
      KD_TRY(make_map(&ma0, op.a0, 3, dims, str, box));
    }
    if (op.c1 > 0) {
      const cuuint64_t dims[3] = {(cuuint64_t)op.c1, (cuuint64_t)p.rows_per_group, (cuuint64_t)op.groups};
      const cuuint64_t str[2] = {(cuuint64_t)op.ld1 * 2, (cuuint64_t)op.ld1 * 2 * p.rows_per_group};
      KD_TRY(make_map(&ma1, op.a1, 3, dims, str, box));
    } else {
      ma1 = ma0;
    }
  }
  {
    const long w_row = (p.taps > 1) ? (long)p.taps * op.w_tap_ld : (long)(op.c0 + op.c1);
    const cuuint64_t dims[3] = {(cuuint64_t)w_row, (cuuint64_t)op.epi.N, (cuuint64_t)op.groups};
    const cuuint64_t str[2] = {(cuuint64_t)op.w_ld * 2,
                               (cuuint64_t)(op.groups > 1 ? op.w_group_stride : (long)op.w_ld * op.epi.N) * 2};
    const cuuint32_t box[3] = {TC_BK, (cuuint32_t)p.nc, 1};
    KD_TRY(make_map(&mw, op.w, 3, dims, str, box));
  }
  p.items = tiles_m * p.n_chunks;
  const int grid = (int)(p.items < (long)g_num_sms ? p.items : (long)g_num_sms);
  const double ktot = (double)p.taps * (op.c0 + op.c1);
  ProfScope prof(PC_GEMM_TC, s, 2.0 * rows * op.epi.N * ktot,
                 2.0 * ((double)rows * (op.c0 + op.c1 + op.epi.N * (op.epi.res ? 2 : 1)) + (double)op.groups * op.epi.N * ktot));
  k_conv_gemm_tc<<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(ma0, ma1, mw, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
