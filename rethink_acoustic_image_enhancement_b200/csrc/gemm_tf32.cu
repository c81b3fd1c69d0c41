// TF32 tensor-core GEMM for the 1x1 convolutions of the training step (fp32 storage, tcgen05 kind::tf32, fp32 accumulation):
//   out[p][n] = sum_k A[p][k] * W[n][k] (+ res[p][n]),   A [P][K] fp32 (row stride ld0), W [N][K] fp32 (per-image groups optional)
// Opt-in (training.set_matmul_precision("tf32") / KDLAE_TRAIN_TF32=1): PyTorch itself runs the reference's nn.Conv2d layers in
// TF32 on an Ampere-or-newer GPU by default (torch.backends.cudnn.allow_tf32), the fp32 CUDA-core path stays the default here
// because the gradient-parity tests hold it to 1e-5 of float64 autograd.
// Operands arrive by TMA as 128-byte-swizzled K-major tiles of 32 floats per row (one swizzle atom), four K = 8 MMAs per tile;
// persistent CTAs, warp 0 producer, warp 1 MMA issuer + TMEM allocator, warps 2-5 epilogue (one TMEM lane quarter each),
// two 256-column accumulators so the epilogue of a tile overlaps the MMAs of the next.  K and N tails are TMA zero fill.
// (wgrad needs MN-major operands - the reduction runs over pixels.  A kind::tf32 MMA over 128B-swizzled MN-major fp32 tiles, the
// layout gram_tc.cu uses for bf16, returned zeros on B200: 32-bit MN-major operands need the 32-byte-atom swizzle mode
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B / descriptor layout type 1), not brought up here - wgrad stays on the CUDA cores.)
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

  auto coord = [&](long item64, int& g, int& r0, int& nchunk) {
    const int item = (int)item64;
    const int mt = fast_div(item, p.n_chunks, p.inv_n_chunks);
    nchunk = item - mt * p.n_chunks;
    g = fast_div(mt, p.tiles_per_group, p.inv_tiles_per_group);
    r0 = (mt - g * p.tiles_per_group) * TF_BM;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    int s = 0; uint32_t ph = 0;
    const uint32_t stage_tx = TF_A_BYTES + (uint32_t)p.nc * 128;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
      int g, r0, nchunk;
      coord(item, g, r0, nchunk);
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait_relaxed(empty_bar(s), ph ^ 1);
        if (elect_one()) {
          const uint32_t a_dst = smem_base + s * TF_STAGE;
          mbar_expect_tx(full_bar(s), stage_tx);
          tma_load_3d(a_dst, &map_a, full_bar(s), kc * TF_BK, r0, g);
          tma_load_3d(a_dst + TF_A_BYTES, &map_w, full_bar(s), kc * TF_BK, nchunk * p.nc, g);
        }
        __syncwarp();
        if (++s == TF_STAGES) { s = 0; ph ^= 1; }
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
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_lo = (((smem_base + s * TF_STAGE) & 0x3FFFF) >> 4) | lo_tag;
        const uint32_t b_lo = (((smem_base + s * TF_STAGE + TF_A_BYTES) & 0x3FFFF) >> 4) | lo_tag;
        const int rem = p.K - kc * TF_BK;
        const int ksn = rem >= TF_BK ? 4 : (rem + 7) >> 3;      // K = 8 steps that hold real channels (the rest is zero fill)
        const uint32_t first = kc != 0 ? 1u : 0u;
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
      const long rr = (long)r0 + r;
      const bool valid = rr < p.rows_per_group;
      const long prow = (long)g * p.rows_per_group + rr;
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

int make_map_f32(CUtensorMap* m, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  KD_CHECK(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KD_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (fp32) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

// 1x1, one source, identity addressing, no row scale / bias / ReLU / statistics: what the training step's GEMMs use
bool gemm_tf32_eligible(const ConvOp& op) {
  const Epilogue& e = op.epi;
  return op.kd == 1 && op.kh == 1 && op.kw == 1 && op.c1 == 0 && op.xpack_cin == 0 && e.mode == OUT_IDENTITY && e.row_scale == nullptr &&
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
  KD_CHECK(p.items < (1L << 24), "gemm_tf32: too many tiles (%ld)", p.items);
  p.inv_n_chunks = 1.0f / (float)p.n_chunks; p.inv_tiles_per_group = 1.0f / (float)p.tiles_per_group;
  p.res = reinterpret_cast<const float*>(e.res); p.res_ld = e.res_ld;
  p.out = reinterpret_cast<float*>(e.out); p.out_ld = e.out_ld; p.out_coff = (int)e.out_coff;
  CUtensorMap ma, mw;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)op.c0, (cuuint64_t)p.rows_per_group, (cuuint64_t)op.groups};
    const cuuint64_t str[2] = {(cuuint64_t)op.ld0 * 4, (cuuint64_t)op.ld0 * 4 * p.rows_per_group};
    const cuuint32_t box[3] = {TF_BK, TF_BM, 1};
    KD_TRY(make_map_f32(&ma, op.a0, dims, str, box));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)op.c0, (cuuint64_t)e.N, (cuuint64_t)op.groups};
    const cuuint64_t str[2] = {(cuuint64_t)op.w_ld * 4, (cuuint64_t)(op.groups > 1 ? op.w_group_stride : (long)op.w_ld * e.N) * 4};
    const cuuint32_t box[3] = {TF_BK, (cuuint32_t)p.nc, 1};
    KD_TRY(make_map_f32(&mw, op.w, dims, str, box));
  }
  const int grid = (int)std::min<long>(p.items, (long)sms);
  ProfScope prof(PC_GEMM_TC, s, 2.0 * rows * e.N * op.c0, 4.0 * ((double)rows * (op.c0 + e.N * (e.res ? 2 : 1)) + (double)op.groups * e.N * op.c0));
  k_gemm_tf32<<<grid, TF_THREADS, TF_SMEM, s>>>(ma, mw, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
