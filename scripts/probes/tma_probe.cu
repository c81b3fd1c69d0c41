// TMA load throughput probe (sm_100a): persistent CTAs, one elected thread streams boxes of a bf16 NHWC tensor into a ring of
// smem slots and a second thread frees each slot as soon as its mbarrier completes - nothing else runs.  Reports bytes per
// clock per SM and GB/s for several box shapes / ring depths.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tma_probe.cu -o tma_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}

// mode 0: 4-D box {64, bx, by, 1} at (0, x, y, img) ; mode 1: 3-D box {64, rows, 1} over the tensor seen as [pixels][C]
__global__ void __launch_bounds__(640, 1) k_probe(const __grid_constant__ CUtensorMap map, int mode, int slots, uint32_t box_bytes,
                                                 int tiles_x, int tiles_y, int bx_step, int by_step, int rows, long items, int use_commit) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slot_bytes = (box_bytes + 1023u) & ~1023u;
  const uint32_t bar = sbase + slots * slot_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s) { mbar_init(bar + 8 * s, 1); mbar_init(bar + 8 * (16 + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {          // producer
    uint32_t i = 0;
    for (long item = blockIdx.x; item < items; item += gridDim.x, ++i) {
      const int s = i % slots;
      mbar_wait(bar + 8 * (16 + s), ((i / slots) & 1) ^ 1);
      mbar_expect(bar + 8 * s, box_bytes);
      if (mode == 0) {
        const int per = tiles_x * tiles_y;
        const int img = (int)(item / per), r = (int)(item % per), ty = r / tiles_x, tx = r % tiles_x;
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(sbase + s * slot_bytes), "l"(&map), "r"(bar + 8 * s), "r"(0), "r"(tx * bx_step - 1), "r"(ty * by_step - 1), "r"(img) : "memory");
      } else {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(sbase + s * slot_bytes), "l"(&map), "r"(bar + 8 * s), "r"(0), "r"((int)(item * rows)), "r"(0) : "memory");
      }
    }
  } else if (threadIdx.x == 32) {  // consumer: free the slot as soon as the data has landed
    uint32_t i = 0;
    for (long item = blockIdx.x; item < items; item += gridDim.x, ++i) {
      const int s = i % slots;
      mbar_wait(bar + 8 * s, (i / slots) & 1);
      if (use_commit) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar + 8 * (16 + s)) : "memory");
      else mbar_arrive(bar + 8 * (16 + s));
    }
  }
  if (blockDim.x > 64) __syncthreads();     // kernel-like: idle lanes / warps park at a CTA barrier while the two threads work
}

__global__ void k_fill(uint32_t* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)i * 2654435761u; v ^= v >> 13; p[i] = (v & 0x7fff7fffu) | 0x30003000u;   // random finite bf16 pairs
  }
}

int main() {
  EncodeTiledFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  const int n = 32, H = 512, W = 512, C = 64;
  void* x; cudaMalloc(&x, (size_t)n * H * W * C * 2);
  k_fill<<<1184, 256>>>((uint32_t*)x, (size_t)n * H * W * C / 2);   // random data (all-zero pages may be compressed)
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  struct Cfg { int mode, bx, by, rows, slots; CUtensorMapSwizzle sw; const char* name; int big_smem; };
  const Cfg cfgs[] = {
      {0, 32, 6, 0, 2, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, step 30x4, 2 slots"},
      {0, 32, 6, 0, 4, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, step 30x4, 4 slots"},
      {0, 32, 6, 0, 2, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, 2 slots, 215 KB smem CTA", 1},
      {0, 32, 6, 0, 2, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, 2 slots, slot freed by tcgen05.commit", 2},
      {0, 32, 6, 0, 2, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, 2 slots, 640 threads parked at bar.sync", 3},
      {0, 32, 6, 0, 4, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, 4 slots, slot freed by tcgen05.commit", 2},
      {0, 32, 6, 0, 8, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,6}, step 30x4, 8 slots"},
      {0, 32, 8, 0, 4, CU_TENSOR_MAP_SWIZZLE_128B, "4-D halo box {64,32,8}, step 30x6, 4 slots"},
      {0, 32, 6, 0, 4, CU_TENSOR_MAP_SWIZZLE_NONE, "4-D halo box {64,32,6}, no swizzle, 4 slots"},
      {1, 0, 0, 128, 2, CU_TENSOR_MAP_SWIZZLE_128B, "3-D linear box {64,128 rows}, 2 slots"},
      {1, 0, 0, 128, 4, CU_TENSOR_MAP_SWIZZLE_128B, "3-D linear box {64,128 rows}, 4 slots"},
      {1, 0, 0, 128, 8, CU_TENSOR_MAP_SWIZZLE_128B, "3-D linear box {64,128 rows}, 8 slots"},
      {1, 0, 0, 256, 4, CU_TENSOR_MAP_SWIZZLE_128B, "3-D linear box {64,256 rows}, 4 slots"},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap m; cuuint32_t es[5] = {1, 1, 1, 1, 1};
    long items; uint32_t box_bytes; int tx = 0, ty = 0, sx = 0, sy = 0;
    if (c.mode == 0) {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
      cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * W, (cuuint64_t)C * 2 * W * H};
      cuuint32_t box[4] = {64, (cuuint32_t)c.bx, (cuuint32_t)c.by, 1};
      enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      sx = c.bx - 2; sy = c.by - 2; tx = (W + sx - 1) / sx; ty = (H + sy - 1) / sy;
      items = (long)n * tx * ty; box_bytes = 128u * c.bx * c.by;
    } else {
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n * H * W, 1};
      cuuint64_t str[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * n * H * W};
      cuuint32_t box[3] = {64, (cuuint32_t)c.rows, 1};
      enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      items = (long)n * H * W / c.rows; box_bytes = 128u * c.rows;
    }
    uint32_t smem = c.slots * ((box_bytes + 1023u) & ~1023u) + 1024 + 512;
    if (c.big_smem == 1) smem = 215 * 1024;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      k_probe<<<sms, c.big_smem == 3 ? 640 : 64, smem>>>(m, c.mode, c.slots, box_bytes, tx, ty, sx, sy, c.rows, items, c.big_smem == 2);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep == 2) {
        const double bytes = (double)items * box_bytes;
        printf("%-48s %8.1f us  %7.0f GB/s into smem  %5.1f B/clk/SM  %5.1f clk per 128-B row  (%s)\n", c.name, ms * 1e3, bytes / ms / 1e6,
               bytes / (ms * 1e-3 * 1.9e9) / sms, (ms * 1e-3 * 1.9e9) / ((double)items / sms * box_bytes / 128), cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
