// Throughput probe: FFMA vs FFMA2 (fma.rn.f32x2) per SM on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 fma_probe.cu -o fma_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, int iters, float w) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  float x = w + threadIdx.x * 1e-6f, y = w * 0.5f;
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
    }
  } else {
    u64 a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(acc[2 * i], acc[2 * i + 1]); a2[i] = *reinterpret_cast<u64*>(&t); }
    float2 xx = make_float2(x, x), yy = make_float2(y, y);
    const u64 x2 = *reinterpret_cast<u64*>(&xx), y2 = *reinterpret_cast<u64*>(&yy);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a2[i] = ffma2(a2[i], x2, y2);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&a2[i]); acc[2 * i] = t.x; acc[2 * i + 1] = t.y; }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<148 * 4, 512>>>(out, iters, 0.999f); else k<1><<<148 * 4, 512>>>(out, iters, 0.999f);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      const double fma = 148.0 * 4 * 512 * 16.0 * iters;
      if (rep) printf("%s: %.3f ms, %.1f GFMA/s, %.1f FMA/clk/SM @1.9GHz\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.9);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
