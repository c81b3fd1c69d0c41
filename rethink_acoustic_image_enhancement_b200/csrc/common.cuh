// Common device/host helpers for the KDLAE / ASDQE sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <atomic>

namespace kd {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local text returned by kdlae_last_error) -------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define KD_CHECK(cond, ...)                          \
  do {                                               \
    if (!(cond)) {                                   \
      ::kd::set_error(__VA_ARGS__);                  \
      return 1;                                      \
    }                                                \
  } while (0)

#define KD_CUDA(expr)                                                              \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      ::kd::set_error("%s:%d CUDA error %s (%s)", __FILE__, __LINE__,              \
                      cudaGetErrorName(_e), cudaGetErrorString(_e));               \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

#define KD_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

#define KD_LAUNCH_CHECK() KD_CUDA(cudaGetLastError())

// ---- launch counter (bench.py reports gpu_launches) -------------------------------------
extern std::atomic<unsigned long long> g_launch_count;
inline void count_launch() { g_launch_count.fetch_add(1, std::memory_order_relaxed); }

// ---- SM budget of the persistent kernels ---------------------------------------------------
// KDLAE_SM_LIMIT=n caps the grid of every persistent kernel at n SMs, so two forwards on two streams can run side by side
// on disjoint SM sets (scripts/two_stream_probe.py: overlapping write-bound with read-bound stages).  Default: all SMs.
inline int sm_limit(int sms) {
  static int lim = -1;
  if (lim < 0) { const char* e = getenv("KDLAE_SM_LIMIT"); lim = e ? atoi(e) : 0; }
  return (lim > 0 && lim < sms) ? lim : sms;
}

// ---- optional per-kernel-class CUDA-event profiler (bench.py roofline; off by default) ----------
enum ProfClass {
  PC_GEMM_TC = 0, PC_GEMM_SIMT, PC_DWCONV, PC_LN_STATS, PC_MDTA_GRAM, PC_MDTA_FOLD, PC_SMALL_CONV, PC_POOL_RESAMPLE, PC_HEAD,
  PC_PWDW,
  PC_COUNT
};
struct ProfScope {
  bool active = false;
  int slot = -1;
  cudaStream_t stream;
  ProfScope(int cls, cudaStream_t s, double flops, double bytes);
  ~ProfScope();
};

// ---- per-device one-time state ---------------------------------------------------------------
// cudaFuncSetAttribute and the SM count belong to the CURRENT device, so "done once" flags are kept per device: a module
// moved from cuda:0 to cuda:1 in one process (or nn.DataParallel threads, one device each) sets its kernels up again there.
constexpr int KD_MAX_DEVICES = 64;
struct DeviceOnce { std::atomic<unsigned long long> mask{0}; };
// *first = the current device has not been marked in `once` yet; *dev = current device
int device_first_use(DeviceOnce& once, bool* first, int* dev);
inline void device_mark(DeviceOnce& once, int dev) { once.mask.fetch_or(1ull << dev, std::memory_order_acq_rel); }
// SM count of the current device after KDLAE_SM_LIMIT (cached per device)
int device_sms(int* sms);

// ---- debug trace (kdlae_debug_trace_begin/end): an order-independent checksum of a stage's output rows after each launch,
// so two forwards that should be bit-identical can be compared stage by stage.  Off unless enabled through the C ABI.
void trace_point(const char* tag, const void* ptr, long rows, long row_bytes, long ld_bytes, cudaStream_t s);

inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- scalar conversion ------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// ---- 8-wide vector load/store (16 B for bf16, 32 B for fp32); pointers must be aligned ---
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)), erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7): branch free, 2 MUFU (rcp, ex2).
// (Measured alternatives that were slower on B200: Newton reciprocal on the FMA pipe, integer-rounded bf16 packing -
// these epilogues are bound by issued instructions, not by the XU pipe.)
__device__ __forceinline__ float gelu_fast(float x) {
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
template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float gelu_erf(float x) {  // exact GELU (F.gelu default)
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- epilogue shared by the SIMT and the tcgen05 implicit-GEMM kernels --------------------
// Row = one output pixel of the conv (b*D+d, y, x) in a D x H x W grid, column n = output channel.
enum OutMode { OUT_IDENTITY = 0, OUT_PIXEL_SHUFFLE = 1, OUT_PIXEL_UNSHUFFLE = 2, OUT_PLANAR_F32 = 3 };

struct Epilogue {
  const float* row_scale = nullptr;  // [rows]  LayerNorm rstd folded behind the GEMM
  const float* row_mu = nullptr;     // [rows]  WithBias LN: v -= rstd*mu*col_s1[n]
  const float* col_s1 = nullptr;     // [N]
  const float* col_bias = nullptr;   // [N]
  int relu = 0;
  const void* res = nullptr;         // residual / skip, indexed like the destination
  long res_ld = 0;
  void* out = nullptr;
  long out_ld = 0;                   // destination row stride in elements
  int out_coff = 0;                  // destination channel offset
  long out_y_ld = 0, res_y_ld = 0;   // tcgen05 spatial path, IDENTITY mode: element stride between image rows of the destination /
                                     // residual (0 = W * ld).  Lets a conv write every other row of a 2x larger map (ConvTranspose
                                     // 2x2 stride 2 = two 1x1 GEMMs, one per output-row phase, KDLAE-S upconv)
  int mode = OUT_IDENTITY;
  int cq = 0;                        // PIXEL_SHUFFLE: channels per sub-pixel (N = 4*cq, packed sub-pixel major)
  int H = 0, W = 0;                  // conv-output pixel grid (per image/frame)
  int N = 0;                         // valid output channels
  // optional: LayerNorm statistics of the produced rows (over the N channels) for the next fused LN
  // (tcgen05 FAST epilogue only, N in a single chunk): rstd = 1/sqrt(var + 1e-5), mu = mean
  float* stat_rstd = nullptr;
  float* stat_mu = nullptr;
  // OUT_PLANAR_F32: fp32 NCHW output (+ fp32 NCHW residual), element (img, n, y, x) at img*img_stride + n*ch_stride + y*W + x
  float* planar_out = nullptr; long planar_img = 0, planar_ch = 0;
  const float* planar_res = nullptr; long planar_res_img = 0, planar_res_ch = 0;
};

// Store 8 consecutive columns n0..n0+7 (n0 % 8 == 0) of one row. `img` = b*D+d, `prow` = linear row.
template <typename T>
__device__ __forceinline__ void epilogue_store8(const Epilogue& e, long prow, int img, int y, int x, int n0, float (&v)[8]) {
  if (n0 >= e.N) return;
  const float rs = e.row_scale ? e.row_scale[prow] : 1.0f;
  const float rmu = e.row_mu ? e.row_mu[prow] * rs : 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + i;
    // explicit operation order (no compiler-chosen contraction): the fused pwdw_f2 conversion warps repeat it bit for bit
    float t = __fmul_rn(v[i], rs);
    if (n < e.N) {
      if (e.row_mu) t = fmaf(-rmu, e.col_s1[n], t);
      if (e.col_bias) t = __fadd_rn(t, e.col_bias[n]);
    }
    v[i] = t;
  }
  if (e.mode == OUT_PLANAR_F32) {
    const long sp = (long)y * e.W + x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i;
      if (n < e.N) {
        float t = v[i];
        if (e.planar_res) t += e.planar_res[(long)img * e.planar_res_img + (long)n * e.planar_res_ch + sp];
        if (e.relu) t = fmaxf(t, 0.f);
        e.planar_out[(long)img * e.planar_img + (long)n * e.planar_ch + sp] = t;
      }
    }
    return;
  }
  T* out = reinterpret_cast<T*>(e.out);
  const T* res = reinterpret_cast<const T*>(e.res);
  if (e.mode == OUT_PIXEL_UNSHUFFLE) {
    // nn.PixelUnshuffle(2): out[c*4 + 2*(y&1) + (x&1), y/2, x/2] = conv[c, y, x]
    const long dpix = ((long)img * (e.H >> 1) + (y >> 1)) * (e.W >> 1) + (x >> 1);
    const int sub = ((y & 1) << 1) | (x & 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i;
      if (n < e.N) {
        float t = v[i];
        const long di = dpix * e.out_ld + e.out_coff + n * 4 + sub;
        if (res) t += to_f<T>(res[dpix * e.res_ld + n * 4 + sub]);
        if (e.relu) t = fmaxf(t, 0.f);
        out[di] = from_f<T>(t);
      }
    }
    return;
  }
  long dpix = prow;
  int c0 = n0;
  if (e.mode == OUT_PIXEL_SHUFFLE) {
    // nn.PixelShuffle(2) (and ConvTranspose 2x2 stride 2): weight rows are packed sub-pixel major,
    // n' = s*cq + c with s = 2*dy + dx  ->  out[c, 2y+dy, 2x+dx]
    const int s = n0 / e.cq;
    c0 = n0 - s * e.cq;
    dpix = ((long)img * (e.H * 2) + (2 * y + (s >> 1))) * (e.W * 2) + (2 * x + (s & 1));
  }
  T* dst = out + dpix * e.out_ld + e.out_coff + c0;
  const T* rsrc = res ? res + dpix * e.res_ld + c0 : nullptr;
  const bool full = (n0 + 8 <= e.N);
  const bool vec_ok = full && ((reinterpret_cast<uintptr_t>(dst) & (sizeof(T) * 8 - 1)) == 0) &&
                      (!rsrc || (reinterpret_cast<uintptr_t>(rsrc) & (sizeof(T) * 8 - 1)) == 0);
  if (vec_ok) {
    if (rsrc) {
      float r[8];
      load8<T>(rsrc, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    if (e.relu) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    store8<T>(dst, v);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (n0 + i < e.N) {
        float t = v[i];
        if (rsrc) t += to_f<T>(rsrc[i]);
        if (e.relu) t = fmaxf(t, 0.f);
        dst[i] = from_f<T>(t);
      }
    }
  }
}

}  // namespace kd
