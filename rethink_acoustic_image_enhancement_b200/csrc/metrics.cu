// Reductions of the validation / training hooks that sit directly behind the forward:
//   psnr_sse   : per-image sum of squared differences (+ max of the first image) for calculate_psnr
//                (Train/basicsr/metrics/psnr_ssim.py:9-70), optionally on the uint8 images tensor2img would produce
//                (img_util.py:67-94: clamp(0,1), *255, round half to even) and with crop_border - one HBM pass instead of
//                .cpu() + numpy per image (image_restoration_model.py:296-334)
//   l1_sr_loss : L1LossSr (Train/basicsr/models/losses/losses.py:135-194): 0.5*L1(hq) + 0.25*L1(sr) + 0.25*(shadow(hq)+shadow(sr)),
//                shadow = L1 of the > 0.1 binarisations (value only: its gradient is zero), forward value and d loss / d pred
//                in one pass per output.
// Sums are accumulated in double (per-thread fp32 partials over <= 16 elements, then double): deterministic for a given
// launch shape because the grid-level combine is a fixed-order second kernel, not atomics.
#include <algorithm>
#include "ops.cuh"

namespace kd {

namespace {

constexpr int RED_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < RED_THREADS / 32) t = sh[threadIdx.x];
  if (warp == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in thread 0
}
__device__ __forceinline__ float block_max(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = -INFINITY;
  if (threadIdx.x < RED_THREADS / 32) t = sh[threadIdx.x];
  if (warp == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
  }
  return t;
}

// tensor2img (img_util.py:67-94): clamp to [0,1], *255, numpy round (half to even) -> the uint8 image as a float
__device__ __forceinline__ float quant_u8(float v) { return rintf(fminf(fmaxf(v, 0.f), 1.f) * 255.0f); }

// grid = (blocks_per_image, B); part[(b*gridDim.x + blockIdx.x)] = {sse, max(a)} over the cropped window, all channels
__global__ void __launch_bounds__(RED_THREADS) k_psnr_part(const float* __restrict__ a, const float* __restrict__ b, int C, int H, int W,
                                                           int crop, int as_u8, double* __restrict__ part_sse,
                                                           float* __restrict__ part_max) {
  __shared__ double shd[RED_THREADS / 32];
  __shared__ float shf[RED_THREADS / 32];
  const int img = blockIdx.y;
  const int hc = H - 2 * crop, wc = W - 2 * crop;
  const long n = (long)C * hc * wc;
  const float* pa = a + (long)img * C * H * W;
  const float* pb = b + (long)img * C * H * W;
  double sse = 0.0;
  float mx = -INFINITY;
  for (long i0 = (long)blockIdx.x * RED_THREADS * 8; i0 < n; i0 += (long)gridDim.x * RED_THREADS * 8) {
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long i = i0 + (long)u * RED_THREADS + threadIdx.x;
      if (i < n) {
        const int c = (int)(i / ((long)hc * wc));
        const long r = i - (long)c * hc * wc;
        const int y = (int)(r / wc), x = (int)(r - (long)y * wc);
        const long o = ((long)c * H + y + crop) * W + x + crop;
        float va = pa[o], vb = pb[o];
        if (as_u8) { va = quant_u8(va); vb = quant_u8(vb); }
        const float d = va - vb;
        s = fmaf(d, d, s);
        mx = fmaxf(mx, va);
      }
    }
    sse += (double)s;
  }
  const double t = block_sum(sse, shd);
  const float m = block_max(mx, shf);
  if (threadIdx.x == 0) {
    part_sse[(long)img * gridDim.x + blockIdx.x] = t;
    part_max[(long)img * gridDim.x + blockIdx.x] = m;
  }
}

// out[b] = {mse, max(a)} (doubles): fixed-order combine of the per-block partials
__global__ void k_psnr_final(const double* __restrict__ part_sse, const float* __restrict__ part_max, int nblk, double inv_n,
                             double* __restrict__ out, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  float m = -INFINITY;
  for (int i = 0; i < nblk; ++i) { s += part_sse[(long)b * nblk + i]; m = fmaxf(m, part_max[(long)b * nblk + i]); }
  out[2 * b] = s * inv_n;
  out[2 * b + 1] = (double)m;
}

// part[blockIdx.x] = {sum |p - t|, sum |[p > 0.1] - [t > 0.1]|}; grad[i] = gscale * sign(p - t)
__global__ void __launch_bounds__(RED_THREADS) k_l1_shadow_part(const float* __restrict__ pred, const float* __restrict__ tgt, long n,
                                                                float gscale, float* __restrict__ grad,
                                                                double* __restrict__ part) {
  __shared__ double shd[RED_THREADS / 32];
  double s_abs = 0.0, s_bin = 0.0;
  for (long i0 = (long)blockIdx.x * RED_THREADS * 8; i0 < n; i0 += (long)gridDim.x * RED_THREADS * 8) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long i = i0 + (long)u * RED_THREADS + threadIdx.x;
      if (i < n) {
        const float p = pred[i], t = tgt[i];
        const float d = p - t;
        sa += fabsf(d);
        sb += ((p > 0.1f) != (t > 0.1f)) ? 1.f : 0.f;
        if (grad) grad[i] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);    // torch: sign(0) = 0
      }
    }
    s_abs += (double)sa;
    s_bin += (double)sb;
  }
  const double ta = block_sum(s_abs, shd);
  const double tb = block_sum(s_bin, shd);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = ta; part[2 * blockIdx.x + 1] = tb; }
}

// loss = w_l1 * S_abs / n + w_sh * S_bin / n, accumulated into *loss (so the hq and sr terms add up)
__global__ void k_l1_shadow_final(const double* __restrict__ part, int nblk, double inv_n, float w_l1, float w_sh, int accumulate,
                                  float* __restrict__ loss, double* __restrict__ terms) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double a = 0.0, b = 0.0;
  for (int i = 0; i < nblk; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
  a *= inv_n; b *= inv_n;
  if (terms) { terms[0] = a; terms[1] = b; }
  const float v = (float)(w_l1 * a + w_sh * b);
  *loss = accumulate ? *loss + v : v;
}

}  // namespace

size_t psnr_scratch_bytes(int B) { return (size_t)B * 64 * (sizeof(double) + sizeof(float)) + 256; }

int psnr_mse(const float* a, const float* b, int B, int C, int H, int W, int crop_border, int as_u8, double* out, void* scratch,
             cudaStream_t s) {
  KD_CHECK(B >= 1 && C >= 1 && H >= 1 && W >= 1 && crop_border >= 0 && 2 * crop_border < H && 2 * crop_border < W,
           "psnr: bad shape %dx%dx%dx%d crop %d", B, C, H, W, crop_border);
  const long n = (long)C * (H - 2 * crop_border) * (W - 2 * crop_border);
  const int nblk = (int)std::min<long>(64, (n + RED_THREADS * 8 - 1) / (RED_THREADS * 8));
  double* part_sse = reinterpret_cast<double*>(scratch);
  float* part_max = reinterpret_cast<float*>(part_sse + (size_t)B * 64);
  ProfScope prof(PC_POOL_RESAMPLE, s, 0.0, (double)B * n * 8.0);
  k_psnr_part<<<dim3(nblk, B), RED_THREADS, 0, s>>>(a, b, C, H, W, crop_border, as_u8, part_sse, part_max);
  count_launch();
  KD_LAUNCH_CHECK();
  k_psnr_final<<<cdiv(B, 128), 128, 0, s>>>(part_sse, part_max, nblk, 1.0 / (double)n, out, B);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

size_t l1_sr_scratch_bytes() { return 2 * 1024 * sizeof(double); }

// one output (hq or sr) of L1LossSr: adds w_l1 * mean|p - t| + w_sh * mean|bin(p) - bin(t)| to *loss (or sets it) and writes
// grad = d(that term)/d pred = w_l1 * sign(p - t) / n  (the shadow term is piecewise constant: zero gradient)
int l1_shadow_term(const float* pred, const float* tgt, long n, float w_l1, float w_sh, int accumulate, float* loss, float* grad,
                   double* terms, void* scratch, cudaStream_t s) {
  KD_CHECK(pred && tgt && loss && scratch && n >= 1, "l1_shadow_term: bad argument");
  const int nblk = (int)std::min<long>(1024, (n + RED_THREADS * 8 - 1) / (RED_THREADS * 8));
  double* part = reinterpret_cast<double*>(scratch);
  ProfScope prof(PC_POOL_RESAMPLE, s, 0.0, (double)n * (grad ? 12.0 : 8.0));
  k_l1_shadow_part<<<nblk, RED_THREADS, 0, s>>>(pred, tgt, n, (float)((double)w_l1 / (double)n), grad, part);
  count_launch();
  KD_LAUNCH_CHECK();
  k_l1_shadow_final<<<1, 32, 0, s>>>(part, nblk, 1.0 / (double)n, w_l1, w_sh, accumulate, loss, terms);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
