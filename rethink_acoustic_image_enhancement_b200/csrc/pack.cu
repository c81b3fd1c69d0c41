// Weight re-layout kernels, run once per load_state_dict: PyTorch [N][Cin][taps] conv weights ->
// GEMM operand layout [N'][taps][C'] (K-major, channel-padded), with LayerNorm / BatchNorm folds,
// the FFN half padding (127 -> 128 ...), PixelShuffle row permutation and ConvTranspose 2x2 phases.
#include "ops.cuh"

namespace kd {

template <typename T>
__global__ void __launch_bounds__(256) k_pack_weights(const PackOp op) {
  const long total = (long)op.n_dst * op.taps * op.c_dst;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % op.c_dst);
  const int t = (int)((idx / op.c_dst) % op.taps);
  const int n = (int)(idx / ((long)op.c_dst * op.taps));
  int ns = n, ks = k;
  if (op.mode == PACK_HALVES) {
    // two halves of h channels each padded to hp:  dst j in [0,hp) -> src j ; dst hp+j -> src h+j
    int& v = op.halves_on_k ? ks : ns;
    const int half = v / op.hp, j = v % op.hp;
    v = (j < op.h && half < 2) ? half * op.h + j : -1;
  } else if (op.mode == PACK_PIXEL_SHUFFLE) {
    // dst row n' = s*cq + c  <-  src row 4*c + s   (nn.PixelShuffle(2) channel order)
    const int cq = op.n_src / 4;
    const int s = n / cq, c = n % cq;
    ns = (s < 4) ? 4 * c + s : -1;
  }
  float v = 0.f;
  if (op.mode == PACK_CONVT) {
    // ConvTranspose3d weight [Cin][Cout][1][2][2]; dst row n' = s*Cout + co, k = ci (taps == 1)
    const int cout = op.n_src;
    const int s = n / cout, co = n % cout;
    if (s < 4 && k < op.c_src) v = op.src[((long)k * cout + co) * 4 + s];
  } else if (ns >= 0 && ns < op.n_src && ks >= 0 && ks < op.c_src) {
    v = op.src[((long)ns * op.c_src + ks) * op.taps + t];
    if (op.kscale) v *= op.kscale[ks];
    if (op.nscale) v *= op.nscale[ns];
  }
  reinterpret_cast<T*>(op.dst)[idx] = from_f<T>(v);
}

template <typename T>
int pack_weights(const PackOp& op, cudaStream_t s) {
  const long total = (long)op.n_dst * op.taps * op.c_dst;
  k_pack_weights<T><<<cdiv(total, 256), 256, 0, s>>>(op);
  KD_LAUNCH_CHECK();
  return 0;
}
template int pack_weights<float>(const PackOp&, cudaStream_t);
template int pack_weights<bf16>(const PackOp&, cudaStream_t);

// WithBias-LN fold column vectors for a packed 1x1 weight (taps == 1, PLAIN or HALVES-on-N):
//   s1[n] = sum_k W'[n][k] (packed, already rounded to T),  s2[n] = sum_k lnb[k] * W[src(n)][k]
template <typename T>
__global__ void k_pack_ln_cols(const PackOp op, const float* __restrict__ lnb, float* __restrict__ s1, float* __restrict__ s2) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= op.n_dst) return;
  int ns = n;
  if (op.mode == PACK_HALVES && !op.halves_on_k) {
    const int half = n / op.hp, j = n % op.hp;
    ns = (j < op.h && half < 2) ? half * op.h + j : -1;
  }
  const T* wp = reinterpret_cast<const T*>(op.dst) + (long)n * op.c_dst;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < op.c_dst; ++k) {
    a += to_f<T>(wp[k]);
    if (ns >= 0 && ns < op.n_src && k < op.c_src) b = fmaf(lnb[k], op.src[(long)ns * op.c_src + k], b);
  }
  s1[n] = a;
  s2[n] = b;
}
template <typename T>
int pack_ln_cols(const PackOp& op, const float* lnb, float* s1, float* s2, cudaStream_t s) {
  KD_CHECK(op.taps == 1 && (op.mode == PACK_PLAIN || (op.mode == PACK_HALVES && !op.halves_on_k)), "pack_ln_cols: unsupported pack mode");
  k_pack_ln_cols<T><<<cdiv(op.n_dst, 128), 128, 0, s>>>(op, lnb, s1, s2);
  KD_LAUNCH_CHECK();
  return 0;
}
template int pack_ln_cols<float>(const PackOp&, const float*, float*, float*, cudaStream_t);
template int pack_ln_cols<bf16>(const PackOp&, const float*, float*, float*, cudaStream_t);

__global__ void k_pack_dw(const float* __restrict__ src, int c_src, int h, int hp, float* __restrict__ dst, int c_dst, int taps) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= taps * c_dst) return;
  const int c = idx % c_dst, t = idx / c_dst;
  int cs = c;
  if (hp > 0) {
    const int half = c / hp, j = c % hp;
    cs = (j < h && half < 2) ? half * h + j : -1;
  }
  dst[idx] = (cs >= 0 && cs < c_src) ? src[(long)cs * taps + t] : 0.f;
}
int pack_dw(const float* src, int c_src, int h, int hp, float* dst, int c_dst, cudaStream_t s) {
  k_pack_dw<<<cdiv(9 * c_dst, 256), 256, 0, s>>>(src, c_src, h, hp, dst, c_dst, 9);
  KD_LAUNCH_CHECK();
  return 0;
}
int pack_dw_bias(const float* src, int c_src, int h, int hp, float* dst, int c_dst, cudaStream_t s) {
  k_pack_dw<<<cdiv(c_dst, 256), 256, 0, s>>>(src, c_src, h, hp, dst, c_dst, 1);
  KD_LAUNCH_CHECK();
  return 0;
}

__global__ void k_pack_few_in(const float* __restrict__ src, int cout, int cin, int taps, const float* __restrict__ nscale,
                              float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * cin * taps) return;
  const int co = idx % cout, c = (idx / cout) % cin, t = idx / (cout * cin);
  float v = src[((long)co * cin + c) * taps + t];
  if (nscale) v *= nscale[co];
  dst[idx] = v;
}
int pack_few_in(const float* src, int cout, int cin, int taps, const float* nscale, float* dst, cudaStream_t s) {
  k_pack_few_in<<<cdiv(cout * cin * taps, 256), 256, 0, s>>>(src, cout, cin, taps, nscale, dst);
  KD_LAUNCH_CHECK();
  return 0;
}
__global__ void k_pack_few_out(const float* __restrict__ src, int cout, int cin, int taps, float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= cout * cin * taps) return;
  const int c = idx % cin, t = (idx / cin) % taps, co = idx / (cin * taps);
  dst[idx] = src[((long)co * cin + c) * taps + t];
}
int pack_few_out(const float* src, int cout, int cin, int taps, float* dst, cudaStream_t s) {
  k_pack_few_out<<<cdiv(cout * cin * taps, 256), 256, 0, s>>>(src, cout, cin, taps, dst);
  KD_LAUNCH_CHECK();
  return 0;
}

__global__ void k_bn_fold(const float* g, const float* beta, const float* mean, const float* var, const float* bias, int n,
                          float eps, float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sc = g[i] / sqrtf(var[i] + eps);
  scale[i] = sc;
  shift[i] = beta[i] + ((bias ? bias[i] : 0.f) - mean[i]) * sc;
}
int bn_fold(const float* g, const float* beta, const float* mean, const float* var, const float* bias, int n, float eps,
            float* scale, float* shift, cudaStream_t s) {
  k_bn_fold<<<cdiv(n, 128), 128, 0, s>>>(g, beta, mean, var, bias, n, eps, scale, shift);
  KD_LAUNCH_CHECK();
  return 0;
}

__global__ void k_copy_f32(const float* src, float* dst, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
int copy_f32(const float* src, float* dst, long n, cudaStream_t s) {
  if (n <= 0) return 0;
  k_copy_f32<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n);
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
