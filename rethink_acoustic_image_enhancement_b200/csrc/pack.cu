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

struct BlockDiagArgs { const float* src[4]; const float* nscale[4]; };
template <typename T>
__global__ void __launch_bounds__(256) k_pack_blockdiag(const BlockDiagArgs a, int ng, int c, int taps, T* __restrict__ dst) {
  const int N = ng * c;
  const long total = (long)N * taps * N;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % N), t = (int)((idx / N) % taps), n = (int)(idx / ((long)N * taps));
  const int g = n / c;
  float v = 0.f;
  if (k / c == g) {
    v = a.src[g][((long)(n - g * c) * c + (k - g * c)) * taps + t];
    if (a.nscale[g]) v *= a.nscale[g][n - g * c];
  }
  dst[idx] = from_f<T>(v);
}
template <typename T>
int pack_blockdiag(const float* const* src, const float* const* nscale, int ng, int c, int taps, T* dst, cudaStream_t s) {
  KD_CHECK(ng >= 1 && ng <= 4, "pack_blockdiag: ng=%d", ng);
  BlockDiagArgs a;
  for (int g = 0; g < 4; ++g) { a.src[g] = g < ng ? src[g] : nullptr; a.nscale[g] = (g < ng && nscale) ? nscale[g] : nullptr; }
  const long total = (long)ng * c * taps * ng * c;
  k_pack_blockdiag<T><<<cdiv(total, 256), 256, 0, s>>>(a, ng, c, taps, dst);
  KD_LAUNCH_CHECK();
  return 0;
}
template int pack_blockdiag<float>(const float* const*, const float* const*, int, int, int, float*, cudaStream_t);
template int pack_blockdiag<bf16>(const float* const*, const float* const*, int, int, int, bf16*, cudaStream_t);

// ---- x-packed narrow conv (ConvOp::xpack_cin, conv3x3_tc.cu) -------------------------------------------------------------
// P = 64 / cin pixels form one 64-channel super-pixel.  Output row n = j*cout + co (pixel j of the super-pixel), K column
// k = i*cin + ci (pixel i).  Tile h = td*3 + ty (< 3*kd) is the centre tile of that (frame, row) tap: W[co][ci][h][dx = i - j]
// for |i - j| <= 1.  Halo tiles follow: each holds P/2 (frame, row) taps with two cin-wide slots per tap - slot 2u: the right
// neighbour's pixel 0 feeding output pixel P-1 (dx = +1), slot 2u+1: the left neighbour's pixel P-1 feeding output pixel 0.
int xconv_tiles(int kd, int cin) {
  const int hpt = (64 / cin) / 2;
  return 3 * kd + (3 * kd + hpt - 1) / hpt;
}
__global__ void __launch_bounds__(256) k_pack_xconv(const float* __restrict__ src, const float* __restrict__ nscale, int cout, int cin,
                                                    int kd, bf16* __restrict__ dst) {
  const int P = 64 / cin, taps = kd * 9, nh = 3 * kd, hpt = P / 2;
  const int ntiles = nh + (nh + hpt - 1) / hpt;
  const long total = (long)P * cout * ntiles * 64;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % 64);
  const int tile = (int)((idx / 64) % ntiles);
  const int n = (int)(idx / (64L * ntiles));
  const int j = n / cout, co = n % cout;
  float v = 0.f;
  if (tile < nh) {
    const int i = k / cin, ci = k % cin, dx = i - j;
    if (dx >= -1 && dx <= 1) v = src[((long)co * cin + ci) * taps + tile * 3 + dx + 1];
  } else {
    const int slot = k / cin, ci = k % cin;
    const int h = (tile - nh) * hpt + slot / 2;
    if (h < nh) {
      if ((slot & 1) == 0) { if (j == P - 1) v = src[((long)co * cin + ci) * taps + h * 3 + 2]; }
      else { if (j == 0) v = src[((long)co * cin + ci) * taps + h * 3 + 0]; }
    }
  }
  if (nscale) v *= nscale[co];
  dst[idx] = __float2bfloat16_rn(v);
}
int pack_xconv(const float* src, const float* nscale, int cout, int cin, int kd, bf16* dst, cudaStream_t s) {
  KD_CHECK((cin == 16 || cin == 32) && (kd == 1 || kd == 3) && cout % 8 == 0, "pack_xconv: unsupported cin=%d kd=%d cout=%d", cin, kd, cout);
  const long total = (long)(64 / cin) * cout * xconv_tiles(kd, cin) * 64;
  k_pack_xconv<<<cdiv(total, 256), 256, 0, s>>>(src, nscale, cout, cin, kd, dst);
  KD_LAUNCH_CHECK();
  return 0;
}

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
