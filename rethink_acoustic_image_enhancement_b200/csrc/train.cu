// First slice of the training step: forward-with-saves and backward of the GDFN half of a TransformerBlock,
//     out = x + project_out( gelu(u[:h]) * u[h:] ),   u = dwconv3x3(project_in(LayerNorm(x)))          (KDLAE_model.py:50-52,101-106,163)
// in the fp32 reference-grade path: NHWC fp32 activations, CUDA-core kernels, every sum in a fixed order (deterministic).
// The 1x1 convs (forward and dgrad) reuse conv_gemm_simt, the depthwise conv and its dgrad reuse dwconv3x3 (dgrad = the same conv
// with the taps reversed); new here: wgrad of the 1x1 convs (a pixel-reduction GEMM), wgrad of the depthwise conv, the GELU-gate
// backward and the BiasFree LayerNorm backward with the residual add and the gamma gradient.  Weight layouts are the packed ones of
// the forward path (project_in [2hp][C] with the two chunk(2) halves padded to hp, depthwise [9][2hp], project_out [C][hp]); the
// Python wrapper (training.py) maps them from / to the reference's parameter shapes.
// The end of the file holds the fused clip-norm + AdamW step over flat buffers.  tcgen05 dgrad / wgrad for the bf16 path and the
// MDTA half of the block are the next steps of this row.
#include <algorithm>
#include <atomic>
#include "ops.cuh"

namespace kd {

namespace {
std::atomic<int> g_train_tf32{-1};
}
void set_train_matmul_tf32(int on) { g_train_tf32.store(on ? 1 : 0); }
int train_matmul_tf32() {
  int v = g_train_tf32.load();
  if (v < 0) { const char* e = getenv("KDLAE_TRAIN_TF32"); v = (e && e[0] == '1') ? 1 : 0; g_train_tf32.store(v); }
  return v;
}

namespace {
// every fp32 GEMM of the training step goes through here: tcgen05 kind::tf32 when switched on and the shape allows, else CUDA cores
int train_gram(const float* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s) {
  if (train_matmul_tf32() && gram_tf32_eligible(qk, ld, C, heads)) return gram_tf32(qk, ld, nimg, HW, C, heads, splits, part, s);
  return mdta_gram<float>(qk, ld, nimg, HW, C, heads, splits, part, s);
}
int train_gemm(const ConvOp& g, cudaStream_t s) {
  if (train_matmul_tf32() && gemm_tf32_eligible(g)) return gemm_tf32(g, s);
  return conv_gemm_simt<float>(g, s);
}

constexpr int WG_T = 64, WG_P = 32;      // wgrad tile: 64 x 64 outputs, 32 pixels per step

// part[split][n][k] = sum_{p in split} A[p][n] * B[p][k]     (A: [P][lda], B: [P][ldb])
__global__ void __launch_bounds__(256) k_wgrad_part(const float* __restrict__ A, long lda, int N, const float* __restrict__ B, long ldb,
                                                    int K, long P, int per, float* __restrict__ part) {
  __shared__ float As[WG_P][WG_T + 4];
  __shared__ float Bs[WG_P][WG_T + 4];
  const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T, split = blockIdx.z;
  const long p_begin = (long)split * per, p_end = min(P, p_begin + per);
  const int tid = threadIdx.x, tn = (tid >> 4) * 4, tk = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long p0 = p_begin; p0 < p_end; p0 += WG_P) {
    for (int e = tid; e < WG_P * WG_T; e += 256) {
      const int pp = e / WG_T, c = e % WG_T;
      const long p = p0 + pp;
      As[pp][c] = (p < p_end && n0 + c < N) ? A[p * lda + n0 + c] : 0.f;
      Bs[pp][c] = (p < p_end && k0 + c < K) ? B[p * ldb + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < WG_P; ++pp) {
      const float4 a = *reinterpret_cast<const float4*>(&As[pp][tn]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[pp][tk]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + tn + i < N && k0 + tk + j < K) part[((long)split * N + n0 + tn + i) * K + k0 + tk + j] = acc[i][j];
}
// out[e] = sum_split part[split][e]  (fixed order)
__global__ void k_sum_parts(const float* __restrict__ part, int splits, long n, float* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[(long)sp * n + e];
  out[e] = s;
}

// depthwise wgrad: part[split][tap][c] = sum_{p in split} du[p][c] * t[p shifted by tap][c]   (zero padding)
__global__ void __launch_bounds__(128) k_dw_wgrad_part(const float* __restrict__ du, const float* __restrict__ t, int nimg, int H, int W,
                                                       int C, int per, float* __restrict__ part) {
  const int c = blockIdx.x * 128 + threadIdx.x, split = blockIdx.y;
  if (c >= C) return;
  const long P = (long)nimg * H * W, p_begin = (long)split * per, p_end = min(P, p_begin + per);
  float acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.f;
  int x = (int)(p_begin % W), y = (int)((p_begin / W) % H);          // advanced incrementally: no division per pixel
  for (long p = p_begin; p < p_end; ++p) {
    const float g = du[p * C + c];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc[k] = fmaf(g, t[(p + (long)(k / 3 - 1) * W + (k % 3 - 1)) * C + c], acc[k]);
    }
    if (++x == W) { x = 0; if (++y == H) y = 0; }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) part[((long)split * 9 + k) * C + c] = acc[k];
}

// y = x * rstd * gamma (BiasFree LayerNorm, x not centred)
__global__ void k_ln_apply(const float* __restrict__ x, const float* __restrict__ rstd, const float* __restrict__ gamma, int C, long n,
                           float* __restrict__ y) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  y[e] = x[e] * rstd[e / C] * gamma[e % C];
}
// g = gelu(u1) * u2
__global__ void k_gate_fwd(const float* __restrict__ u, int hp, long P, float* __restrict__ g) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * hp) return;
  const long p = e / hp; const int j = (int)(e % hp);
  g[e] = gelu_erf(u[p * 2 * hp + j]) * u[p * 2 * hp + hp + j];
}
// du1 = dg * u2 * gelu'(u1), du2 = dg * gelu(u1);  gelu'(x) = Phi(x) + x * phi(x)
__global__ void k_gate_bwd(const float* __restrict__ u, const float* __restrict__ dg, int hp, long P, float* __restrict__ du) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * hp) return;
  const long p = e / hp; const int j = (int)(e % hp);
  const float u1 = u[p * 2 * hp + j], u2 = u[p * 2 * hp + hp + j], d = dg[e];
  const float Phi = 0.5f * (1.0f + erff(u1 * 0.70710678118654752440f));
  const float phi = 0.39894228040143267794f * expf(-0.5f * u1 * u1);
  du[p * 2 * hp + j] = d * u2 * (Phi + u1 * phi);
  du[p * 2 * hp + hp + j] = d * (u1 * Phi);
}
// dx = dout + LN_backward(dy);  part_gamma[block][c] = sum over the block's pixels of dy_c * x_c * rstd
// one warp per pixel; C <= 1024
__global__ void __launch_bounds__(256) k_ln_bwd(const float* __restrict__ x, const float* __restrict__ rstd, const float* __restrict__ mu,
                                                const float* __restrict__ gamma, const float* __restrict__ dy,
                                                const float* __restrict__ dout, int C, long P, int pix_per_block,
                                                float* __restrict__ dx, float* __restrict__ part_gamma) {
  extern __shared__ float sg[];           // [8 warps][C] gamma-gradient partials
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = lane; c < C; c += 32) sg[warp * C + c] = 0.f;
  const long p_begin = (long)blockIdx.x * pix_per_block, p_end = min(P, p_begin + pix_per_block);
  for (long p = p_begin + warp; p < p_end; p += 8) {
    const float r = rstd[p], m = mu[p];
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(dy[p * C + c] * gamma[c], x[p * C + c], s);
    s = warp_sum(s);
    const float k = -r * r * r * s / (float)C;
    for (int c = lane; c < C; c += 32) {
      const float xv = x[p * C + c], d = dy[p * C + c];
      dx[p * C + c] = dout[p * C + c] + r * gamma[c] * d + k * (xv - m);
      sg[warp * C + c] = fmaf(d * xv, r, sg[warp * C + c]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sg[w * C + c];
    part_gamma[(long)blockIdx.x * C + c] = t;
  }
}
__global__ void k_transpose(const float* __restrict__ a, int R, int Cc, float* __restrict__ out) {   // out[c][r] = a[r][c]
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)R * Cc) return;
  const int r = (int)(e / Cc), c = (int)(e % Cc);
  out[(long)c * R + r] = a[e];
}
__global__ void k_flip9(const float* __restrict__ w9c, int C, float* __restrict__ out) {              // out[t][c] = w9c[8 - t][c]
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 9 * C) return;
  out[e] = w9c[(8 - e / C) * C + e % C];
}

int wgrad_1x1(const float* A, long lda, int N, const float* B, long ldb, int K, long P, float* out, float* part, int splits,
              cudaStream_t s) {
  if (train_matmul_tf32() && wgrad_tf32_eligible(A, lda, N, B, ldb, K)) {     // tcgen05 kind::tf32, MN-major operands (gemm_tf32.cu)
    int used = splits;
    KD_TRY(wgrad_tf32(A, lda, N, B, ldb, K, P, part, splits, &used, s));
    k_sum_parts<<<cdiv((long)N * K, 256), 256, 0, s>>>(part, used, (long)N * K, out);
    count_launch();
    KD_LAUNCH_CHECK();
    return 0;
  }
  const int per = (int)((P + splits - 1) / splits);
  ProfScope prof(PC_GEMM_SIMT, s, 2.0 * P * N * K, 4.0 * (double)P * (N + K));
  k_wgrad_part<<<dim3(cdiv(N, WG_T), cdiv(K, WG_T), splits), 256, 0, s>>>(A, lda, N, B, ldb, K, P, per, part);
  count_launch();
  KD_LAUNCH_CHECK();
  k_sum_parts<<<cdiv((long)N * K, 256), 256, 0, s>>>(part, splits, (long)N * K, out);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

// The depthwise wgrad is a per-channel reduction over all pixels: a thread owns a channel, so the only parallelism besides the
// channels is the pixel split - up to 512 splits of >= 32 pixels (with the 32 splits of the dense wgrads a block-backward spent a
// third of its time here, one warp per SM walking thousands of pixels).
constexpr int DW_SPLITS = 512;
int dw_wgrad_splits(long P) { return (int)std::max<long>(1, std::min<long>(DW_SPLITS, P / 32)); }
int dw_wgrad(const float* du, const float* t, int nimg, int H, int W, int C2, float* dw, float* part, cudaStream_t s) {
  const long P = (long)nimg * H * W;
  const int splits = dw_wgrad_splits(P), per = (int)((P + splits - 1) / splits);
  k_dw_wgrad_part<<<dim3(cdiv(C2, 128), splits), 128, 0, s>>>(du, t, nimg, H, W, C2, per, part);
  count_launch();
  KD_LAUNCH_CHECK();
  k_sum_parts<<<cdiv(9L * C2, 256), 256, 0, s>>>(part, splits, 9L * C2, dw);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

int conv1x1_f32(const float* a, int C, const float* w, int N, const float* res, float* out, int nimg, int H, int W, cudaStream_t s) {
  ConvOp g;
  g.a0 = a; g.c0 = C; g.ld0 = C; g.nimg = nimg; g.H = H; g.W = W; g.w = w; g.w_ld = C; g.w_tap_ld = C;
  g.epi.res = res; g.epi.res_ld = N; g.epi.out = out; g.epi.out_ld = N; g.epi.N = N; g.epi.H = H; g.epi.W = W;
  return train_gemm(g, s);
}

struct GdfnWs {
  float *rstd, *mu, *y, *t, *u, *g;                 // saved by the forward
  float *dg, *du, *dt, *dy, *wt, *w9f, *part;       // backward scratch
  size_t total;
};
constexpr int GD_SPLITS = 32, GD_LN_PIX = 64;     // 64 pixels per LayerNorm-backward block: 256 left 128 blocks for a 2 x 128^2 batch

GdfnWs gdfn_layout(float* base, int nimg, int H, int W, int C, int hp) {
  const size_t P = (size_t)nimg * H * W;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
  GdfnWs L;
  L.rstd = take(P); L.mu = take(P); L.y = take(P * C); L.t = take(P * 2 * hp); L.u = take(P * 2 * hp); L.g = take(P * hp);
  L.dg = take(P * hp); L.du = take(P * 2 * hp); L.dt = take(P * 2 * hp); L.dy = take(P * C);
  L.wt = take((size_t)2 * hp * C); L.w9f = take((size_t)9 * 2 * hp);
  const size_t ln_blocks = (P + GD_LN_PIX - 1) / GD_LN_PIX;
  L.part = take(std::max<size_t>({(size_t)GD_SPLITS * 2 * hp * C, (size_t)dw_wgrad_splits((long)P) * 9 * 2 * hp, ln_blocks * C}));
  L.total = off;
  return L;
}

}  // namespace

size_t gdfn_train_ws_floats(int nimg, int H, int W, int C, int hp) { return gdfn_layout(nullptr, nimg, H, W, C, hp).total; }

int gdfn_forward_train(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, float* out, int nimg,
                       int H, int W, int C, int hp, float* ws, cudaStream_t s) {
  KD_CHECK(C % 8 == 0 && hp % 8 == 0 && C <= 1024, "gdfn_forward_train: C=%d hp=%d", C, hp);
  const GdfnWs L = gdfn_layout(ws, nimg, H, W, C, hp);
  const long P = (long)nimg * H * W;
  KD_TRY(ln_stats<float>(x, C, C, P, L.rstd, L.mu, s));
  k_ln_apply<<<cdiv(P * C, 256), 256, 0, s>>>(x, L.rstd, gamma, C, P * C, L.y);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(conv1x1_f32(L.y, C, w_in, 2 * hp, nullptr, L.t, nimg, H, W, s));
  KD_TRY(dwconv3x3<float>(L.t, 2 * hp, L.u, 2 * hp, w_dw, nullptr, nimg, H, W, 2 * hp, 0, s));
  k_gate_fwd<<<cdiv(P * hp, 256), 256, 0, s>>>(L.u, hp, P, L.g);
  count_launch();
  KD_LAUNCH_CHECK();
  return conv1x1_f32(L.g, hp, w_out, C, x, out, nimg, H, W, s);
}

// dout [P][C] -> dx [P][C], dgamma [C], dw_in [2hp][C], dw_dw [9][2hp], dw_out [C][hp]; ws as left by gdfn_forward_train
int gdfn_backward(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, const float* dout,
                  float* dx, float* dgamma, float* dw_in, float* dw_dw, float* dw_out, int nimg, int H, int W, int C, int hp, float* ws,
                  cudaStream_t s) {
  const GdfnWs L = gdfn_layout(ws, nimg, H, W, C, hp);
  const long P = (long)nimg * H * W;
  const int splits = (int)std::min<long>(GD_SPLITS, std::max<long>(1, P / 256));
  // project_out: wgrad and dgrad
  KD_TRY(wgrad_1x1(dout, C, C, L.g, hp, hp, P, dw_out, L.part, splits, s));
  k_transpose<<<cdiv((long)C * hp, 256), 256, 0, s>>>(w_out, C, hp, L.wt);                 // [hp][C]
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(conv1x1_f32(dout, C, L.wt, hp, nullptr, L.dg, nimg, H, W, s));
  // gate
  k_gate_bwd<<<cdiv(P * hp, 256), 256, 0, s>>>(L.u, L.dg, hp, P, L.du);
  count_launch();
  KD_LAUNCH_CHECK();
  // depthwise conv: wgrad, dgrad (= the conv with reversed taps)
  KD_TRY(dw_wgrad(L.du, L.t, nimg, H, W, 2 * hp, dw_dw, L.part, s));
  k_flip9<<<cdiv(9 * 2 * hp, 256), 256, 0, s>>>(w_dw, 2 * hp, L.w9f);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(dwconv3x3<float>(L.du, 2 * hp, L.dt, 2 * hp, L.w9f, nullptr, nimg, H, W, 2 * hp, 0, s));
  // project_in: wgrad and dgrad
  KD_TRY(wgrad_1x1(L.dt, 2 * hp, 2 * hp, L.y, C, C, P, dw_in, L.part, splits, s));
  k_transpose<<<cdiv((long)2 * hp * C, 256), 256, 0, s>>>(w_in, 2 * hp, C, L.wt);           // [C][2hp]
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(conv1x1_f32(L.dt, 2 * hp, L.wt, C, nullptr, L.dy, nimg, H, W, s));
  // LayerNorm backward + residual + gamma gradient
  const int ln_blocks = (int)cdiv(P, GD_LN_PIX);
  k_ln_bwd<<<ln_blocks, 256, sizeof(float) * 8 * C, s>>>(x, L.rstd, L.mu, gamma, L.dy, dout, C, P, GD_LN_PIX, dx, L.part);
  count_launch();
  KD_LAUNCH_CHECK();
  k_sum_parts<<<cdiv(C, 256), 256, 0, s>>>(L.part, ln_blocks, C, dgamma);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd

// ---- fused gradient clipping + AdamW over flat fp32 buffers (image_restoration_model.py:218-220: clip_grad_norm_(params, 0.01),
// then the AdamW step of KDLAET.yml) ----------------------------------------------------------------------------------------------
// Parameters, gradients and both moments live in flat buffers (the same flat gradient buffer the bucketed NCCL all-reduce
// works on), so the step is two passes: a deterministic sum of squares and one fused update that applies the clip
// coefficient min(1, max_norm / (||g|| + 1e-6)) on the fly - the clipped gradient is never written back.
namespace kd {
namespace {
__global__ void __launch_bounds__(256) k_sumsq_part(const float* __restrict__ g, long n, double* __restrict__ part) {
  __shared__ double sh[8];
  double s = 0.0;
  for (long i0 = (long)blockIdx.x * 2048; i0 < n; i0 += (long)gridDim.x * 2048) {
    float t = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const long i = i0 + u * 256 + threadIdx.x;
      if (i < n) { const float v = g[i]; t = fmaf(v, v, t); }
    }
    s += (double)t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    part[blockIdx.x] = t;
  }
}
__global__ void k_sumsq_final(const double* __restrict__ part, int nblk, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nblk; ++i) t += part[i];
  *out = t;
}
__global__ void __launch_bounds__(256) k_adamw(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                               float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float wd,
                                               float bc1, float bc2_sqrt, float max_norm, const double* __restrict__ norm_sq) {
  float coef = 1.0f;
  if (norm_sq != nullptr && max_norm > 0.f) {
    const float total = (float)sqrt(*norm_sq);
    coef = fminf(1.0f, max_norm / (total + 1e-6f));
  }
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * coef;
  float pi = p[i] * (1.0f - lr * wd);                      // decoupled weight decay
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}
}  // namespace

int grad_norm_sq(const float* g, long n, double* out, double* scratch /* 1024 doubles */, cudaStream_t s) {
  KD_CHECK(g && out && scratch && n > 0, "grad_norm_sq: bad argument");
  const int nblk = (int)std::min<long>(1024, (n + 2047) / 2048);
  ProfScope prof(PC_POOL_RESAMPLE, s, 2.0 * n, 4.0 * n);
  k_sumsq_part<<<nblk, 256, 0, s>>>(g, n, scratch);
  count_launch();
  KD_LAUNCH_CHECK();
  k_sumsq_final<<<1, 32, 0, s>>>(scratch, nblk, out);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

int adamw_step(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps, float wd, int step,
               float max_norm, const double* norm_sq, cudaStream_t s) {
  KD_CHECK(p && g && m && v && n > 0 && step >= 1, "adamw_step: bad argument");
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  ProfScope prof(PC_POOL_RESAMPLE, s, 12.0 * n, 28.0 * n);
  k_adamw<<<cdiv(n, 256), 256, 0, s>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, sqrtf(bc2), max_norm, norm_sq);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
}  // namespace kd

// ---- MDTA half of a TransformerBlock (KDLAE_model.py:124-145,162): out = x + project_out(softmax(q^ k^T * temp) v) ----------------
//   q, k, v = chunk(dwconv3x3(qkv(LayerNorm(x))));  q^ = q / max(|q|, 1e-12) over the pixels of each (image, head, channel)
// Forward keeps what the backward needs (y, t, u, o, the summed Gram / squared norms and the softmax).  The backward never forms a
// per-pixel normalised q^, k^: with G^ = q^ k^T and dG^ = temp * dS,
//   dq_i = sum_j dG^_ij / (|q_i| |k_j|) k_j - (sum_j dG^_ij G^_ij) / |q_i|^2 q_i        (and symmetrically for k),
// so d(q, k) is ONE per-image 2C x 2C pixel-wise linear map of (q, k) - a grouped 1x1 GEMM - whose matrix a small kernel builds
// from the ch x ch quantities.  dA = do v^T is the same pixel reduction as the forward Gram (mdta_gram on [do | v]).
namespace kd {
namespace {

__global__ void k_sum_groups(const float* __restrict__ part, int splits, long psz, long total, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long g = i / psz, e = i % psz;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[(g * splits + sp) * psz + e];
  out[i] = s;
}
// grid (heads, nimg), one warp per softmax row: A = softmax(temp * G / (nq nk)); Ab / AbT = blockdiag(A) / its transpose [nimg][C][C]
__global__ void __launch_bounds__(256) k_attn_softmax_train(const float* __restrict__ gs, int C, int heads, const float* __restrict__ temp,
                                                            float* __restrict__ A, float* __restrict__ Ab, float* __restrict__ AbT) {
  const int ch = C / heads, head = blockIdx.x, img = blockIdx.y;
  const long psz = (long)ch * ch + 2 * ch;
  const float* g = gs + ((long)img * heads + head) * psz;
  float* a = A + ((long)img * heads + head) * ch * ch;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float tp = temp[head];
  for (int i = warp; i < ch; i += 8) {
    const float nq = fmaxf(sqrtf(g[ch * ch + i]), 1e-12f);
    float mx = -INFINITY;
    for (int j = lane; j < ch; j += 32) {
      const float l = g[i * ch + j] / (nq * fmaxf(sqrtf(g[ch * ch + ch + j]), 1e-12f)) * tp;
      a[i * ch + j] = l;
      mx = fmaxf(mx, l);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < ch; j += 32) { const float e = expf(a[i * ch + j] - mx); a[i * ch + j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    const int n = head * ch + i;
    __syncwarp();                    // the loop below reads a[] entries written by other lanes
    for (int c = lane; c < C; c += 32) {
      const int j = c - head * ch;
      float v = 0.f;
      if (j >= 0 && j < ch) { v = a[i * ch + j] * inv; a[i * ch + j] = v; }
      Ab[((long)img * C + n) * C + c] = v;
      AbT[((long)img * C + c) * C + n] = v;
    }
  }
  // AbT rows outside this head's columns are written by the other heads' blocks (each block writes column range n of all rows c)
}
__global__ void k_copy_cols(const float* __restrict__ src, long lds, float* __restrict__ dst, long ldd, int C, long P) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P * C) return;
  dst[(e / C) * ldd + e % C] = src[(e / C) * lds + e % C];
}
// grid (heads, nimg): softmax / temperature / normalisation backward on the ch x ch quantities; builds the per-image [2C][2C] matrix of
// the d(q, k) map and the per-(image, head) temperature gradient.  dynamic smem: ch*ch floats (dG^) + 2*ch.
__global__ void __launch_bounds__(256) k_attn_bwd_small(const float* __restrict__ gs, const float* __restrict__ A,
                                                        const float* __restrict__ dAs, int C, int heads, const float* __restrict__ temp,
                                                        float* __restrict__ Wqk, float* __restrict__ dtemp_part) {
  extern __shared__ float sm[];
  const int ch = C / heads, head = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
  const long psz = (long)ch * ch + 2 * ch;
  const float* g = gs + ((long)img * heads + head) * psz;
  const float* a = A + ((long)img * heads + head) * ch * ch;
  const float* da = dAs + ((long)img * heads + head) * psz;      // first ch*ch entries: dA = do v^T
  float* dG = sm;                 // [ch][ch]  temp * dS
  float* sq = sm + ch * ch;       // [ch] s_i
  float* rk = sq + ch;            // [ch] r_j
  __shared__ float red[8];
  const float tp = temp[head];
  const int warp = tid >> 5, lane = tid & 31;
  float dt_acc = 0.f;
  for (int i = warp; i < ch; i += 8) {
    float rd = 0.f;
    for (int j = lane; j < ch; j += 32) rd = fmaf(a[i * ch + j], da[i * ch + j], rd);
    rd = warp_sum(rd);
    const float nq = fmaxf(sqrtf(g[ch * ch + i]), 1e-12f);
    float si = 0.f;
    for (int j = lane; j < ch; j += 32) {
      const float nk = fmaxf(sqrtf(g[ch * ch + ch + j]), 1e-12f);
      const float dS = a[i * ch + j] * (da[i * ch + j] - rd);
      const float gh = g[i * ch + j] / (nq * nk);
      dt_acc = fmaf(dS, gh, dt_acc);
      const float dg = tp * dS;
      dG[i * ch + j] = dg;
      si = fmaf(dg, gh, si);
    }
    si = warp_sum(si);
    if (lane == 0) sq[i] = si;
  }
  dt_acc = warp_sum(dt_acc);
  if (lane == 0) red[warp] = dt_acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    dtemp_part[(long)img * heads + head] = t;
  }
  for (int j = tid; j < ch; j += 256) {
    const float nk = fmaxf(sqrtf(g[ch * ch + ch + j]), 1e-12f);
    float r = 0.f;
    for (int i = 0; i < ch; ++i) r = fmaf(dG[i * ch + j], g[i * ch + j] / (fmaxf(sqrtf(g[ch * ch + i]), 1e-12f) * nk), r);
    rk[j] = r;
  }
  __syncthreads();
  // rows/cols of this head inside the [2C][2C] matrix (the buffer is zeroed beforehand)
  float* Wm = Wqk + (long)img * 4 * C * C;
  const int C2 = 2 * C, q0 = head * ch, k0 = C + head * ch;
  for (int e = tid; e < ch * ch; e += 256) {
    const int i = e / ch, j = e % ch;
    const float nq = fmaxf(sqrtf(g[ch * ch + i]), 1e-12f), nk = fmaxf(sqrtf(g[ch * ch + ch + j]), 1e-12f);
    const float m = dG[e] / (nq * nk);
    Wm[(long)(q0 + i) * C2 + k0 + j] = m;     // dq_i += m * k_j
    Wm[(long)(k0 + j) * C2 + q0 + i] = m;     // dk_j += m * q_i
  }
  for (int i = tid; i < ch; i += 256) {
    const float nq2 = fmaxf(g[ch * ch + i], 1e-24f), nk2 = fmaxf(g[ch * ch + ch + i], 1e-24f);
    Wm[(long)(q0 + i) * C2 + q0 + i] = -sq[i] / nq2;
    Wm[(long)(k0 + i) * C2 + k0 + i] = -rk[i] / nk2;
  }
}
__global__ void k_sum_dtemp(const float* __restrict__ part, int nimg, int heads, float* __restrict__ out) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= heads) return;
  float s = 0.f;
  for (int i = 0; i < nimg; ++i) s += part[(long)i * heads + h];
  out[h] = s;
}
__global__ void k_zero(float* __restrict__ p, long n) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) p[e] = 0.f;
}

struct MdtaWs {
  float *rstd, *mu, *y, *t, *u, *o, *gs, *A, *Ab, *AbT;          // saved by the forward
  float *dov, *dAs, *Wqk, *du, *dt, *dy, *wt, *w9f, *dtp, *part;  // backward scratch
  size_t total;
};
// Pixel splits of the fp32 Gram reductions of the training step: 512 pixels per split (the inference rule, 2048, leaves a batch of
// two 128 x 128 crops on 16-32 CTAs); depends on the image size only, like the inference rule.
int train_gram_splits(int HW, int /*nimg_heads*/) {
  const int s = HW / 512;
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}

MdtaWs mdta_layout(float* base, int nimg, int H, int W, int C, int heads) {
  const size_t P = (size_t)nimg * H * W, ch = C / heads, psz = ch * ch + 2 * ch;
  const int splits = train_gram_splits(H * W, nimg * heads);
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) / 64 * 64; return p; };
  MdtaWs L;
  L.rstd = take(P); L.mu = take(P); L.y = take(P * C); L.t = take(P * 3 * C); L.u = take(P * 3 * C); L.o = take(P * C);
  L.gs = take((size_t)nimg * heads * psz); L.A = take((size_t)nimg * heads * ch * ch);
  L.Ab = take((size_t)nimg * C * C); L.AbT = take((size_t)nimg * C * C);
  L.dov = take(P * 2 * C); L.dAs = take((size_t)nimg * heads * psz); L.Wqk = take((size_t)nimg * 4 * C * C);
  L.du = take(P * 3 * C); L.dt = take(P * 3 * C); L.dy = take(P * C);
  L.wt = take((size_t)3 * C * C); L.w9f = take((size_t)9 * 3 * C); L.dtp = take((size_t)nimg * heads);
  const size_t ln_blocks = (P + GD_LN_PIX - 1) / GD_LN_PIX;
  L.part = take(std::max<size_t>({(size_t)GD_SPLITS * 3 * C * C, (size_t)dw_wgrad_splits((long)P) * 9 * 3 * C, ln_blocks * C,
                                  (size_t)nimg * heads * splits * psz}));
  L.total = off;
  return L;
}
int grouped_1x1_f32(const float* a, int Cin, long lda, const float* w, int N, long w_group_stride, float* out, long ldo, int coff,
                    int nimg, int H, int W, cudaStream_t s) {
  ConvOp g;
  g.a0 = a; g.c0 = Cin; g.ld0 = lda; g.nimg = nimg; g.H = H; g.W = W; g.w = w; g.w_ld = Cin; g.w_tap_ld = Cin;
  g.groups = nimg; g.w_group_stride = w_group_stride;
  g.epi.out = out; g.epi.out_ld = ldo; g.epi.out_coff = coff; g.epi.N = N; g.epi.H = H; g.epi.W = W;
  return train_gemm(g, s);
}
}  // namespace

size_t mdta_train_ws_floats(int nimg, int H, int W, int C, int heads) { return mdta_layout(nullptr, nimg, H, W, C, heads).total; }

int mdta_forward_train(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj, const float* temp,
                       float* out, int nimg, int H, int W, int C, int heads, float* ws, cudaStream_t s) {
  const int ch = heads > 0 ? C / heads : 0;
  KD_CHECK(heads > 0 && C % heads == 0 && ch % 8 == 0 && ch <= 96 && C <= 1024, "mdta_forward_train: C=%d heads=%d (channels per head <= 96)", C, heads);
  const MdtaWs L = mdta_layout(ws, nimg, H, W, C, heads);
  const long P = (long)nimg * H * W;
  const int HW = H * W, splits = train_gram_splits(HW, nimg * heads);
  const long psz = (long)ch * ch + 2 * ch;
  KD_TRY(ln_stats<float>(x, C, C, P, L.rstd, L.mu, s));
  k_ln_apply<<<cdiv(P * C, 256), 256, 0, s>>>(x, L.rstd, gamma, C, P * C, L.y);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(conv1x1_f32(L.y, C, w_qkv, 3 * C, nullptr, L.t, nimg, H, W, s));
  KD_TRY(dwconv3x3<float>(L.t, 3 * C, L.u, 3 * C, w_dw, nullptr, nimg, H, W, 3 * C, 0, s));
  KD_TRY(train_gram(L.u, 3 * C, nimg, HW, C, heads, splits, L.part, s));
  k_sum_groups<<<cdiv((long)nimg * heads * psz, 256), 256, 0, s>>>(L.part, splits, psz, (long)nimg * heads * psz, L.gs);
  count_launch();
  KD_LAUNCH_CHECK();
  k_attn_softmax_train<<<dim3(heads, nimg), 256, 0, s>>>(L.gs, C, heads, temp, L.A, L.Ab, L.AbT);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(grouped_1x1_f32(L.u + 2 * C, C, 3 * C, L.Ab, C, (long)C * C, L.o, C, 0, nimg, H, W, s));        // o = attn @ v
  return conv1x1_f32(L.o, C, w_proj, C, x, out, nimg, H, W, s);                                          // + residual
}

int mdta_backward(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj, const float* temp,
                  const float* dout, float* dx, float* dgamma, float* dw_qkv, float* dw_dw, float* dw_proj, float* dtemp, int nimg, int H,
                  int W, int C, int heads, float* ws, cudaStream_t s) {
  const int ch = C / heads;
  const MdtaWs L = mdta_layout(ws, nimg, H, W, C, heads);
  const long P = (long)nimg * H * W;
  const int HW = H * W, gsplits = train_gram_splits(HW, nimg * heads);
  const int splits = (int)std::min<long>(GD_SPLITS, std::max<long>(1, P / 256));
  const long psz = (long)ch * ch + 2 * ch;
  // project_out: wgrad, dgrad (do -> columns [0, C) of dov)
  KD_TRY(wgrad_1x1(dout, C, C, L.o, C, C, P, dw_proj, L.part, splits, s));
  k_transpose<<<cdiv((long)C * C, 256), 256, 0, s>>>(w_proj, C, C, L.wt);
  count_launch();
  KD_LAUNCH_CHECK();
  {
    ConvOp g;
    g.a0 = dout; g.c0 = C; g.ld0 = C; g.nimg = nimg; g.H = H; g.W = W; g.w = L.wt; g.w_ld = C; g.w_tap_ld = C;
    g.epi.out = L.dov; g.epi.out_ld = 2 * C; g.epi.N = C; g.epi.H = H; g.epi.W = W;
    KD_TRY(train_gemm(g, s));
  }
  k_copy_cols<<<cdiv(P * C, 256), 256, 0, s>>>(L.u + 2 * C, 3 * C, L.dov + C, 2 * C, C, P);              // v -> columns [C, 2C)
  count_launch();
  KD_LAUNCH_CHECK();
  // dA = do v^T (pixel reduction), dv = A^T do
  KD_TRY(train_gram(L.dov, 2 * C, nimg, HW, C, heads, gsplits, L.part, s));
  k_sum_groups<<<cdiv((long)nimg * heads * psz, 256), 256, 0, s>>>(L.part, gsplits, psz, (long)nimg * heads * psz, L.dAs);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(grouped_1x1_f32(L.dov, C, 2 * C, L.AbT, C, (long)C * C, L.du, 3 * C, 2 * C, nimg, H, W, s));
  // softmax / temperature / normalisation backward -> per-image d(q, k) matrix, then one grouped GEMM on (q, k)
  k_zero<<<cdiv((long)nimg * 4 * C * C, 256), 256, 0, s>>>(L.Wqk, (long)nimg * 4 * C * C);
  count_launch();
  KD_LAUNCH_CHECK();
  k_attn_bwd_small<<<dim3(heads, nimg), 256, sizeof(float) * (ch * ch + 2 * ch), s>>>(L.gs, L.A, L.dAs, C, heads, temp, L.Wqk, L.dtp);
  count_launch();
  KD_LAUNCH_CHECK();
  k_sum_dtemp<<<cdiv(heads, 32), 32, 0, s>>>(L.dtp, nimg, heads, dtemp);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(grouped_1x1_f32(L.u, 2 * C, 3 * C, L.Wqk, 2 * C, 4L * C * C, L.du, 3 * C, 0, nimg, H, W, s));
  // depthwise conv: wgrad, dgrad
  KD_TRY(dw_wgrad(L.du, L.t, nimg, H, W, 3 * C, dw_dw, L.part, s));
  k_flip9<<<cdiv(27 * C, 256), 256, 0, s>>>(w_dw, 3 * C, L.w9f);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(dwconv3x3<float>(L.du, 3 * C, L.dt, 3 * C, L.w9f, nullptr, nimg, H, W, 3 * C, 0, s));
  // qkv 1x1: wgrad, dgrad
  KD_TRY(wgrad_1x1(L.dt, 3 * C, 3 * C, L.y, C, C, P, dw_qkv, L.part, splits, s));
  k_transpose<<<cdiv(3L * C * C, 256), 256, 0, s>>>(w_qkv, 3 * C, C, L.wt);
  count_launch();
  KD_LAUNCH_CHECK();
  KD_TRY(conv1x1_f32(L.dt, 3 * C, L.wt, C, nullptr, L.dy, nimg, H, W, s));
  const int ln_blocks = (int)cdiv(P, GD_LN_PIX);
  k_ln_bwd<<<ln_blocks, 256, sizeof(float) * 8 * C, s>>>(x, L.rstd, L.mu, gamma, L.dy, dout, C, P, GD_LN_PIX, dx, L.part);
  count_launch();
  KD_LAUNCH_CHECK();
  k_sum_parts<<<cdiv(C, 256), 256, 0, s>>>(L.part, ln_blocks, C, dgamma);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
}  // namespace kd


// ==========================================================================================================================
// Dense conv (1x1 / 3x3, dilation d, zero padding d * (k / 2), no bias) in training mode, fp32 NHWC: the layers of KDLAE-T
// outside the TransformerBlocks (OverlapPatchEmbed :169-178, Downsample / Upsample bodies :182-200, reduce_chan, output,
// output_param (dilation 2), output2, cen, upen, outputen :239-268).  Weights [Cout][k*k][Cin].
//   forward : out = conv(x, w)                                   (the fp32 implicit-GEMM kernel of the inference path)
//   backward: dx = conv(dout, w^T with the taps mirrored)        (same kernel)
//             dw[n][tap][k] = sum_p dout[p][n] * x[p + off(tap)][k]   (pixel-split partial sums, fixed-order reduction)
// ==========================================================================================================================
namespace kd {
namespace {

// part[split][n][tap][k] = sum_{p in split} dY[p][n] * X[shift_tap(p)][k];  blockIdx.z = split * taps + tap
__global__ void __launch_bounds__(256) k_wgrad_conv_part(const float* __restrict__ dY, int N, const float* __restrict__ X, int K, int H, int W,
                                                         long P, int ks, int dil, int per, float* __restrict__ part) {
  __shared__ float As[WG_P][WG_T + 4];
  __shared__ float Bs[WG_P][WG_T + 4];
  __shared__ long src[WG_P];
  const int taps = ks * ks;
  const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T, split = blockIdx.z / taps, tap = blockIdx.z % taps;
  const int dx = (tap % ks - ks / 2) * dil, dy = (tap / ks - ks / 2) * dil;
  const long p_begin = (long)split * per, p_end = min(P, p_begin + per);
  const int tid = threadIdx.x, tn = (tid >> 4) * 4, tk = (tid & 15) * 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long p0 = p_begin; p0 < p_end; p0 += WG_P) {
    if (tid < WG_P) {                       // source pixel of the tap (or -1: zero padding / past the split)
      const long p = p0 + tid;
      long sp = -1;
      if (p < p_end) {
        const int x = (int)(p % W), y = (int)((p / W) % H);
        if (x + dx >= 0 && x + dx < W && y + dy >= 0 && y + dy < H) sp = p + (long)dy * W + dx;
      }
      src[tid] = sp;
    }
    __syncthreads();
    for (int e = tid; e < WG_P * WG_T; e += 256) {
      const int pp = e / WG_T, c = e % WG_T;
      const long p = p0 + pp, sp = src[pp];
      As[pp][c] = (p < p_end && n0 + c < N) ? dY[p * N + n0 + c] : 0.f;
      Bs[pp][c] = (sp >= 0 && k0 + c < K) ? X[sp * K + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int pp = 0; pp < WG_P; ++pp) {
      const float4 a = *reinterpret_cast<const float4*>(&As[pp][tn]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[pp][tk]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + tn + i < N && k0 + tk + j < K) part[(((long)split * N + n0 + tn + i) * taps + tap) * K + k0 + tk + j] = acc[i][j];
}

// wt[k][taps - 1 - tap][n] = w[n][tap][k]
__global__ void k_conv_wt(const float* __restrict__ w, int N, int taps, int K, float* __restrict__ wt) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)N * taps * K) return;
  const int k = (int)(e % K), tap = (int)((e / K) % taps), n = (int)(e / ((long)K * taps));
  wt[((long)k * taps + (taps - 1 - tap)) * N + n] = w[e];
}

int conv_splits(long P) { return (int)std::max<long>(1, std::min<long>(32, P / 256)); }

int conv_f32(const float* a, int Cin, const float* w, int Cout, float* out, int nimg, int H, int W, int ks, int dil, cudaStream_t s) {
  ConvOp g;
  g.a0 = a; g.c0 = Cin; g.ld0 = Cin; g.nimg = nimg; g.H = H; g.W = W; g.kh = ks; g.kw = ks; g.dil = dil;
  g.w = w; g.w_ld = (long)ks * ks * Cin; g.w_tap_ld = Cin;
  g.epi.out = out; g.epi.out_ld = Cout; g.epi.N = Cout; g.epi.H = H; g.epi.W = W;
  return train_gemm(g, s);
}

}  // namespace

size_t conv_train_ws_floats(int nimg, int H, int W, int Cin, int Cout, int ks) {
  const long P = (long)nimg * H * W;
  const size_t wsz = (size_t)Cin * ks * ks * Cout;
  return (wsz + 63) / 64 * 64 + (size_t)conv_splits(P) * wsz;
}

int conv_train_forward(const float* x, const float* w, float* out, int nimg, int H, int W, int Cin, int Cout, int ks, int dil,
                       cudaStream_t s) {
  KD_CHECK((ks == 1 || ks == 3) && dil >= 1 && Cin > 0 && Cout > 0, "conv_train_forward: ks=%d dil=%d", ks, dil);
  return conv_f32(x, Cin, w, Cout, out, nimg, H, W, ks, dil, s);
}

int conv_train_backward(const float* x, const float* w, const float* dout, float* dx, float* dw, int nimg, int H, int W, int Cin, int Cout,
                        int ks, int dil, float* ws, cudaStream_t s) {
  KD_CHECK((ks == 1 || ks == 3) && dil >= 1 && Cin > 0 && Cout > 0, "conv_train_backward: ks=%d dil=%d", ks, dil);
  const long P = (long)nimg * H * W;
  const int taps = ks * ks, splits = conv_splits(P);
  const size_t wsz = (size_t)Cin * taps * Cout;
  float* wt = ws;
  float* part = ws + (wsz + 63) / 64 * 64;
  if (dx) {
    k_conv_wt<<<cdiv((long)wsz, 256), 256, 0, s>>>(w, Cout, taps, Cin, wt);
    count_launch();
    KD_LAUNCH_CHECK();
    KD_TRY(conv_f32(dout, Cout, wt, Cin, dx, nimg, H, W, ks, dil, s));
  }
  if (ks == 1) return wgrad_1x1(dout, Cout, Cout, x, Cin, Cin, P, dw, part, splits, s);   // same reduction as the blocks' 1x1 convs
  const int per = (int)((P + splits - 1) / splits);
  KD_CHECK((long)splits * taps <= 65535, "conv_train_backward: grid too large");
  {
    ProfScope prof(PC_GEMM_SIMT, s, 2.0 * P * Cout * Cin * taps, 4.0 * (double)P * (Cout + Cin) * taps);
    k_wgrad_conv_part<<<dim3(cdiv(Cout, WG_T), cdiv(Cin, WG_T), splits * taps), 256, 0, s>>>(dout, Cout, x, Cin, H, W, P, ks, dil, per, part);
    count_launch();
    KD_LAUNCH_CHECK();
  }
  k_sum_parts<<<cdiv((long)wsz, 256), 256, 0, s>>>(part, splits, (long)wsz, dw);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
