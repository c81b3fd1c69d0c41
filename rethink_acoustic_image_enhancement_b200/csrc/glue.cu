// Memory-bound glue kernels (NHWC, vectorised 8 channels per access, fp32 math):
// LayerNorm statistics, depthwise 3x3 (+GELU gate), MDTA Gram/norm reduction and softmax fold,
// tiny-channel direct convolutions, pooling, bilinear up-sampling, GAP+MLP head, layout conversion.
#include <type_traits>
#include "ops.cuh"

namespace kd {

// =====================================================================================
// LayerNorm statistics (KDLAE_model.py:50-52, :67-70): biased variance over channels, eps 1e-5
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256) k_ln_stats(const T* __restrict__ x, long ld, int C, long rows,
                                                  float* __restrict__ rstd, float* __restrict__ mu) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const T* p = x + r * ld;
  float s = 0.f;
  float v[8];
  for (int c = 0; c < C; c += 8) {
    load8<T>(p + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
  }
  const float m = s / (float)C;
  float q = 0.f;
  for (int c = 0; c < C; c += 8) {
    load8<T>(p + c, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - m; q = fmaf(d, d, q); }
  }
  rstd[r] = rsqrtf(q / (float)C + 1e-5f);
  if (mu) mu[r] = m;
}

template <typename T>
int ln_stats(const T* x, long ld, int C, long rows, float* rstd, float* mu, cudaStream_t s) {
  KD_CHECK(C % 8 == 0 && ld % 8 == 0, "ln_stats: C=%d ld=%ld must be multiples of 8", C, ld);
  ProfScope prof(PC_LN_STATS, s, 4.0 * rows * C, (double)rows * (C * sizeof(T) + 4.0 * (mu ? 2 : 1)));
  k_ln_stats<T><<<cdiv(rows, 256), 256, 0, s>>>(x, ld, C, rows, rstd, mu);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int ln_stats<float>(const float*, long, int, long, float*, float*, cudaStream_t);
template int ln_stats<bf16>(const bf16*, long, int, long, float*, float*, cudaStream_t);

// =====================================================================================
// Depthwise 3x3 (qkv_dwconv :119, ffn.dwconv :97) with optional fused GELU gate (:103-104)
// thread = (image row y, 4 consecutive x, 8 channels)
// =====================================================================================
template <typename T, int GATE>
__global__ void __launch_bounds__(128) k_dwconv3x3(const T* __restrict__ x, long ldx, T* __restrict__ out, long ldo,
                                                   const float* __restrict__ w9c, const float* __restrict__ bias,
                                                   int nimg, int H, int W, int C) {
  constexpr int PX = 4;
  const int cgroups = (GATE ? C / 2 : C) / 8;
  const int xblocks = (W + PX - 1) / PX;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)nimg * H * xblocks * cgroups;
  if (idx >= total) return;
  const int cg = (int)(idx % cgroups); idx /= cgroups;
  const int xb = (int)(idx % xblocks); idx /= xblocks;
  const int y = (int)(idx % H);
  const int img = (int)(idx / H);
  const int x0 = xb * PX;
  const int hp = C / 2;

  float res[PX][8];
#pragma unroll
  for (int half = 0; half < (GATE ? 2 : 1); ++half) {
    const int c0 = half * hp + cg * 8;
    float wt[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 a = *reinterpret_cast<const float4*>(w9c + (long)t * C + c0);
      const float4 b = *reinterpret_cast<const float4*>(w9c + (long)t * C + c0 + 4);
      wt[t][0] = a.x; wt[t][1] = a.y; wt[t][2] = a.z; wt[t][3] = a.w;
      wt[t][4] = b.x; wt[t][5] = b.y; wt[t][6] = b.z; wt[t][7] = b.w;
    }
    float acc[PX][8];
#pragma unroll
    for (int p = 0; p < PX; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[p][i] = bias ? bias[c0 + i] : 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;
      const T* rowp = x + ((long)img * H + yy) * W * ldx + c0;
#pragma unroll
      for (int cx = 0; cx < PX + 2; ++cx) {
        const int xx = x0 + cx - 1;
        if (xx < 0 || xx >= W) continue;
        float v[8];
        load8<T>(rowp + (long)xx * ldx, v);
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int dx = cx - p;  // tap column 0..2
          if (dx >= 0 && dx < 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(v[i], wt[dy * 3 + dx][i], acc[p][i]);
          }
        }
      }
    }
    if (half == 0) {
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) res[p][i] = acc[p][i];
    } else {
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) res[p][i] = gelu_erf(res[p][i]) * acc[p][i];
    }
  }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int xx = x0 + p;
    if (xx < W) store8<T>(out + (((long)img * H + y) * W + xx) * ldo + cg * 8, res[p]);
  }
}

int dwconv3x3_f2(const bf16* x, long ldx, bf16* out, long ldo, const float* w9c, int nimg, int H, int W, int C, int gate,
                 cudaStream_t s);

template <typename T>
int dwconv3x3(const T* x, long ldx, T* out, long ldo, const float* w9c, const float* bias, int nimg, int H, int W, int C,
              int gate, cudaStream_t s) {
  if (std::is_same<T, bf16>::value && bias == nullptr) {   // packed-FFMA2 TMA-staged kernel (dwconv_f2.cu); < 0: shape not eligible
    const int r = dwconv3x3_f2(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<bf16*>(out), ldo, w9c, nimg, H, W, C, gate, s);
    if (r >= 0) return r;
  }
  KD_CHECK(C % (gate ? 16 : 8) == 0 && ldx % 8 == 0 && ldo % 8 == 0, "dwconv3x3: C=%d ldx=%ld ldo=%ld alignment", C, ldx, ldo);
  const long total = (long)nimg * H * ((W + 3) / 4) * ((gate ? C / 2 : C) / 8);
  ProfScope prof(PC_DWCONV, s, 18.0 * nimg * H * W * C, (double)nimg * H * W * (C + (gate ? C / 2 : C)) * sizeof(T) + 36.0 * C);
  if (gate) k_dwconv3x3<T, 1><<<cdiv(total, 128), 128, 0, s>>>(x, ldx, out, ldo, w9c, bias, nimg, H, W, C);
  else k_dwconv3x3<T, 0><<<cdiv(total, 128), 128, 0, s>>>(x, ldx, out, ldo, w9c, bias, nimg, H, W, C);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int dwconv3x3<float>(const float*, long, float*, long, const float*, const float*, int, int, int, int, int, cudaStream_t);
template int dwconv3x3<bf16>(const bf16*, long, bf16*, long, const float*, const float*, int, int, int, int, int, cudaStream_t);

// =====================================================================================
// MDTA reductions (KDLAE_model.py:134-137): partial Gram q k^T + squared L2 norms over pixels
// grid = (splits, heads, nimg); thread tile 8x8 of the ch x ch Gram, pixel sub-groups share tiles
// =====================================================================================
constexpr int GRAM_PT = 32;  // pixels staged per step

template <typename T>
__global__ void __launch_bounds__(256) k_mdta_gram(const T* __restrict__ qk, long ld, int HW, int C, int heads, int splits,
                                                   float* __restrict__ part) {
  extern __shared__ float sm[];
  const int ch = C / heads;
  const int split = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  float* qs = sm;                   // [PT][ch]
  float* ks = sm + GRAM_PT * ch;    // [PT][ch]
  const int nb = ch / 8, nblk = nb * nb;
  const int psub = max(1, min(GRAM_PT, 256 / nblk));
  const int tid = threadIdx.x;
  const int blk = tid % nblk, ps = tid / nblk;
  const bool active = ps < psub && tid < nblk * psub;
  const int i0 = (blk / nb) * 8, j0 = (blk % nb) * 8;

  const int per = (HW + splits - 1) / splits;
  const int p_begin = split * per, p_end = min(HW, p_begin + per);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float nrm = 0.f;  // thread t < ch: |q_t|^2 ; ch <= t < 2ch: |k_t|^2

  const T* base = qk + (long)img * HW * ld + head * ch;
  const int vec_per_row = ch / 8;
  for (int p0 = p_begin; p0 < p_end; p0 += GRAM_PT) {
    const int np = min(GRAM_PT, p_end - p0);
    for (int e = tid; e < GRAM_PT * vec_per_row * 2; e += 256) {
      const int which = e / (GRAM_PT * vec_per_row);
      const int r = e % (GRAM_PT * vec_per_row);
      const int p = r / vec_per_row, cv = (r % vec_per_row) * 8;
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (p < np) load8<T>(base + (long)(p0 + p) * ld + which * C + cv, v);
      float* dst = (which ? ks : qs) + p * ch + cv;
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = v[i];
    }
    __syncthreads();
    if (active) {
      for (int p = ps; p < GRAM_PT; p += psub) {
        const float4 qa = *reinterpret_cast<const float4*>(qs + p * ch + i0);
        const float4 qb = *reinterpret_cast<const float4*>(qs + p * ch + i0 + 4);
        const float4 ka = *reinterpret_cast<const float4*>(ks + p * ch + j0);
        const float4 kb = *reinterpret_cast<const float4*>(ks + p * ch + j0 + 4);
        const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        const float k[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(q[i], k[j], acc[i][j]);
      }
    }
    if (tid < 2 * ch) {
      const float* src = (tid < ch) ? (qs + tid) : (ks + tid - ch);
      for (int p = 0; p < GRAM_PT; ++p) { const float v = src[p * ch]; nrm = fmaf(v, v, nrm); }
    }
    __syncthreads();
  }
  // deterministic reduction over pixel sub-groups through shared memory
  float* G = sm;  // [ch][ch] (reuses the staging tiles; ch*ch <= 2*PT*ch needs ch <= 64, else extra space was reserved)
  for (int e = tid; e < ch * ch; e += 256) G[e] = 0.f;
  __syncthreads();
  for (int r = 0; r < psub; ++r) {
    if (active && ps == r) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) G[(i0 + i) * ch + j0 + j] += acc[i][j];
    }
    __syncthreads();
  }
  float* dst = part + (((long)img * heads + head) * splits + split) * (long)(ch * ch + 2 * ch);
  for (int e = tid; e < ch * ch; e += 256) dst[e] = G[e];
  if (tid < 2 * ch) dst[ch * ch + tid] = nrm;
}

// Pixel splits per (image, head).  Depends on the image size only, so the reduction order - and therefore the
// result - of an image does not depend on the batch or micro-batch it is processed in.
int mdta_gram_splits(int HW, int /*nimg_heads*/) {
  int s = HW / 2048;
  return s < 1 ? 1 : (s > 64 ? 64 : s);
}

int mdta_gram_tc(const bf16* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s);

template <typename T>
int mdta_gram(const T* qk, long ld, int nimg, int HW, int C, int heads, int splits, float* part, cudaStream_t s) {
  const int ch = C / heads;
  if (std::is_same<T, bf16>::value) {   // tcgen05 Gram (MN-major operands straight from the dwconv output)
    const int r = mdta_gram_tc(reinterpret_cast<const bf16*>(qk), ld, nimg, HW, C, heads, splits, part, s);
    if (r >= 0) return r;
  }
  KD_CHECK(C % heads == 0 && ch % 8 == 0 && ch <= 128 && 2 * ch <= 256, "mdta_gram: unsupported channels/head %d", ch);
  const size_t smem = sizeof(float) * (size_t)max(2 * GRAM_PT * ch, ch * ch);
  static DeviceOnce once;     // one per instantiation (T)
  bool first; int dev;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_mdta_gram<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    device_mark(once, dev);
  }
  dim3 grid(splits, heads, nimg);
  ProfScope prof(PC_MDTA_GRAM, s, 2.0 * nimg * HW * C * ch + 4.0 * nimg * HW * C,
                 (double)nimg * HW * 2 * C * sizeof(T) + 4.0 * nimg * heads * splits * (ch * ch + 2 * ch));
  k_mdta_gram<T><<<grid, 256, smem, s>>>(qk, ld, HW, C, heads, splits, part);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int mdta_gram<float>(const float*, long, int, int, int, int, int, float*, cudaStream_t);
template int mdta_gram<bf16>(const bf16*, long, int, int, int, int, int, float*, cudaStream_t);

// softmax(cosine-Gram * temperature) (KDLAE_model.py:134-138) folded into project_out (:140,:144):
//   project_out(attn @ v) = (Wp . blockdiag(attn)) @ v   ->  per-image C x C matrix Mb
// Kernel 1: grid (ch/8, heads, nimg) - reduce the split partials of 8 Gram rows (fixed order: deterministic), normalise
//           by max(|q_i|,eps) max(|k_j|,eps), scale by the temperature, softmax; the rows are written in place over
//           split 0 of the partial buffer (each CTA only overwrites the rows that it alone reads).
__global__ void __launch_bounds__(256) k_mdta_softmax(float* __restrict__ part, int C, int heads, int splits,
                                                      const float* __restrict__ temperature) {
  extern __shared__ float sm[];
  const int ch = C / heads;
  const int i0 = blockIdx.x * 8, head = blockIdx.y, img = blockIdx.z, tid = threadIdx.x;
  float* G = sm;                 // [8][ch]
  float* nq = sm + 8 * ch;       // [8]
  float* nk = nq + 8;            // [ch]
  const long psz = (long)ch * ch + 2 * ch;
  float* src = part + ((long)img * heads + head) * splits * psz;
  const int nrow = min(8, ch - i0);
  for (int e = tid; e < nrow * ch + nrow + ch; e += 256) {
    long off;
    if (e < nrow * ch) off = (long)(i0 + e / ch) * ch + e % ch;
    else if (e < nrow * ch + nrow) off = (long)ch * ch + i0 + (e - nrow * ch);
    else off = (long)ch * ch + ch + (e - nrow * ch - nrow);
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += src[(long)sp * psz + off];
    if (e < nrow * ch) G[e] = s;
    else if (e < nrow * ch + nrow) nq[e - nrow * ch] = fmaxf(sqrtf(s), 1e-12f);   // F.normalize eps
    else nk[e - nrow * ch - nrow] = fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < nrow) {   // one warp per softmax row
    const float temp = temperature[head];
    float* g = G + warp * ch;
    float mx = -INFINITY;
    for (int j = lane; j < ch; j += 32) { const float l = g[j] / (nq[warp] * nk[j]) * temp; g[j] = l; mx = fmaxf(mx, l); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < ch; j += 32) { const float e = expf(g[j] - mx); g[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    float* dst = src + (long)(i0 + warp) * ch;    // split 0, rows of this CTA
    for (int j = lane; j < ch; j += 32) dst[j] = g[j] * inv;
  }
}
// Kernel 2: grid (C/16, heads, nimg) - Mb[n][head*ch + j] = sum_i Wp[n][head*ch + i] * attn[i][j] for 16 rows n.
template <typename T>
__global__ void __launch_bounds__(256) k_mdta_project(const float* __restrict__ part, int C, int heads, int splits,
                                                      const float* __restrict__ wproj, T* __restrict__ mb, long mb_ld,
                                                      long mb_img_stride) {
  extern __shared__ float sm[];
  const int ch = C / heads;
  const int n0 = blockIdx.x * 16, head = blockIdx.y, img = blockIdx.z, tid = threadIdx.x;
  float* A = sm;                 // attn [ch][ch]
  float* Wr = sm + ch * ch;      // [16][ch]
  const long psz = (long)ch * ch + 2 * ch;
  const float* src = part + ((long)img * heads + head) * splits * psz;
  for (int e = tid; e < ch * ch; e += 256) A[e] = src[e];
  const int nrow = min(16, C - n0);
  for (int e = tid; e < nrow * ch; e += 256) Wr[e] = wproj[(long)(n0 + e / ch) * C + head * ch + e % ch];
  __syncthreads();
  T* dst = mb + (long)img * mb_img_stride + head * ch;
  for (int e = tid; e < nrow * ch; e += 256) {
    const int r = e / ch, j = e % ch;
    float s = 0.f;
    for (int i = 0; i < ch; ++i) s = fmaf(Wr[r * ch + i], A[i * ch + j], s);
    dst[(long)(n0 + r) * mb_ld + j] = from_f<T>(s);
  }
}

template <typename T>
int mdta_fold(float* part, int nimg, int C, int heads, int splits, const float* temperature, const float* wproj, T* mb,
              long mb_ld, long mb_img_stride, cudaStream_t s) {
  const int ch = C / heads;
  static DeviceOnce once;     // one per instantiation (T)
  bool first; int dev;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_mdta_project<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    device_mark(once, dev);
  }
  ProfScope prof(PC_MDTA_FOLD, s, 2.0 * nimg * C * C * ch, 4.0 * nimg * heads * splits * (ch * ch + 2 * ch) + (double)nimg * C * C * (4 + sizeof(T)));
  k_mdta_softmax<<<dim3(cdiv(ch, 8), heads, nimg), 256, sizeof(float) * (9 * ch + 8), s>>>(part, C, heads, splits, temperature);
  count_launch();
  KD_LAUNCH_CHECK();
  k_mdta_project<T><<<dim3(cdiv(C, 16), heads, nimg), 256, sizeof(float) * ((size_t)ch * ch + 16 * ch), s>>>(part, C, heads, splits, wproj, mb,
                                                                                                       mb_ld, mb_img_stride);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int mdta_fold<float>(float*, int, int, int, int, const float*, const float*, float*, long, long, cudaStream_t);
template int mdta_fold<bf16>(float*, int, int, int, int, const float*, const float*, bf16*, long, long, cudaStream_t);

// =====================================================================================
// Direct convolutions with very few input channels (patch_embed, output_param, cen, student conv 1,
// ASDQE stems).  thread = (pixel, 8 output channels); planar fp32 inputs.
// =====================================================================================
// CIN = total input planes (1..4), KD = temporal taps (1 or 3).  All tap inputs are gathered first (independent,
// predicated loads), the [taps*CIN][cout] weights sit in shared memory.
// PXF = consecutive output pixels per thread: one weight fetch feeds PXF pixels (wide outputs: the weight LDS traffic is the
// limiter); narrow outputs (ASDQE stems, 16 channels) keep PXF = 1 so a warp's input loads stay contiguous.
// packed fp32 helpers (exact per lane: fma.rn.f32x2 = two IEEE fmaf): the 8 output channels of a thread are four pairs, so the
// tap loop issues half as many FMA instructions
__device__ __forceinline__ unsigned long long fi_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fi_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fi_unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <typename T, int CIN, int KD, int PXF, int DIL>
__global__ void __launch_bounds__(128) k_conv_few_in(const SmallConv op, int Hin, int Win) {
  extern __shared__ __align__(16) float wsm_in[];   // [KD*9*CIN][cout]
  constexpr int TAPS = KD * 9;
  constexpr int RW = PXF + 2 * DIL;                 // input columns one thread needs per (frame, row, channel)
  for (int e = threadIdx.x; e < TAPS * CIN * op.cout; e += blockDim.x) wsm_in[e] = op.w[e];
  __syncthreads();
  const unsigned cgroups = op.cout / 8;
  const unsigned xgroups = (unsigned)(op.W + PXF - 1) / PXF;
  const unsigned total = (unsigned)op.H * xgroups * cgroups;     // 32-bit index math: one image per blockIdx.y
  const int img = blockIdx.y;
  const int d = img % op.D, b = img / op.D;
  // grid-stride over (row, pixel group, channel group): the weight staging above is paid once per block
  for (unsigned idx0 = blockIdx.x * blockDim.x + threadIdx.x; idx0 < total; idx0 += gridDim.x * blockDim.x) {
    unsigned idx = idx0;
    const int cg = (int)(idx % cgroups); idx /= cgroups;
    const int x0 = (int)(idx % xgroups) * PXF;
    const int y = (int)(idx / xgroups);
    unsigned long long acc2[PXF][4];
#pragma unroll
    for (int q = 0; q < PXF; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        acc2[q][i] = op.bias ? fi_pack2(op.bias[cg * 8 + 2 * i], op.bias[cg * 8 + 2 * i + 1]) : fi_pack2(0.f, 0.f);
#pragma unroll
    for (int td = 0; td < KD; ++td) {
      const int dd = d + td - KD / 2;
      const bool d_ok = dd >= 0 && dd < op.D;
      const long im = (long)b * op.D + min(max(dd, 0), op.D - 1);
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        const int yy = y + (ty - 1) * DIL;
        const bool y_ok = d_ok && yy >= 0 && yy < Hin;
        const long rowo = (long)min(max(yy, 0), Hin - 1) * Win;
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          // the RW input columns of this (frame, row, channel) are loaded once and feed the three horizontal taps of all PXF
          // pixels (the first version re-loaded every tap: 3 * PXF loads for 8 * 3 * PXF FMAs).  Loads are unconditional from a
          // clamped address and zeroed by a select: with `if (in range) load` every LDG sat in its own divergent region
          // (BSSY / BRA / BSYNC) and its latency was paid serially - 3x the FMA time of the 27-tap first layer of KDLAE-S.
          float r[RW];
          const float* src = (c < op.cin0) ? op.in0 + im * op.in0_img + (long)c * op.in0_ch
                                           : op.in1 + im * op.in1_img + (long)(c - op.cin0) * op.in1_ch;
          const long px = (c < op.cin0) ? 1 : op.in1_px;
          const float* sub = (c < op.cin0 && op.sub0) ? op.sub0 + im * op.in0_img + (long)c * op.in0_ch : nullptr;
#pragma unroll
          for (int j = 0; j < RW; ++j) {
            const int xx = x0 - DIL + j;
            const bool ok = y_ok && xx >= 0 && xx < Win;
            const long sp = (rowo + min(max(xx, 0), Win - 1)) * px;
            float t = __ldg(src + sp);
            if (sub) t -= __ldg(sub + sp);
            r[j] = ok ? t : 0.f;
          }
#pragma unroll
          for (int tx = 0; tx < 3; ++tx) {
            const float* wp = wsm_in + (((td * 3 + ty) * 3 + tx) * CIN + c) * op.cout + cg * 8;
            const ulonglong2 wa = *reinterpret_cast<const ulonglong2*>(wp);        // channel pairs (0,1) (2,3)
            const ulonglong2 wb = *reinterpret_cast<const ulonglong2*>(wp + 4);    //               (4,5) (6,7)
#pragma unroll
            for (int q = 0; q < PXF; ++q) {
              const float v = r[q + tx * DIL];
              const unsigned long long v2 = fi_pack2(v, v);
              acc2[q][0] = fi_ffma2(v2, wa.x, acc2[q][0]); acc2[q][1] = fi_ffma2(v2, wa.y, acc2[q][1]);
              acc2[q][2] = fi_ffma2(v2, wb.x, acc2[q][2]); acc2[q][3] = fi_ffma2(v2, wb.y, acc2[q][3]);
            }
          }
        }
      }
    }
    T* out = reinterpret_cast<T*>(op.out);
    float acc[PXF][8];
#pragma unroll
    for (int q = 0; q < PXF; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) fi_unpack2(acc2[q][i], acc[q][2 * i], acc[q][2 * i + 1]);
#pragma unroll
    for (int q = 0; q < PXF; ++q) {
      if (x0 + q < op.W) {
        if (op.relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[q][i] = fmaxf(acc[q][i], 0.f);
        }
        store8<T>(out + (((long)img * op.H + y) * op.W + x0 + q) * op.out_ld + cg * 8, acc[q]);
      }
    }
  }
}

template <typename T>
int conv_few_in_sized(const SmallConv& op, int Hin, int Win, cudaStream_t s) {
  KD_CHECK(op.cout % 8 == 0 && op.out_ld % 8 == 0, "conv_few_in: cout=%d must be a multiple of 8", op.cout);
  const int cin = op.cin0 + op.cin1;
  KD_CHECK(cin >= 1 && cin <= 4 && (op.kd == 1 || (op.kd == 3 && cin == 1)), "conv_few_in: unsupported cin=%d kd=%d", cin, op.kd);
  KD_CHECK(op.dil == 1 || (op.dil == 2 && op.kd == 1), "conv_few_in: unsupported dilation %d", op.dil);
  const int pxf = (op.W >= 16) ? 4 : 1;
  const long total = (long)op.H * ((op.W + pxf - 1) / pxf) * (op.cout / 8);
  KD_CHECK((long)op.H * op.W * (op.cout / 8) < (1L << 31) && op.nimg <= 65535, "conv_few_in: image too large");
  const double fi_pix = (double)op.nimg * op.H * op.W;
  ProfScope prof(PC_SMALL_CONV, s, 2.0 * fi_pix * op.cout * cin * 9 * op.kd,
                 fi_pix * (4.0 * cin * (op.sub0 ? 2 : 1) + (double)op.cout * sizeof(T)));
  const dim3 grid((unsigned)std::min<long>(cdiv(total, 128), std::max<long>(1, 148L * 16 / op.nimg)), op.nimg);
  const size_t smem = sizeof(float) * (size_t)op.kd * 9 * cin * op.cout;
  KD_CHECK(smem <= 48 * 1024, "conv_few_in: weights do not fit shared memory");
#define KD_FEW_IN(CI, KDD, DL) do { if (pxf == 4) k_conv_few_in<T, CI, KDD, 4, DL><<<grid, 128, smem, s>>>(op, Hin, Win); \
                                    else k_conv_few_in<T, CI, KDD, 1, DL><<<grid, 128, smem, s>>>(op, Hin, Win); } while (0)
#define KD_FEW_IN_D(CI) do { if (op.dil == 2) KD_FEW_IN(CI, 1, 2); else KD_FEW_IN(CI, 1, 1); } while (0)
  if (op.kd == 3) KD_FEW_IN(1, 3, 1);
  else if (cin == 1) KD_FEW_IN_D(1);
  else if (cin == 2) KD_FEW_IN_D(2);
  else if (cin == 3) KD_FEW_IN_D(3);
  else KD_FEW_IN_D(4);
#undef KD_FEW_IN_D
#undef KD_FEW_IN
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template <typename T> int conv_few_in(const SmallConv& op, cudaStream_t s) { return conv_few_in_sized<T>(op, op.H, op.W, s); }
template int conv_few_in<float>(const SmallConv&, cudaStream_t);
template int conv_few_in<bf16>(const SmallConv&, cudaStream_t);
template int conv_few_in_sized<float>(const SmallConv&, int, int, cudaStream_t);
template int conv_few_in_sized<bf16>(const SmallConv&, int, int, cudaStream_t);

// Few output channels (output, output2, outputen, student out_conv): thread = pixel, weights in smem.
template <typename T>
__global__ void __launch_bounds__(128) k_conv_few_out(const SmallConvOut op) {
  extern __shared__ __align__(16) float wsm[];  // [cout][taps][cin]
  const int taps = op.k * op.k;
  const int wn = op.cout * taps * op.cin;
  for (int e = threadIdx.x; e < wn; e += blockDim.x) wsm[e] = op.w[e];
  __syncthreads();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // 32-bit index math: one image per blockIdx.y
  if (idx >= (unsigned)op.H * op.W) return;
  const int x = (int)(idx % (unsigned)op.W);
  const int y = (int)(idx / (unsigned)op.W);
  const int img = blockIdx.y;
  const T* in = reinterpret_cast<const T*>(op.in);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int hk = op.k / 2;
  for (int ty = 0; ty < op.k; ++ty) {
    const int yy = y + ty - hk;
    if (yy < 0 || yy >= op.H) continue;
    for (int tx = 0; tx < op.k; ++tx) {
      const int xx = x + tx - hk;
      if (xx < 0 || xx >= op.W) continue;
      const T* p = in + (((long)img * op.H + yy) * op.W + xx) * op.in_ld;
      const float* wt = wsm + (ty * op.k + tx) * op.cin;
      for (int c = 0; c < op.cin; c += 8) {
        float v[8];
        load8<T>(p + c, v);
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          if (co < op.cout) {
            const float* wc = wt + co * taps * op.cin + c;     // 32-byte aligned: cin % 8 == 0
            const float4 w0 = *reinterpret_cast<const float4*>(wc), w1 = *reinterpret_cast<const float4*>(wc + 4);
            acc[co] = fmaf(v[0], w0.x, acc[co]); acc[co] = fmaf(v[1], w0.y, acc[co]);
            acc[co] = fmaf(v[2], w0.z, acc[co]); acc[co] = fmaf(v[3], w0.w, acc[co]);
            acc[co] = fmaf(v[4], w1.x, acc[co]); acc[co] = fmaf(v[5], w1.y, acc[co]);
            acc[co] = fmaf(v[6], w1.z, acc[co]); acc[co] = fmaf(v[7], w1.w, acc[co]);
          }
        }
      }
    }
  }
  const long sp = (long)y * op.W + x;
  for (int co = 0; co < op.cout; ++co) {
    float t = acc[co] + (op.bias ? op.bias[co] : 0.f);
    if (op.res) t += op.res[(long)img * op.res_img + (long)co * op.res_ch + sp];
    op.out[(long)img * op.out_img + (long)co * op.out_ch + sp] = t;
  }
}

template <typename T>
int conv_few_out(const SmallConvOut& op, cudaStream_t s) {
  KD_CHECK(op.cout >= 1 && op.cout <= 4 && op.cin % 8 == 0 && op.in_ld % 8 == 0, "conv_few_out: cout=%d cin=%d", op.cout, op.cin);
  const size_t smem = sizeof(float) * (size_t)op.cout * op.k * op.k * op.cin;
  KD_CHECK(smem <= 48 * 1024, "conv_few_out: weights do not fit shared memory");
  const long total = (long)op.nimg * op.H * op.W;
  KD_CHECK((long)op.H * op.W < (1L << 31) && op.nimg <= 65535, "conv_few_out: image too large");
  ProfScope prof(PC_SMALL_CONV, s, 2.0 * total * op.cout * op.cin * op.k * op.k,
                 (double)total * (op.cin * sizeof(T) + 4.0 * op.cout * (op.res ? 2 : 1)));
  k_conv_few_out<T><<<dim3(cdiv((long)op.H * op.W, 128), op.nimg), 128, smem, s>>>(op);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
// ---- 3x3 tap gather-sum over per-tap partial planes (fp32 planar) ----
__global__ void __launch_bounds__(256) k_tap_sum(const float* __restrict__ part, int cout, int H, int W, const float* __restrict__ res,
                                                 long res_img, long res_ch, float* __restrict__ out, long out_img, long out_ch) {
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int img = blockIdx.z / cout, co = blockIdx.z - img * cout;
  if (x >= W || y >= H) return;
  const long HW = (long)H * W;
  const float* pp = part + ((long)img * cout + co) * 9 * HW;
  float acc = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc += __ldg(pp + t * HW + (long)yy * W + xx);
  }
  const long sp = (long)y * W + x;
  if (res) acc += __ldg(res + img * res_img + co * res_ch + sp);
  out[img * out_img + co * out_ch + sp] = acc;
}

int tap_sum(const float* part, int cout, int nimg, int H, int W, const float* res, long res_img, long res_ch, float* out, long out_img,
            long out_ch, cudaStream_t s) {
  KD_CHECK((long)nimg * cout <= 65535, "tap_sum: too many planes");
  const double px = (double)nimg * cout * H * W;
  ProfScope prof(PC_SMALL_CONV, s, 9.0 * px, px * (9 + 1 + (res ? 1 : 0)) * 4.0);
  k_tap_sum<<<dim3(cdiv(W, 64), cdiv(H, 4), nimg * cout), 256, 0, s>>>(part, cout, H, W, res, res_img, res_ch, out, out_img, out_ch);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

template int conv_few_out<float>(const SmallConvOut&, cudaStream_t);
template int conv_few_out<bf16>(const SmallConvOut&, cudaStream_t);

// =====================================================================================
// MaxPool 2x2 (MaxPool3d(1,2,2) / MaxPool2d(2)), bilinear x2 align_corners=True, GAP+MLP, layout
// =====================================================================================
template <typename T>
__global__ void __launch_bounds__(256) k_maxpool2x2(const T* __restrict__ x, T* __restrict__ out, int nimg, int H, int W, int C) {
  const int OH = H / 2, OW = W / 2, cg = C / 8;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)nimg * OH * OW * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 8; idx /= cg;
  const int ox = (int)(idx % OW); idx /= OW;
  const int oy = (int)(idx % OH);
  const int img = (int)(idx / OH);
  float m[8], v[8];
  const T* p = x + (((long)img * H + 2 * oy) * W + 2 * ox) * C + c;
  load8<T>(p, m);
  load8<T>(p + C, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
  load8<T>(p + (long)W * C, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
  load8<T>(p + (long)W * C + C, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
  store8<T>(out + (((long)img * OH + oy) * OW + ox) * C + c, m);
}
template <typename T>
int maxpool2x2(const T* x, T* out, int nimg, int H, int W, int C, cudaStream_t s) {
  KD_CHECK(C % 8 == 0, "maxpool2x2: C=%d", C);
  const long total = (long)nimg * (H / 2) * (W / 2) * (C / 8);
  ProfScope prof(PC_POOL_RESAMPLE, s, 0.0, (double)nimg * H * W * C * sizeof(T) * 1.25);
  k_maxpool2x2<T><<<cdiv(total, 256), 256, 0, s>>>(x, out, nimg, H, W, C);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int maxpool2x2<float>(const float*, float*, int, int, int, int, cudaStream_t);
template int maxpool2x2<bf16>(const bf16*, bf16*, int, int, int, int, cudaStream_t);

// grid (x chunks, output row, image): 32-bit index math only (the first version decoded a 64-bit linear index with three
// 64-bit divisions per thread and ran at 1.8 TB/s)
template <typename T>
__global__ void __launch_bounds__(256) k_upsample2x(const T* __restrict__ x, T* __restrict__ out, int H, int W, int C, int OH, int OW,
                                                    float sy, float sx) {
  const unsigned cg = (unsigned)C / 8;
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (unsigned)OW * cg) return;
  const int c = (int)(e % cg) * 8, ox = (int)(e / cg);
  const int oy = blockIdx.y, img = blockIdx.z;
  // align_corners=True: src = dst * (in-1)/(out-1)   (ASDQE_model.py:54)
  const float fy = sy * oy, fx = sx * ox;
  const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - y0, lx = fx - x0;
  float a[8], b[8], cc[8], d[8], r[8];
  const T* base = x + (long)img * H * W * C + c;
  load8<T>(base + ((long)y0 * W + x0) * C, a);
  load8<T>(base + ((long)y0 * W + x1) * C, b);
  load8<T>(base + ((long)y1 * W + x0) * C, cc);
  load8<T>(base + ((long)y1 * W + x1) * C, d);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    r[i] = (1.f - ly) * ((1.f - lx) * a[i] + lx * b[i]) + ly * ((1.f - lx) * cc[i] + lx * d[i]);
  store8<T>(out + (((long)img * OH + oy) * OW + ox) * C + c, r);
}
template <typename T>
int upsample_bilinear2x(const T* x, T* out, int nimg, int H, int W, int C, int OH, int OW, cudaStream_t s) {
  KD_CHECK(C % 8 == 0 && OH <= 65535 && nimg <= 65535, "upsample: C=%d OH=%d nimg=%d", C, OH, nimg);
  ProfScope prof(PC_POOL_RESAMPLE, s, 0.0, (double)nimg * C * sizeof(T) * ((double)H * W + (double)OH * OW));
  const float sy = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float sx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  k_upsample2x<T><<<dim3(cdiv((long)OW * (C / 8), 256), OH, nimg), 256, 0, s>>>(x, out, H, W, C, OH, OW, sy, sx);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int upsample_bilinear2x<float>(const float*, float*, int, int, int, int, int, int, cudaStream_t);
template int upsample_bilinear2x<bf16>(const bf16*, bf16*, int, int, int, int, int, int, cudaStream_t);

// GAP partial sums: grid (chunks, nimg), each block sums a pixel range for all C channels (C <= 64)
template <typename T>
__global__ void __launch_bounds__(256) k_gap_partial(const T* __restrict__ f, int HW, int C, int chunks, float* __restrict__ scratch) {
  __shared__ float red[256];
  const int chunk = blockIdx.x, img = blockIdx.y, tid = threadIdx.x;
  const int per = (HW + chunks - 1) / chunks;
  const int p0 = chunk * per, p1 = min(HW, p0 + per);
  const int lanes_per_c = 256 / C;       // pixel lanes per channel
  const int c = tid % C, lane = tid / C;
  float s = 0.f;
  if (lane < lanes_per_c)
    for (int p = p0 + lane; p < p1; p += lanes_per_c) s += to_f<T>(f[((long)img * HW + p) * C + c]);
  red[tid] = (lane < lanes_per_c) ? s : 0.f;
  __syncthreads();
  if (tid < C) {
    float t = 0.f;
    for (int l = 0; l < lanes_per_c; ++l) t += red[l * C + tid];
    scratch[((long)img * chunks + chunk) * C + tid] = t;
  }
}
// regressor (ASDQE_model.py:143-154): mean -> Linear(C,256) ReLU -> Linear(256,64) ReLU -> Linear(64,1) -> tanh
__global__ void __launch_bounds__(256) k_mlp_tanh(const float* __restrict__ scratch, int chunks, int C0, float inv_hw,
                                                  const float* __restrict__ outc_w, const float* __restrict__ outc_b, int C,
                                                  const float* __restrict__ w1, const float* __restrict__ b1,
                                                  const float* __restrict__ w2, const float* __restrict__ b2,
                                                  const float* __restrict__ w3, const float* __restrict__ b3,
                                                  float* __restrict__ score) {
  __shared__ float g[64], f[64], h1[256], h2[64];
  const int img = blockIdx.x, tid = threadIdx.x;
  if (tid < C0) {
    float t = 0.f;
    for (int k = 0; k < chunks; ++k) t += scratch[((long)img * chunks + k) * C0 + tid];
    g[tid] = t * inv_hw;
  }
  __syncthreads();
  if (tid < C) {                       // outc on the channel means (mean and the 1x1 conv commute)
    float t = outc_b[tid];
    for (int c = 0; c < C0; ++c) t = fmaf(outc_w[tid * C0 + c], g[c], t);
    f[tid] = t;
  }
  __syncthreads();
  {
    float t = b1[tid];
    for (int c = 0; c < C; ++c) t = fmaf(w1[tid * C + c], f[c], t);
    h1[tid] = fmaxf(t, 0.f);
  }
  __syncthreads();
  if (tid < 64) {
    float t = b2[tid];
    for (int c = 0; c < 256; ++c) t = fmaf(w2[tid * 256 + c], h1[c], t);
    h2[tid] = fmaxf(t, 0.f);
  }
  __syncthreads();
  if (tid == 0) {
    float t = b3[0];
    for (int c = 0; c < 64; ++c) t = fmaf(w3[c], h2[c], t);
    score[img] = tanhf(t);
  }
}
template <typename T>
int gap_mlp_tanh(const T* feat, int nimg, int HW, int C, const float* outc_w, const float* outc_b, int Cf, const float* w1,
                 const float* b1, const float* w2, const float* b2, const float* w3, const float* b3, float* score, float* scratch,
                 cudaStream_t s) {
  KD_CHECK(C <= 64 && Cf <= 64 && 256 % C == 0, "gap_mlp_tanh: C=%d Cf=%d unsupported", C, Cf);
  const int chunks = 64;
  ProfScope prof(PC_HEAD, s, 0.0, (double)nimg * HW * C * sizeof(T));
  k_gap_partial<T><<<dim3(chunks, nimg), 256, 0, s>>>(feat, HW, C, chunks, scratch);
  count_launch();
  KD_LAUNCH_CHECK();
  k_mlp_tanh<<<nimg, 256, 0, s>>>(scratch, chunks, C, 1.f / (float)HW, outc_w, outc_b, Cf, w1, b1, w2, b2, w3, b3, score);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int gap_mlp_tanh<float>(const float*, int, int, int, const float*, const float*, int, const float*, const float*, const float*,
                                 const float*, const float*, const float*, float*, float*, cudaStream_t);
template int gap_mlp_tanh<bf16>(const bf16*, int, int, int, const float*, const float*, int, const float*, const float*, const float*,
                                const float*, const float*, const float*, float*, float*, cudaStream_t);

template <typename T>
__global__ void __launch_bounds__(256) k_nhwc_to_planar(const T* __restrict__ x, long ld, float* __restrict__ out, int HW, int C, long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int p = (int)(idx % HW);
  const int c = (int)((idx / HW) % C);
  const long img = idx / ((long)HW * C);
  out[idx] = to_f<T>(x[(img * HW + p) * ld + c]);
}
template <typename T>
int nhwc_to_planar(const T* x, long ld, float* out, int nimg, int HW, int C, cudaStream_t s) {
  const long total = (long)nimg * HW * C;
  ProfScope prof(PC_HEAD, s, 0.0, (double)total * (sizeof(T) + 4.0));
  k_nhwc_to_planar<T><<<cdiv(total, 256), 256, 0, s>>>(x, ld, out, HW, C, total);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}
template int nhwc_to_planar<float>(const float*, long, float*, int, int, int, cudaStream_t);
template int nhwc_to_planar<bf16>(const bf16*, long, float*, int, int, int, cudaStream_t);

}  // namespace kd
