// CUDA-core implicit-GEMM convolution (fp32 accumulate) for the reference-grade fp32 path and for
// the shapes the tcgen05 kernel does not take.  Implements the same ConvOp / Epilogue contract as
// gemm_tc.cu so both paths share the host orchestration.
//
// Tile: 128 pixels x 64 output channels x 16 K per step, 256 threads, 4x8 outputs per thread.
#include <type_traits>
#include "ops.cuh"

namespace kd {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

struct SimtParams {
  ConvOp op;
  int ctot, taps, ktot;     // channels per tap, number of taps, taps*ctot
  long rows_per_group;      // rows handled by one weight group
};

template <typename T>
__global__ void __launch_bounds__(NT) k_conv_gemm_simt(const SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const ConvOp& op = p.op;
  const int tid = threadIdx.x;
  const int g = blockIdx.z;
  const long row0 = (long)g * p.rows_per_group + (long)blockIdx.x * BM;
  const long row_end = (long)(g + 1) * p.rows_per_group;
  const int n0 = blockIdx.y * BN;
  const T* __restrict__ a0 = reinterpret_cast<const T*>(op.a0);
  const T* __restrict__ a1 = reinterpret_cast<const T*>(op.a1);
  const T* __restrict__ w = reinterpret_cast<const T*>(op.w) + (long)g * op.w_group_stride;

  // A loader: thread -> (row = tid % 128, k half = tid / 128 -> 8 consecutive k)
  const int lrow = tid & (BM - 1);
  const int lk0 = (tid >> 7) * 8;
  const long prow = row0 + lrow;
  const bool row_ok = prow < row_end;
  int px = 0, py = 0, pd = 0, pb = 0;
  if (row_ok) {
    long t = prow;
    px = (int)(t % op.W); t /= op.W;
    py = (int)(t % op.H); t /= op.H;
    pd = (int)(t % op.D); pb = (int)(t / op.D);
  }
  // B loader: thread -> (n = tid / 4 -> 64 rows, k quarter = tid % 4 -> 4 consecutive k)
  const int bn = tid >> 2, bk0 = (tid & 3) * 4;

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int rg = tid >> 3, cg = tid & 7;  // 32 row groups x 8 col groups
  const int hk = op.kh / 2, hw_ = op.kw / 2, hd = op.kd / 2;

  for (int k0 = 0; k0 < p.ktot; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kk = k0 + lk0 + i;
      float v = 0.f;
      if (row_ok && kk < p.ktot) {
        const int tap = kk / p.ctot, c = kk - tap * p.ctot;
        int tx = tap % op.kw, ty = (tap / op.kw) % op.kh, td = tap / (op.kw * op.kh);
        const int x = px + (tx - hw_) * op.dil, y = py + (ty - hk) * op.dil, d = pd + (td - hd);
        if (x >= 0 && x < op.W && y >= 0 && y < op.H && d >= 0 && d < op.D) {
          const long sp = (((long)pb * op.D + d) * op.H + y) * op.W + x;
          v = (c < op.c0) ? to_f<T>(a0[sp * op.ld0 + c]) : to_f<T>(a1[sp * op.ld1 + (c - op.c0)]);
        }
      }
      As[lk0 + i][lrow] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = k0 + bk0 + i, n = n0 + bn;
      float v = 0.f;
      if (kk < p.ktot && n < op.epi.N) {
        const int tap = kk / p.ctot, c = kk - tap * p.ctot;
        v = to_f<T>(w[(long)n * op.w_ld + (long)tap * op.w_tap_ld + c]);
      }
      Bs[bk0 + i][bn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[8];
      const float4 av = *reinterpret_cast<const float4*>(&As[k][rg * 4]);
      a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][cg * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][cg * 8 + 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long r = row0 + rg * 4 + i;
    if (r >= row_end) continue;
    long t = r;
    const int x = (int)(t % op.W); t /= op.W;
    const int y = (int)(t % op.H); t /= op.H;   // t = b*D + d
    epilogue_store8<T>(op.epi, r, (int)t, y, x, n0 + cg * 8, acc[i]);
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// fp32 1x1 conv = plain SGEMM  out[P][N] = A[P][K] . W[N][K]^T : the dominant kernel of the fp32 training step (every qkv /
// project / GDFN 1x1, forward and dgrad) and of the fp32 inference path.  8 x 8 outputs per thread, K in steps of 8 with the
// next step's global float4 loads in flight under the FMAs (register prefetch + two shared-memory buffers), 16-byte loads on both
// operands.  CG = column groups: 16 -> tile 128 x 128, 8 -> tile 256 x 64 (narrow N such as the 48-channel project_out).
// Same ConvOp / Epilogue contract as the generic kernel (row scale, bias, residual, ragged N, per-image weight groups).
// ---------------------------------------------------------------------------------------------------------------------------
// KS = 3: the same kernel as an implicit GEMM over the 9 taps of a dense 3x3 conv (dilation d, zero padding): K = 9 * C with
// C % 8 == 0, so a K step of 8 stays inside one tap and only the A loader changes (shifted pixel, zero outside the image).
template <int CG, int KS>
__global__ void __launch_bounds__(256, 2) k_sgemm_1x1(const SimtParams p) {
  constexpr int RG = 256 / CG, SBM = RG * 8, SBN = CG * 8, SBK = 8;
  constexpr int NA = SBM / 128;                 // float4 loads of A per thread and step
  __shared__ __align__(16) float As[2][SBK][SBM + 4];
  __shared__ __align__(16) float Bs[2][SBK][SBN + 4];
  const ConvOp& op = p.op;
  const int tid = threadIdx.x, g = blockIdx.z;
  const long row0 = (long)g * p.rows_per_group + (long)blockIdx.x * SBM;
  const long row_end = (long)(g + 1) * p.rows_per_group;
  const int n0 = blockIdx.y * SBN, K = op.c0 * KS * KS;
  const float* __restrict__ a = reinterpret_cast<const float*>(op.a0);
  const float* __restrict__ w = reinterpret_cast<const float*>(op.w) + (long)g * op.w_group_stride;
  // loaders: float4 f of a tile covers row f / 2, k offset (f & 1) * 4
  const int lk = (tid & 1) * 4;
  const float* ap[NA];
  bool a_ok[NA];
  int ay[NA], ax[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const long r = row0 + (tid >> 1) + i * 128;
    a_ok[i] = r < row_end;
    ap[i] = a + (a_ok[i] ? r : row0) * op.ld0 + lk;
    ax[i] = (int)(r % op.W); ay[i] = (int)((r / op.W) % op.H);
  }
  const int brow = tid >> 1;
  const bool b_act = brow < SBN, b_ok = b_act && (n0 + brow) < op.epi.N;
  const float* bp = w + (long)(b_ok ? n0 + brow : 0) * op.w_ld + lk;
  float4 ra[NA], rb;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto gload = [&](int k0) {
    const bool k_ok = k0 + lk < K;              // K % 4 == 0: a float4 is entirely inside or outside
    if (KS == 1) {
#pragma unroll
      for (int i = 0; i < NA; ++i) ra[i] = (a_ok[i] && k_ok) ? __ldg(reinterpret_cast<const float4*>(ap[i] + k0)) : zero4;
    } else {
      const int tap = k0 / op.c0, c = k0 - tap * op.c0;
      const int dy = (tap / 3 - 1) * op.dil, dx = (tap % 3 - 1) * op.dil;
      const long shift = ((long)dy * op.W + dx) * op.ld0 + c;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const bool in = a_ok[i] && k_ok && ay[i] + dy >= 0 && ay[i] + dy < op.H && ax[i] + dx >= 0 && ax[i] + dx < op.W;
        ra[i] = in ? __ldg(reinterpret_cast<const float4*>(ap[i] + shift)) : zero4;
      }
    }
    rb = (b_ok && k_ok) ? __ldg(reinterpret_cast<const float4*>(bp + k0)) : zero4;
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int r = (tid >> 1) + i * 128;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
    }
    if (b_act) { Bs[buf][lk + 0][brow] = rb.x; Bs[buf][lk + 1][brow] = rb.y; Bs[buf][lk + 2][brow] = rb.z; Bs[buf][lk + 3][brow] = rb.w; }
  };
  const int ty = tid / CG, tx = tid % CG;       // rows ty * 4 + {0..3} and SBM / 2 + ty * 4 + {0..3}; columns tx * 8 + {0..7}
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = (K + SBK - 1) / SBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * SBK);
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][SBM / 2 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);                          // the other buffer was last read before the previous barrier
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long r = row0 + (i < 4 ? ty * 4 + i : SBM / 2 + ty * 4 + (i - 4));
    if (r >= row_end) continue;
    epilogue_store8<float>(op.epi, r, 0, 0, 0, n0 + tx * 8, acc[i]);     // identity addressing: (image, y, x) are not used
  }
}

bool sgemm_1x1_ok(const ConvOp& op) {
  const bool k1 = op.kh == 1 && op.kw == 1 && op.c0 % 4 == 0;
  const bool k3 = op.kh == 3 && op.kw == 3 && op.c0 % 8 == 0 && op.w_tap_ld == op.c0 && op.groups == 1;   // taps contiguous in K
  return op.kd == 1 && (k1 || k3) && op.c1 == 0 && op.epi.mode == OUT_IDENTITY && op.ld0 % 4 == 0 &&
         op.w_ld % 4 == 0 && op.w_group_stride % 4 == 0 && !(reinterpret_cast<uintptr_t>(op.a0) & 15) &&
         !(reinterpret_cast<uintptr_t>(op.w) & 15);
}

}  // namespace

template <typename T>
int conv_gemm_simt(const ConvOp& op, cudaStream_t s) {
  SimtParams p;
  p.op = op;
  p.ctot = op.c0 + op.c1;
  p.taps = op.kd * op.kh * op.kw;
  p.ktot = p.taps * p.ctot;
  const long rows = (long)op.nimg * op.H * op.W;
  KD_CHECK(op.groups >= 1 && rows % op.groups == 0, "conv_gemm_simt: rows %ld not divisible by groups %d", rows, op.groups);
  p.rows_per_group = rows / op.groups;
  KD_CHECK(op.epi.N > 0 && (op.epi.out != nullptr || op.epi.planar_out != nullptr), "conv_gemm_simt: bad epilogue");
  if (op.epi.mode == OUT_PIXEL_SHUFFLE) KD_CHECK(op.epi.cq % 8 == 0 && op.epi.N == 4 * op.epi.cq, "pixel-shuffle needs cq%%8==0");
  dim3 grid(cdiv(p.rows_per_group, BM), cdiv(op.epi.N, BN), op.groups);
  KD_CHECK(grid.y <= 65535 && grid.z <= 65535, "conv_gemm_simt: grid too large");
  ProfScope prof(PC_GEMM_SIMT, s, 2.0 * rows * op.epi.N * p.ktot,
                 sizeof(T) * ((double)rows * (p.ctot + op.epi.N * (op.epi.res ? 2 : 1)) + (double)op.groups * op.epi.N * p.ktot));
  if constexpr (std::is_same<T, float>::value) {
    if (sgemm_1x1_ok(op)) {
      const int N = op.epi.N;
      const bool narrow = cdiv(N, 64) * 64 < cdiv(N, 128) * 128;       // narrow / ragged N: the 256 x 64 tile pads less
      const dim3 g8(cdiv(p.rows_per_group, 256), cdiv(N, 64), op.groups), g16(cdiv(p.rows_per_group, 128), cdiv(N, 128), op.groups);
      if (op.kh == 3) {
        if (narrow) k_sgemm_1x1<8, 3><<<g8, 256, 0, s>>>(p);
        else k_sgemm_1x1<16, 3><<<g16, 256, 0, s>>>(p);
      } else {
        if (narrow) k_sgemm_1x1<8, 1><<<g8, 256, 0, s>>>(p);
        else k_sgemm_1x1<16, 1><<<g16, 256, 0, s>>>(p);
      }
      count_launch();
      KD_LAUNCH_CHECK();
      return 0;
    }
  }
  k_conv_gemm_simt<T><<<grid, NT, 0, s>>>(p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

template int conv_gemm_simt<float>(const ConvOp&, cudaStream_t);
template int conv_gemm_simt<bf16>(const ConvOp&, cudaStream_t);

template <> int conv_gemm<float>(const ConvOp& op, cudaStream_t s) { return conv_gemm_simt<float>(op, s); }
template <> int conv_gemm<bf16>(const ConvOp& op, cudaStream_t s) {
  if (conv_gemm_tc_eligible(op)) return conv_gemm_tc(op, s);
  return conv_gemm_simt<bf16>(op, s);
}

}  // namespace kd
