// Depthwise 3x3 (+ optional GELU gate) for NHWC bf16, TMA-staged (KDLAE_model.py:97,103-104,119).
//
// One CTA = a 32x8 pixel tile x 64 channels.  A single 4-D TMA box {64 ch, 34, 10, 1} brings the tile and
// its 1-pixel halo into shared memory; the conv's zero padding is TMA's out-of-bounds fill, so the
// compute loop has no boundary tests.  thread = (8 channels, 8 consecutive pixels of one row): 30 LDS.128
// feed 576 FMAs; the three weight vectors of a filter row are re-read per row to stay under 128 registers
// (2 CTAs/SM so one CTA's TMA load overlaps the other's math).  The gate variant loads the matching 64
// channels of the second chunk(2) half with a second box and emits gelu(dw(x1)) * dw(x2).
#include <algorithm>
#include <cstdlib>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int DW_TH = 8, DW_CB = 64;
constexpr int DW_IN_H = DW_TH + 2;
// PX = consecutive output pixels per thread; the tile is 4*PX pixels wide (PX=8: 32x8 tile, PX=4: 16x8 tile)
template <int PX> struct DwCfg {
  static constexpr int TW = 4 * PX, IN_W = TW + 2;
  static constexpr uint32_t TILE_BYTES = IN_W * DW_IN_H * DW_CB * 2;
};

template <int GATE, int PX>
__global__ void __launch_bounds__(256, PX == 8 ? (GATE ? 1 : 2) : (GATE ? 2 : 3))
k_dwconv_tma(const __grid_constant__ CUtensorMap map, bf16* __restrict__ out, long ldo, const float* __restrict__ w9c,
             int H, int W, int C, int Cout, int tiles_x, int tiles_y, int cblocks, int total_tiles) {
  constexpr int DW_TW = DwCfg<PX>::TW, DW_IN_W = DwCfg<PX>::IN_W;
  constexpr uint32_t DW_TILE_BYTES = DwCfg<PX>::TILE_BYTES;
  constexpr uint32_t STAGE = (GATE ? 2 : 1) * DW_TILE_BYTES;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + 2 * STAGE;   // two "full" barriers, one per stage
  const int hp = C / 2;

  auto issue = [&](int tile, int stage) {   // one elected thread: arm the barrier and launch the TMA box(es)
    int b = tile;
    const int cb = b % cblocks; b /= cblocks;
    const int txi = b % tiles_x; b /= tiles_x;
    const int tyi = b % tiles_y;
    const int img = b / tiles_y;
    const uint32_t bar = bar0 + 8u * stage, dst = sbase + stage * STAGE;
    mbar_expect_tx(bar, STAGE);
    tma_load_4d(dst, &map, bar, cb * DW_CB, txi * DW_TW - 1, tyi * DW_TH - 1, img);
    if (GATE) tma_load_4d(dst + DW_TILE_BYTES, &map, bar, hp + cb * DW_CB, txi * DW_TW - 1, tyi * DW_TH - 1, img);
  };

  auto prefetch_l2 = [&](int tile) {   // bring a later tile's boxes into L2 (2-3 tiles ahead of the smem load)
    int b = tile;
    const int cb = b % cblocks; b /= cblocks;
    const int txi = b % tiles_x; b /= tiles_x;
    const int tyi = b % tiles_y;
    const int img = b / tiles_y;
    tma_prefetch_4d(&map, cb * DW_CB, txi * DW_TW - 1, tyi * DW_TH - 1, img);
    if (GATE) tma_prefetch_4d(&map, hp + cb * DW_CB, txi * DW_TW - 1, tyi * DW_TH - 1, img);
  };

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < total_tiles) issue(blockIdx.x, 0);
    if ((long)blockIdx.x + 2L * gridDim.x < total_tiles) prefetch_l2(blockIdx.x + 2 * gridDim.x);
  }
  __syncthreads();

  const int cg = threadIdx.x & 7;
  const int pg = threadIdx.x >> 3;
  const int row = pg >> 2, xs = (pg & 3) * PX;

  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    if (threadIdx.x == 0) {
      if (tile + (int)gridDim.x < total_tiles) issue(tile + gridDim.x, stage ^ 1);
      if ((long)tile + 3L * gridDim.x < total_tiles) prefetch_l2(tile + 3 * gridDim.x);
    }
    int b = tile;
    const int cb = b % cblocks; b /= cblocks;
    const int txi = b % tiles_x; b /= tiles_x;
    const int tyi = b % tiles_y;
    const int img = b / tiles_y;
    const int x0 = txi * DW_TW, y0 = tyi * DW_TH;
    const int ch = cb * DW_CB + cg * 8;    // first of this thread's 8 output channels
    const bool ch_ok = ch < Cout;

    mbar_wait(bar0 + 8u * stage, (it >> 1) & 1);

    float res[PX][8];
#pragma unroll
    for (int half = 0; half < (GATE ? 2 : 1); ++half) {
      float acc[PX][8];
#pragma unroll
      for (int p = 0; p < PX; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = 0.f;
      const bf16* tl = reinterpret_cast<const bf16*>(sgen + stage * STAGE + half * DW_TILE_BYTES);
      const int wch = (GATE && half) ? hp + ch : ch;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        float wt[3][8];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          if (ch_ok) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(w9c + (long)(dy * 3 + t) * C + wch));
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(w9c + (long)(dy * 3 + t) * C + wch + 4));
            wt[t][0] = a.x; wt[t][1] = a.y; wt[t][2] = a.z; wt[t][3] = a.w;
            wt[t][4] = c4.x; wt[t][5] = c4.y; wt[t][6] = c4.z; wt[t][7] = c4.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) wt[t][i] = 0.f;
          }
        }
        const bf16* rp = tl + ((row + dy) * DW_IN_W + xs) * DW_CB + cg * 8;
#pragma unroll
        for (int cx = 0; cx < PX + 2; ++cx) {
          float v[8];
          load8<bf16>(rp + cx * DW_CB, v);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            const int p = cx - t;   // output pixel fed by this input column through filter column t
            if (p >= 0 && p < PX) {
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(v[i], wt[t][i], acc[p][i]);
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
    const int y = y0 + row;
    if (ch_ok && y < H) {
      bf16* op = out + (((long)img * H + y) * W + x0 + xs) * ldo + ch;
#pragma unroll
      for (int p = 0; p < PX; ++p)
        if (x0 + xs + p < W) store8<bf16>(op + (long)p * ldo, res[p]);
    }
    __syncthreads();   // everyone is done reading this stage before it is refilled two iterations later
  }
}

}  // namespace

static int g_dw_sms = 148;
static int g_dw_px = 8;

template <int GATE, int PX>
static int launch_dw(const bf16* x, long ldx, bf16* out, long ldo, const float* w9c, int nimg, int H, int W, int C, cudaStream_t s) {
  constexpr uint32_t smem = 2 * (GATE ? 2 : 1) * DwCfg<PX>::TILE_BYTES + 256;
  static bool attr = false;
  if (!attr) {
    KD_CUDA(cudaFuncSetAttribute(k_dwconv_tma<GATE, PX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  CUtensorMap map;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
  const cuuint64_t str[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)ldx * 2 * W, (cuuint64_t)ldx * 2 * W * H};
  const cuuint32_t box[4] = {DW_CB, (cuuint32_t)DwCfg<PX>::IN_W, DW_IN_H, 1};
  KD_TRY(make_map(&map, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE));
  const int Cout = GATE ? C / 2 : C;
  const int tiles_x = cdiv(W, DwCfg<PX>::TW), tiles_y = cdiv(H, DW_TH), cblocks = cdiv(Cout, DW_CB);
  const long total = (long)nimg * tiles_x * tiles_y * cblocks;
  KD_CHECK(total < (1L << 31), "dwconv3x3: too many tiles");
  const int per_sm = PX == 8 ? (GATE ? 1 : 2) : (GATE ? 2 : 3);
  const long grid = std::min<long>(total, (long)g_dw_sms * per_sm);
  ProfScope prof(PC_DWCONV, s, 18.0 * nimg * H * W * C, (double)nimg * H * W * (C + Cout) * 2.0 + 36.0 * C);
  k_dwconv_tma<GATE, PX><<<(unsigned)grid, 256, smem, s>>>(map, out, ldo, w9c, H, W, C, Cout, tiles_x, tiles_y, cblocks, (int)total);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

// bf16 specialisation entry used by dwconv3x3<bf16>; returns -1 when the shape is not eligible
int dwconv3x3_tma(const bf16* x, long ldx, bf16* out, long ldo, const float* w9c, int nimg, int H, int W, int C, int gate,
                  cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) || ldx % 8 || ldo % 8 || C % (gate ? 16 : 8)) return -1;
  static bool init = false;
  if (!init) {
    int dev = 0;
    KD_CUDA(cudaGetDevice(&dev));
    KD_CUDA(cudaDeviceGetAttribute(&g_dw_sms, cudaDevAttrMultiProcessorCount, dev));
    g_dw_sms = sm_limit(g_dw_sms);
    const char* e = getenv("KDLAE_DW_PX");
    if (e && atoi(e) == 4) g_dw_px = 4;
    if (e && atoi(e) == 8) g_dw_px = 8;
    init = true;
  }
  if (g_dw_px == 4)
    return gate ? launch_dw<1, 4>(x, ldx, out, ldo, w9c, nimg, H, W, C, s) : launch_dw<0, 4>(x, ldx, out, ldo, w9c, nimg, H, W, C, s);
  return gate ? launch_dw<1, 8>(x, ldx, out, ldo, w9c, nimg, H, W, C, s) : launch_dw<0, 8>(x, ldx, out, ldo, w9c, nimg, H, W, C, s);
}

}  // namespace kd
