// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation).
//
//   out[pixel, n] = epilogue( sum_{tap, c} A[shift(pixel, tap), c] * W[n, tap, c] )
//
// * A (NHWC activations) is staged by TMA into 128B-swizzled K-major smem tiles of 128 pixels x 64
//   channels.  1x1 convs use a linear-row tensor map; 3x3 convs use a 4-D (C, W, H, image) map whose
//   box is a 16x8 pixel patch shifted by the tap offset - the conv's zero padding is TMA's
//   out-of-bounds fill, so there is no im2col buffer.  Channel tails (C % 64 != 0) and the channel
//   concat of two sources (skip connections) are also handled by TMA zero fill / two maps.
// * W tiles ([n-chunk <= 256] x 64, K-major) come from a 3-D map (k, n, group); group = image for
//   the per-image attention matrices (MDTA folded into project_out).
// * One elected thread issues tcgen05.mma (M=128, N=n-chunk, K=16) into one of two TMEM accumulator
//   buffers (2 x 256 columns); 8 epilogue warps drain the other buffer with tcgen05.ld.
// * FAST epilogue (identity addressing: every 1x1 and the ASDQE 3x3 convs): the output tile leaves through a
//   ring of four 128 x 64 bf16 shared-memory slabs (128B swizzle) and TMA bulk stores, and the residual tile
//   arrives in the same slab by TMA load and is updated in place - so every byte that crosses HBM moves in
//   full 128-byte lines and the epilogue warps never form a global address.
//   Generic epilogue (PixelShuffle / PixelUnshuffle scatter, ragged N, WithBias LayerNorm): direct stores.
// * Persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer + TMEM
//   allocator, warps 2..9 epilogue, warp 10 residual-slab producer, warp 11 slab store issuer.
//   mbarrier pipelines: smem ring full/empty, TMEM full/empty, slab full/empty (+ named barriers per slab).
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int TC_BM = 128;         // pixels per tile (UMMA M)
constexpr int TC_BK = 64;          // K elements per stage (one 128B swizzle atom of bf16)
constexpr int TC_NC_MAX = 256;     // max N per accumulator (UMMA N)
constexpr int TC_STAGES = 3;        // ring depth with a full-width (256-column) weight tile per stage
constexpr int TC_MAX_STAGES = 8;    // narrower weight tiles shrink the stage and deepen the ring inside the same bytes
constexpr int TC_NSLAB = 4;        // output slabs of 128 rows x 64 columns
constexpr int TC_TW = 16, TC_TH = 8;   // spatial tile of the 3x3 path
constexpr int TC_EPI_WARPS = 16;      // 4 TMEM lane quarters x 4 column quarters of a 64-column slab
constexpr int TC_THREADS = (2 + TC_EPI_WARPS + 2) * 32;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr uint32_t TC_B_BYTES = TC_NC_MAX * TC_BK * 2;
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr uint32_t TC_SLAB_BYTES = TC_BM * 64 * 2;
constexpr uint32_t TC_STAT_BYTES = TC_NSLAB * 4 * TC_BM * 2 * 4;   // [slab buffer][column quarter][row]{sum, sum of squares}
constexpr uint32_t TC_RING_BYTES = TC_STAGES * TC_STAGE_BYTES;
constexpr uint32_t TC_SMEM_BYTES = TC_RING_BYTES + TC_NSLAB * TC_SLAB_BYTES + TC_STAT_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_STORE_BAR_THREADS = TC_EPI_WARPS * 32 + 32;

struct TcParams {
  int spatial;          // 0: linear rows (1x1), 1: 16x8 patches with 3x3 (x3) taps
  int taps, kw, dil;
  int kd, D;            // 3-D convs (Conv3d 3x3x3): temporal taps and frames per batch element (5-D A tensor map)
  float inv_D;
  int kc0, kc1;         // 64-wide K chunks per tap from source 0 / 1
  int c0;               // weight column offset of source 1
  long w_tap_ld;
  int n_chunks, nc;     // N split
  int stages;           // ring depth: TC_RING_BYTES / (A tile + nc weight rows), at most TC_MAX_STAGES.  The write- and
                        // read-heavy 1x1 GEMMs are HBM bound and three 16 KB A tiles in flight per SM do not cover the latency
  uint32_t stage_bytes;
  long items;           // tiles_m * n_chunks
  // linear mode
  long rows_per_group;
  int tiles_per_group;
  // spatial mode
  int H, W, tiles_x, tiles_y;
  int has_res;
  float inv_n_chunks, inv_tiles_per_group, inv_tiles_x, inv_tiles_y;   // fast_div reciprocals (items < 2^24)
  Epilogue epi;
};

struct TileCoord { int g, r0, tx0, ty0, img, nchunk; };

__device__ __forceinline__ TileCoord tile_coord(const TcParams& p, long item64) {
  TileCoord t;
  const int item = (int)item64;
  const int mt = fast_div(item, p.n_chunks, p.inv_n_chunks);
  t.nchunk = item - mt * p.n_chunks;
  t.g = 0; t.r0 = 0; t.tx0 = 0; t.ty0 = 0; t.img = 0;
  if (p.spatial) {
    const int rowt = fast_div(mt, p.tiles_x, p.inv_tiles_x);
    const int txi = mt - rowt * p.tiles_x;
    t.img = fast_div(rowt, p.tiles_y, p.inv_tiles_y);
    const int tyi = rowt - t.img * p.tiles_y;
    t.tx0 = txi * TC_TW; t.ty0 = tyi * TC_TH;
  } else {
    t.g = fast_div(mt, p.tiles_per_group, p.inv_tiles_per_group);
    t.r0 = (mt - t.g * p.tiles_per_group) * TC_BM;
  }
  return t;
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------
// FAST epilogue feature mask (compile-time so the per-vector inner loop has no runtime branches)
enum { EF_ROW_SCALE = 1, EF_BIAS = 2, EF_RES = 4, EF_RELU = 8, EF_STATS = 16, EF_RUNTIME = 32 };

// FAST = 0: generic epilogue. FAST = 1: slab/TMA epilogue specialised on MASK (EF_RUNTIME: flags read at run time).
template <int FAST, int MASK>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_conv_gemm_tc(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
               const __grid_constant__ CUtensorMap map_res, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slab_base = smem_base + TC_RING_BYTES;
  const uint32_t stat_base = slab_base + TC_NSLAB * TC_SLAB_BYTES;
  const uint32_t bar_base = stat_base + TC_STAT_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TC_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_MAX_STAGES + 2 + a); };
  auto sfull_bar = [&](int b) { return bar_base + 8u * (2 * TC_MAX_STAGES + 4 + b); };
  auto sempty_bar = [&](int b) { return bar_base + 8u * (2 * TC_MAX_STAGES + 4 + TC_NSLAB + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_MAX_STAGES + 4 + 2 * TC_NSLAB);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* slab_gen = smem_raw + (slab_base - smem_u32(smem_raw));
  float2* stat_gen = reinterpret_cast<float2*>(smem_raw + (stat_base - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.kc1 > 0) prefetch_tmap(&map_a1);
    prefetch_tmap(&map_w);
    if (FAST) { prefetch_tmap(&map_out); if (p.has_res) prefetch_tmap(&map_res); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TC_EPI_WARPS); }
    for (int b = 0; b < TC_NSLAB; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(sempty_bar(b), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int kchunks = p.kc0 + p.kc1;
  const int kblocks = p.taps * kchunks;
  const uint32_t stage_tx = TC_A_BYTES + (uint32_t)p.nc * TC_BK * 2;
  const int nslabs = (p.nc + 63) / 64;

  if (warp == 0) {
    // ===================== TMA producer (A and W tiles): warp-uniform loop, elected issuer (see elect_one()) =====================
    {
      int s = 0; uint32_t ph = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const TileCoord t = tile_coord(p, item);
        int fb = 0, fd = 0;                       // 3-D: batch element and frame of this tile's image index
        if (p.kd == 3) { fb = fast_div(t.img, p.D, p.inv_D); fd = t.img - fb * p.D; }
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dx = (tap % p.kw - p.kw / 2) * p.dil, dy = ((tap / p.kw) % 3 - p.kw / 2) * p.dil;
          const int dd = (p.kd == 3) ? tap / 9 - 1 : 0;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait_relaxed(empty_bar(s), ph ^ 1);
            const uint32_t a_dst = smem_base + s * p.stage_bytes;
            const uint32_t b_dst = a_dst + TC_A_BYTES;
            const bool src1 = kc >= p.kc0;
            const CUtensorMap* ma = src1 ? &map_a1 : &map_a0;
            const int cc = (src1 ? kc - p.kc0 : kc) * TC_BK;
            const int wk = (int)(tap * p.w_tap_ld) + (src1 ? p.c0 : 0) + cc;
            if (elect_one()) {
              mbar_expect_tx(full_bar(s), stage_tx);
              if (p.kd == 3) tma_load_5d(a_dst, ma, full_bar(s), cc, t.tx0 + dx, t.ty0 + dy, fd + dd, fb);
              else if (p.spatial) tma_load_4d(a_dst, ma, full_bar(s), cc, t.tx0 + dx, t.ty0 + dy, t.img);
              else tma_load_3d(a_dst, ma, full_bar(s), cc, t.r0, t.g);
              tma_load_3d(b_dst, &map_w, full_bar(s), wk, t.nchunk * p.nc, t.g);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loop, one elected lane issues =====================
    {
      const uint32_t idesc = make_idesc(p.nc);
      uint32_t it = 0;
      int s = 0; uint32_t ph = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_NC_MAX;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * p.stage_bytes;
          const uint32_t b_addr = a_addr + TC_A_BYTES;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) {
              const uint64_t ad = make_smem_desc(a_addr + k * 32);
              const uint64_t bd = make_smem_desc(b_addr + k * 32);
              umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(s));      // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tfull_bar(acc));      // accumulator complete
        __syncwarp();
      }
    }
  } else if (warp < 2 + TC_EPI_WARPS) {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int cq = ew >> 2;               // which 16-column quarter of a 64-column slab / which column groups
    const int r = quarter * 32 + lane;    // accumulator row (pixel within the tile)
    uint32_t it = 0, slab_ctr = 0;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      TileCoord t = tile_coord(p, item);
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      long prow; int y = 0, x = 0; bool valid;
      if (p.spatial) {
        x = t.tx0 + (r % TC_TW); y = t.ty0 + (r / TC_TW);
        valid = (x < p.W) && (y < p.H);
        prow = ((long)t.img * p.H + y) * p.W + x;
      } else {
        const long rr = (long)t.r0 + r;
        valid = rr < p.rows_per_group;
        prow = (long)t.g * p.rows_per_group + rr;
        if (p.epi.mode == OUT_PLANAR_F32) {     // 1x1 to fp32 planes: (image, pixel) from the linear row; epilogue uses y*W + x
          const long hw = (long)p.H * p.W;
          t.img = (int)(prow / hw);
          x = (int)(prow - (long)t.img * hw); y = 0;
        }
      }
      const uint32_t t_row = tmem_base + acc * TC_NC_MAX + ((uint32_t)(quarter * 32) << 16);
      const int nbase = t.nchunk * p.nc;
      if (FAST) {
        const bool rt = (MASK & EF_RUNTIME) != 0;
        const bool f_rs = rt ? (p.epi.row_scale != nullptr) : (MASK & EF_ROW_SCALE) != 0;
        const bool f_bias = rt ? (p.epi.col_bias != nullptr) : (MASK & EF_BIAS) != 0;
        const bool f_res = rt ? (p.has_res != 0) : (MASK & EF_RES) != 0;
        const bool f_relu = rt ? (p.epi.relu != 0) : (MASK & EF_RELU) != 0;
        const bool f_stats = rt ? (p.epi.stat_rstd != nullptr) : (MASK & EF_STATS) != 0;
        const float* __restrict__ bias_p = p.epi.col_bias;
        const int n_valid = p.epi.N, nc_ = p.nc;
        const float rs = (f_rs && valid) ? __ldg(p.epi.row_scale + prow) : 1.0f;
        float st_mean = 0.f, st_m2 = 0.f;   // running sum / sum of squares of this thread's part of the row
        mbar_wait_relaxed(tfull_bar(acc), aph);
        tc_fence_after();
        for (int j = 0; j < nslabs; ++j, ++slab_ctr) {
          const int b = slab_ctr % TC_NSLAB;
          const uint32_t sph = (slab_ctr / TC_NSLAB) & 1;
          // the slab holds the residual tile (TMA-loaded) or is simply free again (its last store has been read out)
          if (f_res) mbar_wait_relaxed(sfull_bar(b), sph);
          else mbar_wait_relaxed(sempty_bar(b), sph ^ 1);
          const int col0 = j * 64 + cq * 16;
          if (col0 < nc_) {
            uint32_t v[16];
            tmem_ld16_issue(t_row + col0, v);
            tmem_ld16_wait(v);
            uint8_t* srow = slab_gen + b * TC_SLAB_BYTES + r * 128;
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
              const int chunk = cq * 2 + q4;                   // 16-byte chunk of the 128-byte slab row
              const int n = nbase + j * 64 + chunk * 8;
              uint4* sp = reinterpret_cast<uint4*>(srow + ((chunk ^ (r & 7)) << 4));   // 128B swizzle
              if (n < n_valid) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = f_rs ? __uint_as_float(v[q4 * 8 + i]) * rs : __uint_as_float(v[q4 * 8 + i]);
                if (f_bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_p + n));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_p + n + 4));
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (f_res) {
                  const uint4 rr4 = *sp;
                  const uint32_t w4[4] = {rr4.x, rr4.y, rr4.z, rr4.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    f[2 * i] += __uint_as_float(w4[i] << 16);
                    f[2 * i + 1] += __uint_as_float(w4[i] & 0xffff0000u);
                  }
                }
                if (f_relu) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
                }
                if (f_stats) {   // plain sums; the residual stream is roughly centred, fp32 is ample for C <= 256
#pragma unroll
                  for (int i = 0; i < 8; ++i) { st_mean += f[i]; st_m2 = fmaf(f[i], f[i], st_m2); }
                }
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                *sp = o;
              }
            }
          }
          if (f_stats && j == nslabs - 1) stat_gen[(b * 4 + cq) * TC_BM + r] = make_float2(st_mean, st_m2);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the TMA store
          if (j == nslabs - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));                 // TMEM buffer drained
          }
          named_bar_arrive(1 + b, TC_STORE_BAR_THREADS);                 // slab written (store warp syncs on it)
        }
      } else {
        mbar_wait(tfull_bar(acc), aph);
        tc_fence_after();
        const int ngroups = (p.nc + 31) / 32;
        for (int cgp = cq; cgp < ngroups; cgp += 4) {
          uint32_t v[32];
          tmem_ld32(t_row + cgp * 32, v);
          if (valid) {
#pragma unroll 1
            for (int q4 = 0; q4 < 4; ++q4) {
              const int col = cgp * 32 + q4 * 8;
              if (col < p.nc) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[q4 * 8 + i]);
                epilogue_store8<bf16>(p.epi, prow, t.img, y, x, nbase + col, f);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
    }
  } else if (warp == 2 + TC_EPI_WARPS) {
    // ===================== residual-slab producer =====================
    if (FAST && p.has_res && lane == 0) {
      uint32_t slab_ctr = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const TileCoord t = tile_coord(p, item);
        for (int j = 0; j < nslabs; ++j, ++slab_ctr) {
          const int b = slab_ctr % TC_NSLAB;
          mbar_wait_relaxed(sempty_bar(b), ((slab_ctr / TC_NSLAB) & 1) ^ 1);
          mbar_expect_tx(sfull_bar(b), TC_SLAB_BYTES);
          const int n0 = t.nchunk * p.nc + j * 64;
          if (p.spatial) tma_load_4d(slab_base + b * TC_SLAB_BYTES, &map_res, sfull_bar(b), n0, t.tx0, t.ty0, t.img);
          else tma_load_3d(slab_base + b * TC_SLAB_BYTES, &map_res, sfull_bar(b), n0, t.r0, t.g);
        }
      }
    }
  } else {
    // ===================== slab store issuer =====================
    if (FAST) {
      uint32_t slab_ctr = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const TileCoord t = tile_coord(p, item);
        for (int j = 0; j < nslabs; ++j, ++slab_ctr) {
          const int b = slab_ctr % TC_NSLAB;
          named_bar_sync(1 + b, TC_STORE_BAR_THREADS);                   // all 256 epilogue threads have written slab b
          if (p.epi.stat_rstd && j == nslabs - 1) {
            // add the two column halves of every row and publish rstd / mean for the next LayerNorm
            const float inv_n = 1.0f / (float)p.epi.N;
#pragma unroll
            for (int k = 0; k < TC_BM / 32; ++k) {
              const int row = k * 32 + lane;
              const long rr = (long)t.r0 + row;
              if (rr < p.rows_per_group) {
                float sx = 0.f, sq = 0.f;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) { const float2 a = stat_gen[(b * 4 + c4) * TC_BM + row]; sx += a.x; sq += a.y; }
                const float mean = sx * inv_n;
                const float var = fmaxf(sq * inv_n - mean * mean, 0.f);
                const long prow = (long)t.g * p.rows_per_group + rr;
                p.epi.stat_rstd[prow] = rsqrtf(var + 1e-5f);
                if (p.epi.stat_mu) p.epi.stat_mu[prow] = mean;
              }
            }
          }
          if (lane == 0) {
            const int n0 = t.nchunk * p.nc + j * 64;
            if (p.spatial) tma_store_4d(&map_out, slab_base + b * TC_SLAB_BYTES, n0, t.tx0, t.ty0, t.img);
            else tma_store_3d(&map_out, slab_base + b * TC_SLAB_BYTES, n0, t.r0, t.g);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");    // stores <= slab_ctr-2 have left smem
            if (slab_ctr >= 2) mbar_arrive(sempty_bar((slab_ctr - 2) % TC_NSLAB));
          }
          __syncwarp();
        }
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---------------------------------------------------------------------------------------
// Host side: tensor maps + launch
// ---------------------------------------------------------------------------------------
// N is split in chunks of at most 256 accumulator columns; with several chunks the chunk width is a multiple of 64 so
// that no 64-column output slab straddles two chunks.
int pick_nc(int N, int* n_chunks) {
  const int n16 = (N + 15) / 16 * 16;
  if (n16 <= TC_NC_MAX) { *n_chunks = 1; return n16; }
  int best_nc = 256, best_cost = 1 << 30;
  for (int nc = 256; nc >= 128; nc -= 64) {
    const int chunks = (N + nc - 1) / nc;
    const int cost = chunks * nc;
    if (cost < best_cost) { best_cost = cost; best_nc = nc; }
  }
  *n_chunks = (N + best_nc - 1) / best_nc;
  return best_nc;
}

// KDLAE_CONV3X3_LEGACY=1 keeps 3x3 convs on this file's tap-by-tap path (A/B comparison in scripts/time_models.py)
bool conv3x3_legacy() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("KDLAE_CONV3X3_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

cudaError_t set_smem_attr() {
  cudaError_t e = cudaSuccess;
#define KD_TC_ATTR(F, M) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_conv_gemm_tc<F, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  KD_TC_ATTR(0, 0) KD_TC_ATTR(1, 0) KD_TC_ATTR(1, EF_ROW_SCALE) KD_TC_ATTR(1, EF_RES) KD_TC_ATTR(1, EF_RES | EF_STATS)
  KD_TC_ATTR(1, EF_BIAS) KD_TC_ATTR(1, EF_BIAS | EF_RELU) KD_TC_ATTR(1, EF_RUNTIME)
#undef KD_TC_ATTR
  return e;
}

}  // namespace

bool conv_gemm_tc_eligible(const ConvOp& op) {
  const bool k11 = (op.kd == 1 && op.kh == 1 && op.kw == 1);
  const bool k33 = (op.kd == 1 && op.kh == 3 && op.kw == 3);
  const bool k333 = (op.kd == 3 && op.kh == 3 && op.kw == 3 && op.dil == 1 && op.c1 == 0 && op.D >= 1 && op.nimg % op.D == 0);
  if (!k11 && !k33 && !k333) return false;
  if (op.c0 % 8 || op.ld0 % 8 || (reinterpret_cast<uintptr_t>(op.a0) & 15)) return false;
  if (op.c1 > 0 && (op.c1 % 8 || op.ld1 % 8 || (reinterpret_cast<uintptr_t>(op.a1) & 15))) return false;
  if (op.w_ld % 8 || op.w_tap_ld % 8 || (reinterpret_cast<uintptr_t>(op.w) & 15) || op.w_group_stride % 8) return false;
  if ((k33 || k333) && op.groups != 1) return false;
  // 1x1 with PixelShuffle addressing (ConvTranspose 2x2 stride 2 of KDLAE-S = 1x1 GEMM with 4 phases): runs on the spatial
  // (y, x)-tiled path with a single tap, generic epilogue
  if (k11 && op.epi.mode == OUT_PIXEL_UNSHUFFLE) return false;
  if (k11 && op.epi.mode == OUT_PIXEL_SHUFFLE && (op.groups != 1 || op.epi.row_scale != nullptr)) return false;
  if (op.epi.mode == OUT_PIXEL_SHUFFLE && (op.epi.cq % 8 != 0)) return false;
  return true;
}

int conv_gemm_tc(const ConvOp& op, cudaStream_t s) {
  KD_CHECK(conv_gemm_tc_eligible(op), "conv_gemm_tc: shape not eligible");
  static DeviceOnce once;
  bool first; int dev, g_num_sms;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(set_smem_attr());
    device_mark(once, dev);
  }
  KD_TRY(device_sms(&g_num_sms));
  if (op.kh == 3 && op.dil == 1 && op.epi.row_scale == nullptr && op.epi.row_mu == nullptr && op.epi.stat_rstd == nullptr &&
      !conv3x3_legacy()) {
    int n_chunks = 1;
    const int nc = pick_nc(op.epi.N, &n_chunks);
    return conv3x3_tc(op, nc, n_chunks, s);     // one activation load per tile, taps = shifted descriptors (conv3x3_tc.cu)
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.spatial = (op.kh == 3 || op.epi.mode == OUT_PIXEL_SHUFFLE || op.epi.out_y_ld != 0) ? 1 : 0;
  p.taps = op.kd * op.kh * op.kw;
  p.kw = op.kw;
  p.kd = op.kd; p.D = op.D; p.inv_D = 1.0f / (float)(op.D > 0 ? op.D : 1);
  p.dil = op.dil;
  p.kc0 = (op.c0 + TC_BK - 1) / TC_BK;
  p.kc1 = (op.c1 + TC_BK - 1) / TC_BK;
  p.c0 = op.c0;
  p.w_tap_ld = op.w_tap_ld;
  p.nc = pick_nc(op.epi.N, &p.n_chunks);
  p.stage_bytes = TC_A_BYTES + (uint32_t)p.nc * TC_BK * 2;       // nc % 16 == 0: the weight tile keeps the 1024-byte alignment
  p.stages = (int)std::min<uint32_t>(TC_RING_BYTES / p.stage_bytes, TC_MAX_STAGES);
  { const char* e = getenv("KDLAE_TC_STAGES"); if (e && atoi(e) >= 2) p.stages = std::min(p.stages, atoi(e)); }   // A/B probe
  p.H = op.H; p.W = op.W;
  p.epi = op.epi;
  const Epilogue& e = op.epi;
  p.has_res = e.res != nullptr;
  const bool fast = e.mode == OUT_IDENTITY && e.N % 8 == 0 && e.row_mu == nullptr && e.out_ld % 8 == 0 && e.out_coff % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 &&
                    (e.res == nullptr || (e.res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(e.res) & 15) == 0)) &&
                    (e.col_bias == nullptr || (reinterpret_cast<uintptr_t>(e.col_bias) & 15) == 0);

  KD_CHECK(e.out_y_ld == 0 || (fast && op.kh == 1), "conv_gemm_tc: a destination row stride needs the FAST 1x1 epilogue");
  KD_CHECK(e.stat_rstd == nullptr || (fast && p.n_chunks == 1 && !p.spatial),
           "conv_gemm_tc: LayerNorm statistics need the FAST 1x1 epilogue with N in one chunk (N=%d)", e.N);
  CUtensorMap ma0, ma1, mw, mout, mres;
  const long rows = (long)op.nimg * op.H * op.W;
  long tiles_m;
  // activation-like maps (A sources, output, residual): {channels, pixels...}
  auto act_map = [&](CUtensorMap* m, const void* base, int ch, long ld, int box_ch, long y_ld = 0) -> int {
    if (p.spatial) {
      if (y_ld == 0) y_ld = ld * op.W;
      const cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.nimg};
      const cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)y_ld * 2, (cuuint64_t)y_ld * 2 * op.H};
      const cuuint32_t box[4] = {(cuuint32_t)box_ch, TC_TW, TC_TH, 1};
      return make_map(m, base, 4, dims, str, box);
    }
    const cuuint64_t dims[3] = {(cuuint64_t)ch, (cuuint64_t)p.rows_per_group, (cuuint64_t)op.groups};
    const cuuint64_t str[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * p.rows_per_group};
    const cuuint32_t box[3] = {(cuuint32_t)box_ch, TC_BM, 1};
    return make_map(m, base, 3, dims, str, box);
  };
  if (p.spatial) {
    p.tiles_x = cdiv(op.W, TC_TW);
    p.tiles_y = cdiv(op.H, TC_TH);
    tiles_m = (long)op.nimg * p.tiles_x * p.tiles_y;
  } else {
    KD_CHECK(rows % op.groups == 0, "conv_gemm_tc: rows %ld not divisible by groups %d", rows, op.groups);
    p.rows_per_group = rows / op.groups;
    p.tiles_per_group = cdiv(p.rows_per_group, TC_BM);
    tiles_m = (long)p.tiles_per_group * op.groups;
  }
  if (op.kd == 3) {   // 5-D {C, W, H, frames, batch}: the temporal zero padding is TMA out-of-bounds fill as well
    const cuuint64_t dims[5] = {(cuuint64_t)op.c0, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.D, (cuuint64_t)(op.nimg / op.D)};
    const cuuint64_t f = (cuuint64_t)op.ld0 * 2 * op.W * op.H;
    const cuuint64_t str[4] = {(cuuint64_t)op.ld0 * 2, (cuuint64_t)op.ld0 * 2 * op.W, f, f * op.D};
    const cuuint32_t box[5] = {TC_BK, TC_TW, TC_TH, 1, 1};
    KD_TRY(make_map(&ma0, op.a0, 5, dims, str, box));
  } else {
    KD_TRY(act_map(&ma0, op.a0, op.c0, op.ld0, TC_BK));
  }
  if (op.c1 > 0) KD_TRY(act_map(&ma1, op.a1, op.c1, op.ld1, TC_BK));
  else ma1 = ma0;
  if (fast) {
    KD_TRY(act_map(&mout, reinterpret_cast<const bf16*>(e.out) + e.out_coff, e.N, e.out_ld, 64, e.out_y_ld));
    if (e.res) KD_TRY(act_map(&mres, e.res, e.N, e.res_ld, 64, e.res_y_ld));
    else mres = mout;
  } else {
    mout = ma0; mres = ma0;
  }
  {
    const long w_row = (p.taps > 1) ? (long)p.taps * op.w_tap_ld : (long)(op.c0 + op.c1);
    const cuuint64_t dims[3] = {(cuuint64_t)w_row, (cuuint64_t)op.epi.N, (cuuint64_t)op.groups};
    const cuuint64_t str[2] = {(cuuint64_t)op.w_ld * 2,
                               (cuuint64_t)(op.groups > 1 ? op.w_group_stride : (long)op.w_ld * op.epi.N) * 2};
    const cuuint32_t box[3] = {TC_BK, (cuuint32_t)p.nc, 1};
    KD_TRY(make_map(&mw, op.w, 3, dims, str, box));
  }
  p.items = tiles_m * p.n_chunks;
  KD_CHECK(p.items < (1L << 24), "conv_gemm_tc: too many tiles (%ld)", p.items);
  p.inv_n_chunks = 1.0f / (float)p.n_chunks;
  p.inv_tiles_per_group = p.tiles_per_group ? 1.0f / (float)p.tiles_per_group : 0.f;
  p.inv_tiles_x = p.tiles_x ? 1.0f / (float)p.tiles_x : 0.f;
  p.inv_tiles_y = p.tiles_y ? 1.0f / (float)p.tiles_y : 0.f;
  const int grid = (int)(p.items < (long)g_num_sms ? p.items : (long)g_num_sms);
  const double ktot = (double)p.taps * (op.c0 + op.c1);
  ProfScope prof(PC_GEMM_TC, s, 2.0 * rows * op.epi.N * ktot,
                 2.0 * ((double)rows * (op.c0 + op.c1 + op.epi.N * (op.epi.res ? 2 : 1)) + (double)op.groups * op.epi.N * ktot));
  if (!fast) {
    k_conv_gemm_tc<0, 0><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(ma0, ma1, mw, mout, mres, p);
  } else {
    const int mask = (e.row_scale ? EF_ROW_SCALE : 0) | (e.col_bias ? EF_BIAS : 0) | (e.res ? EF_RES : 0) | (e.relu ? EF_RELU : 0) |
                     (e.stat_rstd ? EF_STATS : 0);
#define KD_TC_CASE(M) case M: k_conv_gemm_tc<1, M><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(ma0, ma1, mw, mout, mres, p); break;
    switch (mask) {
      KD_TC_CASE(0)                                  // reduce_chan
      KD_TC_CASE(EF_ROW_SCALE)                       // qkv, project_in (LayerNorm folded)
      KD_TC_CASE(EF_RES)                             // attention apply / project_out + residual
      KD_TC_CASE(EF_RES | EF_STATS)                  //   ... + statistics for the next LayerNorm
      KD_TC_CASE(EF_BIAS)                            // ASDQE outc
      KD_TC_CASE(EF_BIAS | EF_RELU)                  // ASDQE conv + BN + ReLU
      default: k_conv_gemm_tc<1, EF_RUNTIME><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(ma0, ma1, mw, mout, mres, p); break;
    }
#undef KD_TC_CASE
  }
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
