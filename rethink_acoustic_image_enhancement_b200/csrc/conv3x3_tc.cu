// Dense 3x3 / 3x3x3 convolution on tcgen05 with ONE activation load per (tile, K chunk[, frame]) - bf16 path.
//   Downsample/Upsample convs of KDLAE-T (KDLAE_model.py:186,196), output convs (:258,:261,:268), every Conv3d of KDLAE-S
//   (:389-392) and every Conv3x3+BN+ReLU of ASDQE (ASDQE_model.py:24-31).
// gemm_tc.cu's spatial mode re-loaded the activation tile for each of the 9 (27) taps.  Here the halo tile
// {64 ch, 32 px, 6 rows} is brought in once (TMA, 128B swizzle, row pitch exactly 32 pixels, zero padding = OOB fill) and
// the 9 spatial taps are 9 shifted shared-memory descriptors over it: start address += (dy*32 + dx) * 128 bytes (the
// swizzle XOR follows the absolute address, verified on B200 in round 1).  Only the weight tiles stream per tap.
// A tile yields 4 rows x 30 pixels (columns 30,31 of each row are garbage and never stored).
// Two smem rings: A (2 slots, each consumed by 9 taps) and W (3 slots, one per tap).  TMEM double buffered, epilogues as
// in gemm_tc.cu: FAST = swizzled slabs + TMA stores/loads (bias, ReLU, residual), generic = epilogue_store8
// (PixelShuffle / PixelUnshuffle / planar fp32).  Warps: 0 producer, 1 MMA, 2..17 epilogue, 18 residual loads, 19 stores.
#include <algorithm>
#include "sm100.cuh"

namespace kd {

namespace {

constexpr int C3_TW = 32, C3_OW = 30, C3_OH = 4, C3_IH = 6;
constexpr int C3_NC_MAX = 256;
constexpr int C3_ASLOTS = 2, C3_ASLOTS_MAX = 4, C3_BSLOTS = 3, C3_BSLOTS_MAX = 8, C3_NSLAB = 4;   // A ring: 2 slots, up to 4 in weight-stationary mode
constexpr uint32_t C3_A_BYTES = C3_TW * C3_IH * 128;          // 24576
constexpr uint32_t C3_A_SLOT = C3_A_BYTES + 1024;             // shifted reads run 2 pixels past the tile
constexpr uint32_t C3_B_SLOT = C3_NC_MAX * 128;               // 32768
constexpr uint32_t C3_SLAB = 15360;                           // 120 px x 128 B
constexpr int C3_EPI_WARPS = 16;
constexpr int C3_THREADS = (2 + C3_EPI_WARPS + 2) * 32;
constexpr uint32_t C3_SMEM = C3_ASLOTS * C3_A_SLOT + C3_BSLOTS * C3_B_SLOT + C3_NSLAB * C3_SLAB + 1024 + 512;   // streaming layout
constexpr uint32_t C3_SMEM_MAX = 227 * 1024;
constexpr int C3_STORE_BAR_THREADS = C3_EPI_WARPS * 32 + 32;
#ifndef KDLAE_C3_RES2
#define KDLAE_C3_RES2 1
#endif

struct C3Params {
  int kd, D;                 // temporal taps (1 or 3) and frames per batch element
  int kc0, kc1, c0, c1;      // 64-wide K chunks of source 0 / 1; channels of source 0 (= weight column offset of source 1) / 1
  long w_tap_ld;
  int n_chunks, nc;
  long items;
  int H, W, tiles_x, tiles_y;
  int has_res, relu;
  int a_slots;               // depth of the activation-tile ring (weight-stationary mode: the freed weight-ring space deepens it)
  uint32_t b_bytes;          // bytes of the weight region
  int b_group, b_slots;      // streaming mode: taps per weight load (3 = one kernel row, or 1) and ring depth
  uint32_t b_slot;           // bytes per weight-ring slot
  int w_resident;            // 1: all 9 (27) tap tiles of the (single) K chunk fit shared memory -> loaded once per CTA
  int w_tiles;               // weight tiles held in resident mode (9 * kd, or xconv_tiles() for an x-packed conv)
  int xpack_cin;             // x-packed narrow conv: channels per pixel (16 / 32), 0 = ordinary conv (see ConvOp::xpack_cin)
  int nslab;                 // output slab ring depth (4; 2 when the resident weights leave no room)
  float inv_n_chunks, inv_tiles_x, inv_tiles_y, inv_D;
  Epilogue epi;
};

__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

struct C3Tile { int img, y0, x0, nchunk; };
__device__ __forceinline__ C3Tile c3_tile(const C3Params& p, long item64) {
  C3Tile t;
  const int item = (int)item64;
  const int mt = fast_div(item, p.n_chunks, p.inv_n_chunks);
  t.nchunk = item - mt * p.n_chunks;
  const int rowt = fast_div(mt, p.tiles_x, p.inv_tiles_x);
  const int txi = mt - rowt * p.tiles_x;
  t.img = fast_div(rowt, p.tiles_y, p.inv_tiles_y);
  const int tyi = rowt - t.img * p.tiles_y;
  t.x0 = txi * C3_OW; t.y0 = tyi * C3_OH;
  return t;
}

template <int FAST>
__global__ void __launch_bounds__(C3_THREADS, 1)
k_conv3_tc(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
           const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
           const __grid_constant__ CUtensorMap map_res, const C3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = sbase;
  const uint32_t b_base = a_base + p.a_slots * C3_A_SLOT;
  const uint32_t slab_base = b_base + p.b_bytes;
  const uint32_t bar_base = slab_base + p.nslab * C3_SLAB;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (C3_ASLOTS_MAX + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * C3_ASLOTS_MAX + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * C3_ASLOTS_MAX + C3_BSLOTS_MAX + s); };
  const uint32_t bar2 = bar_base + 8u * (2 * C3_ASLOTS_MAX + 2 * C3_BSLOTS_MAX);
  auto tfull_bar = [&](int a) { return bar2 + 8u * a; };
  auto tempty_bar = [&](int a) { return bar2 + 8u * (2 + a); };
  auto sfull_bar = [&](int b) { return bar2 + 8u * (4 + b); };
  auto sempty_bar = [&](int b) { return bar2 + 8u * (4 + C3_NSLAB + b); };
  const uint32_t tmem_slot = bar2 + 8u * (4 + 2 * C3_NSLAB);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  uint8_t* slab_gen = smem_raw + (slab_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0); prefetch_tmap(&map_w);
    if (p.kc1 > 0) prefetch_tmap(&map_a1);
    if (FAST) { prefetch_tmap(&map_out); if (p.has_res) prefetch_tmap(&map_res); }
    for (int s = 0; s < C3_ASLOTS_MAX; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < C3_BSLOTS_MAX; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), C3_EPI_WARPS); }
    for (int b = 0; b < C3_NSLAB; ++b) { mbar_init(sfull_bar(b), 1); mbar_init(sempty_bar(b), 1); }
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
  const int agroups = kchunks * p.kd;           // activation tiles per output tile: (K chunk, frame tap)
  const int nslabs = (p.nc + 63) / 64;

  if (warp == 0) {
    // ===================== TMA producer: A tiles (one per 9 taps) and W tiles (one per tap) =====================
    // (warp-uniform loop, one elected lane issues - see elect_one())
    {
      uint32_t aidx = 0, bidx = 0;
      if (p.w_resident) {          // weight-stationary: 9 (27) x [nc x 64] tiles, one barrier, no per-tap handshakes afterwards
        if (elect_one()) {
          mbar_expect_tx(b_full(0), (uint32_t)p.w_tiles * (uint32_t)p.nc * 128);
          if (kchunks == 1) {
            for (int tap = 0; tap < p.w_tiles; tap += p.b_group)
              tma_load_3d(b_base + tap * p.nc * 128, &map_w, b_full(0), 0, 0, tap);
          } else {                 // 2-D conv with two K chunks: tiles in [chunk][tap] order, the order the MMA loop walks them
            for (int kc = 0; kc < kchunks; ++kc) {
              const int wk = kc >= p.kc0 ? p.c0 + (kc - p.kc0) * 64 : kc * 64;
              for (int tap = 0; tap < 9; tap += p.b_group)
                tma_load_3d(b_base + (kc * 9 + tap) * p.nc * 128, &map_w, b_full(0), wk, 0, tap);
            }
          }
        }
        __syncwarp();
      }
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const C3Tile t = c3_tile(p, item);
        int fb = 0, fd = 0;
        if (p.kd == 3) { fb = fast_div(t.img, p.D, p.inv_D); fd = t.img - fb * p.D; }
        for (int kc = 0; kc < kchunks; ++kc) {
          const bool src1 = kc >= p.kc0;
          const CUtensorMap* ma = src1 ? &map_a1 : &map_a0;
          const int cc = (src1 ? kc - p.kc0 : kc) * 64;
          for (int td = 0; td < p.kd; ++td, ++aidx) {
            const int s = aidx % p.a_slots;
            mbar_wait_relaxed(a_empty(s), ((aidx / p.a_slots) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(a_full(s), C3_A_BYTES);
              if (p.kd == 3) tma_load_5d(a_base + s * C3_A_SLOT, ma, a_full(s), cc, t.x0 - 1, t.y0 - 1, fd + td - 1, fb);
              else tma_load_4d(a_base + s * C3_A_SLOT, ma, a_full(s), cc, t.x0 - 1, t.y0 - 1, t.img);
            }
            __syncwarp();
            for (int tap = 0; tap < 9 && !p.w_resident; tap += p.b_group, ++bidx) {
              const int bs = bidx % p.b_slots;
              mbar_wait_relaxed(b_empty(bs), ((bidx / p.b_slots) & 1) ^ 1);
              if (elect_one()) {
                mbar_expect_tx(b_full(bs), (uint32_t)p.b_group * p.nc * 128);
                tma_load_3d(b_base + bs * p.b_slot, &map_w, b_full(bs), (src1 ? p.c0 : 0) + cc, t.nchunk * p.nc, td * 9 + tap);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t lo_tag = 1u << 16;
      uint32_t aidx = 0, bidx = 0, it = 0;
      if (p.w_resident) mbar_wait(b_full(0), 0);
      for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t acc = it & 1, aph = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C3_NC_MAX;
        for (int ag = 0; ag < agroups; ++ag, ++aidx) {
          const int s = aidx % p.a_slots;
          mbar_wait(a_full(s), (aidx / p.a_slots) & 1);
          tc_fence_after();
          const uint32_t a_lo0 = (((a_base + s * C3_A_SLOT) & 0x3FFFF) >> 4) | lo_tag;
          // K = 16 steps that hold real channels in this chunk (a 16-channel layer needs 1 of the 4: the rest of the box is TMA
          // zero fill and would only burn tensor-pipe time - the KDLAE-S / ASDQE layers are 16..64 channels wide)
          const int kci = ag / p.kd;
          const int rem = kci < p.kc0 ? p.c0 - kci * 64 : p.c1 - (kci - p.kc0) * 64;
          const int ksn = rem >= 64 ? 4 : (rem + 15) >> 4;
          if (p.xpack_cin) {
            // x-packed narrow conv, weight-stationary: per kernel row ty the centre tile (4 K steps on the super-pixel itself)
            // and the two halo slots (kpc K steps each: the right neighbour's first pixel, the left neighbour's last pixel).
            // A and B descriptors advance independently, so the halo slots are packed densely in their tiles.
            // Every offset is computed here, in warp-uniform code without divisions (single K chunk: ag is the frame tap): the
            // first version did `h / hpt`, `ag % kd` inside the elected region and the issuing thread spent ~2500 instructions per
            // activation tile on address arithmetic - the tensor pipe sat at 30 % waiting for it.
            const uint32_t b_tap = (uint32_t)p.nc * 8;
            const uint32_t b_lo0 = ((b_base & 0x3FFFF) >> 4) | lo_tag;
            const int kpc = p.xpack_cin >> 4;                 // K steps per pixel: 1 (16 channels) or 2 (32 channels)
            const int hsh = (p.xpack_cin == 16) ? 1 : 0;      // (frame, row) taps per halo tile = 1 << hsh
            const int h0 = ag * 3;
            uint32_t a_row[3], b_c[3], b_h[3];
#pragma unroll
            for (int ty = 0; ty < 3; ++ty) {
              const int h = h0 + ty;
              a_row[ty] = a_lo0 + (uint32_t)(ty * C3_TW * 8);
              b_c[ty] = b_lo0 + (uint32_t)h * b_tap;
              b_h[ty] = b_lo0 + (uint32_t)(3 * p.kd + (h >> hsh)) * b_tap + (uint32_t)((h & ((1 << hsh) - 1)) * 4 * kpc);
            }
            const uint32_t first = (ag == 0) ? 0u : 1u;
            if (elect_one()) {
#pragma unroll
              for (int ty = 0; ty < 3; ++ty) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16_lohi(d_tmem, a_row[ty] + 8 + ks * 2, b_c[ty] + ks * 2, desc_hi, idesc, (ty | ks) != 0 ? 1u : first);
                if (kpc == 1) {
                  umma_bf16_lohi(d_tmem, a_row[ty] + 16, b_h[ty], desc_hi, idesc, 1u);            // right neighbour's first pixel
                  umma_bf16_lohi(d_tmem, a_row[ty] + 6, b_h[ty] + 2, desc_hi, idesc, 1u);         // left neighbour's last pixel
                } else {
                  umma_bf16_lohi(d_tmem, a_row[ty] + 16, b_h[ty], desc_hi, idesc, 1u);
                  umma_bf16_lohi(d_tmem, a_row[ty] + 18, b_h[ty] + 2, desc_hi, idesc, 1u);
                  umma_bf16_lohi(d_tmem, a_row[ty] + 4, b_h[ty] + 4, desc_hi, idesc, 1u);
                  umma_bf16_lohi(d_tmem, a_row[ty] + 6, b_h[ty] + 6, desc_hi, idesc, 1u);
                }
              }
            }
            __syncwarp();
          } else if (p.w_resident) {
            // weight-stationary: no per-tap handshake - the 9 x ksn MMAs of this A tile go out back to back from one elected lane
            // (descriptor offsets in warp-uniform code outside the elected region, as in the x-packed branch: inside it ptxas keeps
            // them per thread and wraps every MMA in R2UR moves)
            const uint32_t b_tap = (uint32_t)p.nc * 8;            // (nc * 128 bytes) >> 4 per tap tile
            // the A group index is the frame tap (single K chunk) or the K chunk (2-D conv): its 9 spatial tap tiles start at 9 * ag
            const uint32_t b_lo0 = (((b_base & 0x3FFFF) >> 4) | lo_tag) + (uint32_t)ag * 9u * b_tap;
            uint32_t b_t[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) b_t[tap] = b_lo0 + tap * b_tap;
            const uint32_t first = (ag == 0) ? 0u : 1u;
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t a_lo = a_lo0 + (uint32_t)(((tap / 3) * C3_TW + tap % 3) * 8);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  if (ks < ksn) umma_bf16_lohi(d_tmem, a_lo + ks * 2, b_t[tap] + ks * 2, desc_hi, idesc, (tap | ks) != 0 ? 1u : first);
              }
            }
            __syncwarp();
          } else {
#pragma unroll 1
          for (int tap0 = 0; tap0 < 9; tap0 += p.b_group, ++bidx) {     // one weight load = b_group taps (a kernel row, or one tap)
            const int bs = bidx % p.b_slots;
            mbar_wait(b_full(bs), (bidx / p.b_slots) & 1);
            tc_fence_after();
            // taps tap0 .. tap0 + b_group - 1 share the kernel row: descriptor offsets are computed in warp-uniform code, outside
            // the elected region (inside it ptxas falls back to per-thread values and R2UR wrappers around every MMA)
            const uint32_t a_row = a_lo0 + (uint32_t)(((tap0 / 3) * C3_TW + tap0 % 3) * 8);
            const uint32_t b_lo0 = (((b_base + bs * p.b_slot) & 0x3FFFF) >> 4) | lo_tag;
            const uint32_t b_tap = (uint32_t)p.nc * 8;
            const int ng = p.b_group;
            if (elect_one()) {
#pragma unroll
              for (int tt = 0; tt < 3; ++tt) {
                if (tt < ng) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    if (ks < ksn)
                      umma_bf16_lohi(d_tmem, a_row + tt * 8 + ks * 2, b_lo0 + tt * b_tap + ks * 2, desc_hi, idesc, (ag | (tap0 + tt) | ks) != 0 ? 1u : 0u);
                }
              }
              umma_commit(b_empty(bs));
            }
            __syncwarp();
          }
          }
          if (elect_one()) umma_commit(a_empty(s));
          __syncwarp();
        }
        if (elect_one()) umma_commit(tfull_bar(acc));
        __syncwarp();
      }
    }
  } else if (warp < 2 + C3_EPI_WARPS) {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int cq = ew >> 2;
    const int r = quarter * 32 + lane;
    const int ry = r / C3_TW, rx = r % C3_TW;
    const int opix = ry * C3_OW + rx;
    const bool in_box = rx < C3_OW;
    uint32_t it = 0, slab_ctr = 0;
    for (long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const C3Tile t = c3_tile(p, item);
      const uint32_t acc = it & 1, aph = (it >> 1) & 1;
      const int y = t.y0 + ry, x = t.x0 + rx;
      const bool valid = in_box && y < p.H && x < p.W;
      const long prow = ((long)t.img * p.H + y) * p.W + x;
      const uint32_t t_row = tmem_base + acc * C3_NC_MAX + ((uint32_t)(quarter * 32) << 16);
      const int nbase = t.nchunk * p.nc;
      mbar_wait_relaxed(tfull_bar(acc), aph);
      tc_fence_after();
      if (FAST) {
        const float* __restrict__ bias_p = p.epi.col_bias;
        for (int j = 0; j < nslabs; ++j, ++slab_ctr) {
          const int b = slab_ctr % p.nslab;
          const uint32_t sph = (slab_ctr / p.nslab) & 1;
          if (p.has_res) mbar_wait_relaxed(sfull_bar(b), sph);
          else mbar_wait_relaxed(sempty_bar(b), sph ^ 1);
          const int col0 = j * 64 + cq * 16;
          if (col0 < p.nc) {
            uint32_t v[16];
            tmem_ld16_issue(t_row + col0, v);
            tmem_ld16_wait(v);
            if (in_box) {
              uint8_t* srow = slab_gen + b * C3_SLAB + opix * 128;
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                const int chunk = cq * 2 + q2;
                const int n = nbase + j * 64 + chunk * 8;
                uint4* sp = reinterpret_cast<uint4*>(srow + ((chunk ^ (opix & 7)) << 4));
                if (n < p.epi.N) {
                  float f[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q2 * 8 + e]);
                  if (bias_p) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias_p + n));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias_p + n + 4));
                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                  }
                  if (p.has_res) {
                    const uint4 rr4 = *sp;
                    const uint32_t w4[4] = {rr4.x, rr4.y, rr4.z, rr4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      f[2 * e] += __uint_as_float(w4[e] << 16);
                      f[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
                    }
                  }
                  if (p.relu) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                  }
                  uint4 o;
                  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                  *sp = o;
                }
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (j == nslabs - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          named_bar_arrive(1 + b, C3_STORE_BAR_THREADS);
        }
      } else {
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
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q4 * 8 + e]);
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
  } else if (warp == 2 + C3_EPI_WARPS) {
    // ===================== residual-slab producer =====================
    if (FAST && p.has_res && lane == 0) {
      uint32_t slab_ctr = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const C3Tile t = c3_tile(p, item);
        for (int j = 0; j < nslabs; ++j, ++slab_ctr) {
          const int b = slab_ctr % p.nslab;
          mbar_wait_relaxed(sempty_bar(b), ((slab_ctr / p.nslab) & 1) ^ 1);
          mbar_expect_tx(sfull_bar(b), C3_OW * C3_OH * 128);
          tma_load_4d(slab_base + b * C3_SLAB, &map_res, sfull_bar(b), t.nchunk * p.nc + j * 64, t.x0, t.y0, t.img);
        }
      }
    }
  } else {
    // ===================== slab store issuer =====================
    if (FAST) {
      uint32_t slab_ctr = 0;
      for (long item = blockIdx.x; item < p.items; item += gridDim.x) {
        const C3Tile t = c3_tile(p, item);
        for (int j = 0; j < nslabs; ++j, ++slab_ctr) {
          const int b = slab_ctr % p.nslab;
          named_bar_sync(1 + b, C3_STORE_BAR_THREADS);
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(&map_out), "r"(slab_base + b * C3_SLAB), "r"(t.nchunk * p.nc + j * 64), "r"(t.x0), "r"(t.y0), "r"(t.img)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // a slab is handed back once its store has been read out of shared memory; the number of stores left in flight is
            // the ring depth minus two (the slab being written and the one just committed)
            if (p.nslab >= 4) {
              asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
              if (slab_ctr >= 2) mbar_arrive(sempty_bar((slab_ctr - 2) % p.nslab));
            } else {
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              mbar_arrive(sempty_bar(slab_ctr % p.nslab));
            }
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

}  // namespace

// dense 3x3 (kd = 1) or 3x3x3 (kd = 3) conv, dilation 1; called by conv_gemm_tc for every spatial op
int conv3x3_tc(const ConvOp& op, int nc, int n_chunks, cudaStream_t s) {
  static DeviceOnce once;
  bool first; int dev, g_c3_sms;
  KD_TRY(device_first_use(once, &first, &dev));
  if (first) {
    KD_CUDA(cudaFuncSetAttribute(k_conv3_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM_MAX));
    KD_CUDA(cudaFuncSetAttribute(k_conv3_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SMEM_MAX));
    device_mark(once, dev);
  }
  KD_TRY(device_sms(&g_c3_sms));
  const Epilogue& e = op.epi;
  C3Params p;
  memset(&p, 0, sizeof(p));
  p.kd = op.kd; p.D = op.D;
  p.kc0 = (op.c0 + 63) / 64; p.kc1 = (op.c1 + 63) / 64; p.c0 = op.c0; p.c1 = op.c1;
  p.w_tap_ld = op.w_tap_ld;
  p.nc = nc; p.n_chunks = n_chunks;
  p.H = op.H; p.W = op.W;
  p.tiles_x = cdiv(op.W, C3_OW); p.tiles_y = cdiv(op.H, C3_OH);
  p.items = (long)op.nimg * p.tiles_x * p.tiles_y * n_chunks;
  KD_CHECK(p.items < (1L << 24), "conv3x3_tc: too many tiles (%ld)", p.items);
  p.has_res = e.res != nullptr; p.relu = e.relu;
  // weight-stationary when every tap tile of a single-K-chunk conv fits next to two activation slots: all 2-D layers with
  // N <= 64 and the 3x3x3 layers of KDLAE-S up to 32 output channels (27 * 32 * 128 B = 108 KB).  Streaming the 27 tap tiles
  // per output tile (9 dependent TMA round trips of ~1 us) made the 16 -> 16 layer at 512^2 take 8500 clk per 120-pixel tile.
  p.xpack_cin = op.xpack_cin;
  p.w_tiles = op.xpack_cin ? xconv_tiles(op.kd, op.xpack_cin) : 9 * op.kd * (p.kc0 + p.kc1);
  p.nslab = C3_NSLAB;
  const uint32_t w_res_bytes = ((uint32_t)p.w_tiles * nc * 128 + 1023u) & ~1023u;
  p.w_resident = (p.kc0 + p.kc1 == 1 && n_chunks == 1 &&
                  w_res_bytes + C3_ASLOTS * C3_A_SLOT + C3_NSLAB * C3_SLAB + 1024 + 512 <= C3_SMEM_MAX &&
                  (op.kd == 3 || op.xpack_cin || 9u * nc * 128 <= C3_BSLOTS * C3_B_SLOT)) ? 1 : 0;
  // 2-D convs whose 18 tap tiles (two K chunks x N <= 64) or 9 tap tiles (one K chunk x N <= 128) take 144 KB: resident next to
  // two activation slots and a two-deep slab ring (the 128 -> 64 and 64 -> 128 layers of ASDQE at 512^2 / 256^2)
  const bool res2 = KDLAE_C3_RES2 && !p.w_resident && op.kd == 1 && !op.xpack_cin && n_chunks == 1 && p.kc0 + p.kc1 <= 2 &&
                    w_res_bytes + C3_ASLOTS * C3_A_SLOT + 2 * C3_SLAB + 1024 + 512 <= C3_SMEM_MAX;
  if (res2) { p.w_resident = 1; p.nslab = 2; }
  if (op.xpack_cin) {
    KD_CHECK((op.xpack_cin == 16 || op.xpack_cin == 32) && op.c0 == 64 && op.c1 == 0 && n_chunks == 1 && nc % 16 == 0 && op.w_tap_ld == 64,
             "conv3x3_tc: bad x-packed conv (cin %d, c0 %d, N %d)", op.xpack_cin, op.c0, e.N);
    // a two-deep slab ring leaves the room to an extra activation slot (3x3x3 tiles of a 64-wide output take 112-144 KB)
    p.nslab = 2;
    p.w_resident = (w_res_bytes + C3_ASLOTS * C3_A_SLOT + 2 * C3_SLAB + 1024 + 512 <= C3_SMEM_MAX) ? 1 : 0;
    KD_CHECK(p.w_resident, "conv3x3_tc: x-packed weights (%u bytes) do not fit shared memory", w_res_bytes);
  }
  p.a_slots = C3_ASLOTS; p.b_bytes = C3_BSLOTS * C3_B_SLOT;
  p.b_group = (2u * 3u * nc * 128 <= p.b_bytes && !op.xpack_cin) ? 3 : 1;      // a kernel row of taps per weight load when two such slots fit
  p.b_slot = (uint32_t)p.b_group * nc * 128;
  p.b_slots = (int)std::min<uint32_t>(C3_BSLOTS, p.b_bytes / p.b_slot);     // deeper weight rings measured slower (they delay the A loads)
  uint32_t smem = C3_SMEM;
  if (p.w_resident) {   // resident weights take 9 * kd * nc * 128 bytes; the rest of the per-tap ring deepens the A ring
    p.b_bytes = w_res_bytes;
    const uint32_t fixed = p.b_bytes + p.nslab * C3_SLAB + 1024 + 512;
    p.a_slots = (int)std::min<uint32_t>(C3_ASLOTS_MAX, (C3_SMEM_MAX - fixed) / C3_A_SLOT);
    smem = p.a_slots * C3_A_SLOT + fixed;
  }
  p.inv_n_chunks = 1.0f / (float)n_chunks; p.inv_tiles_x = 1.0f / (float)p.tiles_x; p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.inv_D = 1.0f / (float)(op.D > 0 ? op.D : 1);
  p.epi = e;
  const bool fast = e.mode == OUT_IDENTITY && e.N % 8 == 0 && e.row_mu == nullptr && e.row_scale == nullptr && e.stat_rstd == nullptr &&
                    e.out_ld % 8 == 0 && e.out_coff % 8 == 0 && (reinterpret_cast<uintptr_t>(e.out) & 15) == 0 &&
                    (e.res == nullptr || (e.res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(e.res) & 15) == 0)) &&
                    (e.col_bias == nullptr || (reinterpret_cast<uintptr_t>(e.col_bias) & 15) == 0);
  CUtensorMap ma0, ma1, mw, mout, mres;
  auto a_map = [&](CUtensorMap* m, const void* base, int ch, long ld) -> int {
    if (op.kd == 3) {
      const cuuint64_t dims[5] = {(cuuint64_t)ch, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.D, (cuuint64_t)(op.nimg / op.D)};
      const cuuint64_t f = (cuuint64_t)ld * 2 * op.W * op.H;
      const cuuint64_t str[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * op.W, f, f * op.D};
      const cuuint32_t box[5] = {64, C3_TW, C3_IH, 1, 1};
      return make_map(m, base, 5, dims, str, box);
    }
    const cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * op.W, (cuuint64_t)ld * 2 * op.W * op.H};
    const cuuint32_t box[4] = {64, C3_TW, C3_IH, 1};
    return make_map(m, base, 4, dims, str, box);
  };
  auto o_map = [&](CUtensorMap* m, const void* base, long ld) -> int {
    const cuuint64_t dims[4] = {(cuuint64_t)e.N, (cuuint64_t)op.W, (cuuint64_t)op.H, (cuuint64_t)op.nimg};
    const cuuint64_t str[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * op.W, (cuuint64_t)ld * 2 * op.W * op.H};
    const cuuint32_t box[4] = {64, C3_OW, C3_OH, 1};
    return make_map(m, base, 4, dims, str, box);
  };
  KD_TRY(a_map(&ma0, op.a0, op.c0, op.ld0));
  if (op.c1 > 0) KD_TRY(a_map(&ma1, op.a1, op.c1, op.ld1));
  else ma1 = ma0;
  if (fast) {
    KD_TRY(o_map(&mout, reinterpret_cast<const bf16*>(e.out) + e.out_coff, e.out_ld));
    if (e.res) KD_TRY(o_map(&mres, e.res, e.res_ld));
    else mres = mout;
  } else {
    mout = ma0; mres = ma0;
  }
  {
    // weights [n][tap][c] seen as {c within a tap, n, tap}: one box = b_group consecutive taps of nc rows, each a [nc][64] K-major tile
    const int taps = op.xpack_cin ? p.w_tiles : 9 * op.kd;
    const cuuint64_t dims[3] = {(cuuint64_t)op.w_tap_ld, (cuuint64_t)e.N, (cuuint64_t)taps};
    const cuuint64_t str[2] = {(cuuint64_t)op.w_ld * 2, (cuuint64_t)op.w_tap_ld * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)nc, (cuuint32_t)p.b_group};
    KD_TRY(make_map(&mw, op.w, 3, dims, str, box));
  }
  const int grid = (int)(p.items < (long)g_c3_sms ? p.items : (long)g_c3_sms);
  // algorithmic work: an x-packed conv does the FLOPs of the cin -> N/P conv it stands for (rows are super-pixels of P pixels)
  const double rows = (double)op.nimg * op.H * op.W;
  const double ktot = op.xpack_cin ? 9.0 * op.kd * op.xpack_cin : 9.0 * op.kd * (op.c0 + op.c1);
  const double nalg = op.xpack_cin ? (double)e.N : (double)e.N;
  ProfScope prof(PC_GEMM_TC, s, 2.0 * rows * nalg * ktot, 2.0 * (rows * (op.c0 + op.c1 + e.N * (e.res ? 2 : 1)) + e.N * ktot));
  if (fast) k_conv3_tc<1><<<grid, C3_THREADS, smem, s>>>(ma0, ma1, mw, mout, mres, p);
  else k_conv3_tc<0><<<grid, C3_THREADS, smem, s>>>(ma0, ma1, mw, mout, mres, p);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
