// uint8 pre / post-processing passes around the KDLAE-T forward (reference: KDLAE/KDLAE_T.ipynb cell 5).
//   pre : uint8 HWC image -> fp32 NCHW in [0,1] (/255), reflect-padded bottom/right to the model's size multiple, plus the
//         constant denoise-rate map the module's forward takes (one float per image broadcast to H x W)
//   post: fp32 NCHW prediction (hq, or sr at scale 2) -> clamp(0,1) -> crop -> rint(x*255) (skimage.img_as_ubyte) ->
//         zero wherever every channel of the source pixel is 0 (the sonar blind zone; the sr mask is the 2x2-repeated one)
//         -> uint8 HWC.
// Both are single HBM passes: 1 byte per element across PCIe instead of 4 in each direction.
#include "ops.cuh"

namespace kd {

namespace {

// thread = one padded pixel (all channels): coalesced fp32 plane writes, byte gather from the HWC source
__global__ void __launch_bounds__(256) k_pre_u8(const uint8_t* __restrict__ src, int h, int w, int c, const float* __restrict__ rates,
                                                float* __restrict__ img, float* __restrict__ rate_map, int H, int W, long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long HW = (long)H * W;
  const int b = (int)(idx / HW);
  const long sp = idx - (long)b * HW;
  const int y = (int)(sp / W), x = (int)(sp - (long)y * W);
  const int ys = y < h ? y : 2 * (h - 1) - y, xs = x < w ? x : 2 * (w - 1) - x;   // F.pad(..., 'reflect') on the far edges
  const uint8_t* p = src + (((long)b * h + ys) * w + xs) * c;
  for (int k = 0; k < c; ++k) img[((long)b * c + k) * HW + sp] = (float)p[k] / 255.0f;
  if (rate_map) rate_map[idx] = rates[b];
}

// thread = one output pixel of the cropped (h*scale) x (w*scale) image
__global__ void __launch_bounds__(256) k_post_u8(const float* __restrict__ pred, const uint8_t* __restrict__ src, int h, int w, int c,
                                                 int Hp, int Wp, int scale, uint8_t* __restrict__ out, long total) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ho = h * scale, wo = w * scale;
  const long hw = (long)ho * wo;
  const int b = (int)(idx / hw);
  const long sp = idx - (long)b * hw;
  const int y = (int)(sp / wo), x = (int)(sp - (long)y * wo);
  const uint8_t* ps = src + (((long)b * h + y / scale) * w + x / scale) * c;
  bool blind = true;
  for (int k = 0; k < c; ++k) blind = blind && ps[k] == 0;
  const long plane = (long)Hp * Wp;
  uint8_t* po = out + idx * c;
  for (int k = 0; k < c; ++k) {
    float v = pred[((long)b * c + k) * plane + (long)y * Wp + x];
    v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;
    po[k] = blind ? (uint8_t)0 : (uint8_t)__float2int_rn(v);    // rint: round half to even, as numpy.rint
  }
}

}  // namespace

int preprocess_u8(const uint8_t* src, int B, int h, int w, int c, const float* rates, float* img, float* rate_map, int H, int W,
                  cudaStream_t s) {
  KD_CHECK(B >= 1 && h >= 1 && w >= 1 && c >= 1 && c <= 4 && H >= h && W >= w, "preprocess_u8: bad shape %dx%dx%d -> %dx%d", h, w, c, H, W);
  KD_CHECK(H - h < h && W - w < w, "preprocess_u8: reflect padding needs pad < size (%d,%d of %d,%d)", H - h, W - w, h, w);
  KD_CHECK(rate_map == nullptr || rates != nullptr, "preprocess_u8: rate map requested without rates");
  const long total = (long)B * H * W;
  ProfScope prof(PC_POOL_RESAMPLE, s, 0.0, (double)B * h * w * c + (double)total * (c + (rate_map ? 1 : 0)) * 4.0);
  k_pre_u8<<<(unsigned)cdiv(total, 256), 256, 0, s>>>(src, h, w, c, rates, img, rate_map, H, W, total);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

int postprocess_u8(const float* pred, const uint8_t* src, int B, int h, int w, int c, int Hp, int Wp, int scale, uint8_t* out,
                   cudaStream_t s) {
  KD_CHECK(B >= 1 && c >= 1 && c <= 4 && (scale == 1 || scale == 2) && Hp >= h * scale && Wp >= w * scale,
           "postprocess_u8: bad shape %dx%dx%d scale %d in %dx%d", h, w, c, scale, Hp, Wp);
  const long total = (long)B * h * scale * w * scale;
  ProfScope prof(PC_POOL_RESAMPLE, s, 0.0, (double)total * c * 5.0 + (double)B * h * w * c);
  k_post_u8<<<(unsigned)cdiv(total, 256), 256, 0, s>>>(pred, src, h, w, c, Hp, Wp, scale, out, total);
  count_launch();
  KD_LAUNCH_CHECK();
  return 0;
}

}  // namespace kd
