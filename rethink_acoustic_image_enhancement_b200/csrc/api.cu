// extern "C" boundary (include/kdlae_b200.h): argument checking, precision dispatch, error text.
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>
#include "models.cuh"

namespace kd {

static thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launch_count{0};

int device_first_use(DeviceOnce& once, bool* first, int* dev) {
  KD_CUDA(cudaGetDevice(dev));
  KD_CHECK(*dev >= 0 && *dev < KD_MAX_DEVICES, "device index %d out of range", *dev);
  *first = !(once.mask.load(std::memory_order_acquire) & (1ull << *dev));
  return 0;
}
int device_sms(int* sms) {
  static std::atomic<int> cache[KD_MAX_DEVICES];
  int dev = 0;
  KD_CUDA(cudaGetDevice(&dev));
  KD_CHECK(dev >= 0 && dev < KD_MAX_DEVICES, "device index %d out of range", dev);
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    KD_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    v = sm_limit(v);
    cache[dev].store(v, std::memory_order_relaxed);
  }
  *sms = v;
  return 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ---- profiler ----------------------------------------------------------------------------
struct ProfRec { cudaEvent_t a, b; int cls; double flops, bytes; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;   // launches may come from several host threads (nn.DataParallel replicas)
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_event_pool;
static size_t g_pool_used = 0;
static const char* kProfNames[PC_COUNT] = {"conv_gemm_tcgen05", "conv_gemm_simt", "dwconv3x3", "ln_stats", "mdta_gram",
                                           "mdta_softmax_fold", "small_channel_conv", "pool_resample", "gap_mlp_head",
                                           "fused_conv1x1_dwconv3x3_tcgen05"};

static cudaEvent_t prof_event() {
  if (g_pool_used == g_event_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_event_pool.push_back(e);
  }
  return g_event_pool[g_pool_used++];
}
ProfScope::ProfScope(int cls, cudaStream_t s, double flops, double bytes) : stream(s) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  active = true;
  ProfRec r;
  r.a = prof_event(); r.b = prof_event(); r.cls = cls; r.flops = flops; r.bytes = bytes;
  cudaEventRecord(r.a, s);
  slot = (int)g_recs.size();
  g_recs.push_back(r);
}
ProfScope::~ProfScope() {
  if (!active) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_recs.size()) cudaEventRecord(g_recs[slot].b, stream);
}

// ---- debug trace -------------------------------------------------------------------------
static bool g_trace_on = false;
static unsigned long long* g_trace_dev = nullptr;
static std::vector<std::string> g_trace_tags;
constexpr int TRACE_MAX = 8192;

__global__ void k_trace_checksum(const uint8_t* __restrict__ p, long rows, long row_words, long ld_bytes, unsigned long long* slot) {
  unsigned long long a = 0, b = 0;
  const long total = rows * row_words;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long r = i / row_words, c = i - r * row_words;
    const unsigned int w = *reinterpret_cast<const unsigned int*>(p + r * ld_bytes + c * 4);
    a += w;
    b += (unsigned long long)w * (unsigned long long)(i % 65521 + 1);
  }
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(slot, a); atomicAdd(slot + 1, b); }
}

void trace_point(const char* tag, const void* ptr, long rows, long row_bytes, long ld_bytes, cudaStream_t s) {
  if (!g_trace_on || (int)g_trace_tags.size() >= TRACE_MAX || ptr == nullptr || rows <= 0 || row_bytes < 4) return;
  const long row_words = row_bytes / 4;
  const long total = rows * row_words;
  const int blocks = (int)std::min<long>(2048, (total + 255) / 256);
  k_trace_checksum<<<blocks, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(ptr), rows, row_words, ld_bytes,
                                          g_trace_dev + 2 * g_trace_tags.size());
  g_trace_tags.emplace_back(tag);
}

}  // namespace kd

using namespace kd;

#define API_BEGIN() kd::set_error("")
#define CHECK_PREC(p) KD_CHECK((p) == KDLAE_PREC_FP32 || (p) == KDLAE_PREC_BF16, "precision must be KDLAE_PREC_FP32 or KDLAE_PREC_BF16")

extern "C" {

int kdlae_abi_version(void) { return KDLAE_ABI_VERSION; }
const char* kdlae_last_error(void) { return kd::get_error(); }
unsigned long long kdlae_launch_count(void) { return kd::g_launch_count.load(); }

int kdlae_device_check(int device) {
  API_BEGIN();
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    kd::set_error("no CUDA device available (%s)", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return 3;
  }
  KD_CHECK(device >= 0 && device < n, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  KD_CUDA(cudaGetDeviceProperties(&prop, device));
  KD_CHECK(prop.major == 10, "device %d (%s, sm_%d%d) is not an sm_100 (B200) GPU; this library only carries sm_100a code",
           device, prop.name, prop.major, prop.minor);
  return 0;
}

// ---------------- profiler ----------------
int kdlae_profile_num_classes(void) { return kd::PC_COUNT; }
const char* kdlae_profile_class_name(int cls) { return (cls >= 0 && cls < kd::PC_COUNT) ? kd::kProfNames[cls] : ""; }
int kdlae_profile_begin(void) {
  API_BEGIN();
  std::lock_guard<std::mutex> lk(kd::g_prof_mu);
  kd::g_recs.clear();
  kd::g_pool_used = 0;
  kd::g_prof_on = true;
  return 0;
}
int kdlae_profile_launches(int max_launches, int* cls, double* ms, double* flops, double* bytes) {
  API_BEGIN();
  KD_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(kd::g_prof_mu);
  const int n = (int)std::min<size_t>(kd::g_recs.size(), (size_t)std::max(0, max_launches));
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, kd::g_recs[i].a, kd::g_recs[i].b) != cudaSuccess) { (void)cudaGetLastError(); t = -1.f; }
    if (cls) cls[i] = kd::g_recs[i].cls;
    if (ms) ms[i] = t;
    if (flops) flops[i] = kd::g_recs[i].flops;
    if (bytes) bytes[i] = kd::g_recs[i].bytes;
  }
  return n;
}
int kdlae_profile_end(int n_classes, double* ms, double* flops, double* bytes, long long* launches) {
  API_BEGIN();
  kd::g_prof_on = false;
  KD_CHECK(n_classes == kd::PC_COUNT && ms && flops && bytes && launches, "kdlae_profile_end: bad arguments");
  for (int i = 0; i < kd::PC_COUNT; ++i) { ms[i] = 0; flops[i] = 0; bytes[i] = 0; launches[i] = 0; }
  KD_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(kd::g_prof_mu);
  for (const auto& r : kd::g_recs) {
    float t = 0.f;
    KD_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cls] += t; flops[r.cls] += r.flops; bytes[r.cls] += r.bytes; launches[r.cls] += 1;
  }
  kd::g_recs.clear();
  kd::g_pool_used = 0;
  return 0;
}

// ---------------- teacher ----------------
int kdlae_teacher_num_tensors(const kdlae_teacher_cfg* cfg) { return cfg ? kd::teacher_num_tensors(*cfg) : -1; }

size_t kdlae_teacher_packed_bytes(const kdlae_teacher_cfg* cfg, int precision) {
  if (!cfg) return 0;
  return precision == KDLAE_PREC_BF16 ? kd::teacher_packed_bytes<bf16>(*cfg) : kd::teacher_packed_bytes<float>(*cfg);
}
int kdlae_teacher_pack(const kdlae_teacher_cfg* cfg, const float* const* tensors, int n_tensors, void* packed, size_t packed_bytes,
                       int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(cfg && tensors && packed, "kdlae_teacher_pack: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16 ? kd::teacher_pack<bf16>(*cfg, tensors, n_tensors, packed, packed_bytes, s)
                                      : kd::teacher_pack<float>(*cfg, tensors, n_tensors, packed, packed_bytes, s);
}
size_t kdlae_teacher_workspace_bytes(const kdlae_teacher_cfg* cfg, int micro_batch, int H, int W, int precision) {
  if (!cfg || micro_batch < 1 || H < 8 || W < 8) return 0;
  return precision == KDLAE_PREC_BF16 ? kd::teacher_workspace_bytes<bf16>(*cfg, micro_batch, H, W)
                                      : kd::teacher_workspace_bytes<float>(*cfg, micro_batch, H, W);
}
int kdlae_teacher_forward(const kdlae_teacher_cfg* cfg, const void* packed, const float* img, const float* rate, int rate_per_image,
                          float* hq, float* sr, int B, int H, int W, int micro_batch, void* workspace, size_t workspace_bytes,
                          int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(cfg && packed && img && hq && workspace, "kdlae_teacher_forward: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16
             ? kd::teacher_forward<bf16>(*cfg, packed, img, rate, rate_per_image, hq, sr, B, H, W, micro_batch, workspace,
                                         workspace_bytes, s)
             : kd::teacher_forward<float>(*cfg, packed, img, rate, rate_per_image, hq, sr, B, H, W, micro_batch, workspace,
                                          workspace_bytes, s);
}

// ---------------- student ----------------
size_t kdlae_student_packed_bytes(const kdlae_student_cfg* cfg, int precision) {
  if (!cfg) return 0;
  return precision == KDLAE_PREC_BF16 ? kd::student_packed_bytes<bf16>(*cfg) : kd::student_packed_bytes<float>(*cfg);
}
int kdlae_student_pack(const kdlae_student_cfg* cfg, const float* const* tensors, int n_tensors, void* packed, size_t packed_bytes,
                       int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(cfg && tensors && packed, "kdlae_student_pack: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16 ? kd::student_pack<bf16>(*cfg, tensors, n_tensors, packed, packed_bytes, s)
                                      : kd::student_pack<float>(*cfg, tensors, n_tensors, packed, packed_bytes, s);
}
size_t kdlae_student_workspace_bytes(const kdlae_student_cfg* cfg, int micro_batch, int F, int H, int W, int precision) {
  if (!cfg || micro_batch < 1 || F < 1 || H < 4 || W < 4) return 0;
  return precision == KDLAE_PREC_BF16 ? kd::student_workspace_bytes<bf16>(*cfg, micro_batch, F, H, W)
                                      : kd::student_workspace_bytes<float>(*cfg, micro_batch, F, H, W);
}
int kdlae_student_forward(const kdlae_student_cfg* cfg, const void* packed, const float* x, float* y, int B, int F, int H, int W,
                          int micro_batch, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(cfg && packed && x && y && workspace, "kdlae_student_forward: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16
             ? kd::student_forward<bf16>(*cfg, packed, x, y, B, F, H, W, micro_batch, workspace, workspace_bytes, s)
             : kd::student_forward<float>(*cfg, packed, x, y, B, F, H, W, micro_batch, workspace, workspace_bytes, s);
}

// ---------------- ASDQE ----------------
size_t kdlae_asdqe_packed_bytes(const kdlae_asdqe_cfg* cfg, int precision) {
  if (!cfg) return 0;
  return precision == KDLAE_PREC_BF16 ? kd::asdqe_packed_bytes<bf16>(*cfg) : kd::asdqe_packed_bytes<float>(*cfg);
}
int kdlae_asdqe_pack(const kdlae_asdqe_cfg* cfg, const float* const* tensors, int n_tensors, void* packed, size_t packed_bytes,
                     int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(cfg && tensors && packed, "kdlae_asdqe_pack: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16 ? kd::asdqe_pack<bf16>(*cfg, tensors, n_tensors, packed, packed_bytes, s)
                                      : kd::asdqe_pack<float>(*cfg, tensors, n_tensors, packed, packed_bytes, s);
}
size_t kdlae_asdqe_workspace_bytes(const kdlae_asdqe_cfg* cfg, int micro_batch, int H, int W, int precision) {
  if (!cfg || micro_batch < 1 || H < 1 || W < 1) return 0;
  return precision == KDLAE_PREC_BF16 ? kd::asdqe_workspace_bytes<bf16>(*cfg, micro_batch, H, W)
                                      : kd::asdqe_workspace_bytes<float>(*cfg, micro_batch, H, W);
}
int kdlae_asdqe_forward(const kdlae_asdqe_cfg* cfg, const void* packed, const float* lq, const float* gt, float* score, float* feat,
                        int B, int H, int W, int micro_batch, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(cfg && packed && lq && gt && score && workspace, "kdlae_asdqe_forward: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16
             ? kd::asdqe_forward<bf16>(*cfg, packed, lq, gt, score, feat, B, H, W, micro_batch, workspace, workspace_bytes, s)
             : kd::asdqe_forward<float>(*cfg, packed, lq, gt, score, feat, B, H, W, micro_batch, workspace, workspace_bytes, s);
}

// ---------------- single stages ----------------
int kdlae_conv_gemm(const void* a, int C, const void* w, int N, int nimg, int H, int W, int ksize, const float* row_scale,
                    const float* col_bias, int relu, const void* res, void* out, int precision, int force_simt, void* stream) {
  API_BEGIN();
  KD_CHECK(a && w && out, "kdlae_conv_gemm: NULL argument");
  KD_CHECK(ksize == 1 || ksize == 3, "kdlae_conv_gemm: ksize must be 1 or 3");
  CHECK_PREC(precision);
  ConvOp g;
  g.a0 = a; g.c0 = C; g.ld0 = C; g.nimg = nimg; g.H = H; g.W = W; g.kh = g.kw = ksize;
  g.w = w; g.w_ld = (long)ksize * ksize * C; g.w_tap_ld = C;
  g.epi.row_scale = row_scale; g.epi.col_bias = col_bias; g.epi.relu = relu; g.epi.res = res; g.epi.res_ld = N;
  g.epi.out = out; g.epi.out_ld = N; g.epi.N = N; g.epi.H = H; g.epi.W = W;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (precision == KDLAE_PREC_FP32) return kd::conv_gemm_simt<float>(g, s);
  if (force_simt) return kd::conv_gemm_simt<bf16>(g, s);
  KD_CHECK(kd::conv_gemm_tc_eligible(g), "kdlae_conv_gemm: shape not eligible for the tcgen05 kernel (C=%d N=%d)", C, N);
  return kd::conv_gemm_tc(g, s);
}

namespace kd {
__global__ void k_sum_splits(const float* __restrict__ part, int splits, long psz, long total, float* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long g = i / psz, e = i % psz;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += part[(g * splits + sp) * psz + e];
  out[i] = s;
}
}  // namespace kd

size_t kdlae_mdta_gram_scratch_floats(int nimg, int HW, int C, int heads) {
  const int ch = heads > 0 ? C / heads : 0;
  return (size_t)nimg * heads * kd::mdta_gram_splits(HW, nimg * heads) * ((size_t)ch * ch + 2 * ch);
}

int kdlae_mdta_gram(const void* qk, long ld, int nimg, int HW, int C, int heads, float* gram, float* scratch, int precision,
                    void* stream) {
  API_BEGIN();
  KD_CHECK(qk && gram && scratch && heads > 0 && C % heads == 0, "kdlae_mdta_gram: bad argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int splits = kd::mdta_gram_splits(HW, nimg * heads);
  if (precision == KDLAE_PREC_BF16) KD_TRY(kd::mdta_gram<bf16>(reinterpret_cast<const bf16*>(qk), ld, nimg, HW, C, heads, splits, scratch, s));
  else KD_TRY(kd::mdta_gram<float>(reinterpret_cast<const float*>(qk), ld, nimg, HW, C, heads, splits, scratch, s));
  const int ch = C / heads;
  const long psz = (long)ch * ch + 2 * ch, total = (long)nimg * heads * psz;
  kd::k_sum_splits<<<kd::cdiv(total, 256), 256, 0, s>>>(scratch, splits, psz, total, gram);
  KD_LAUNCH_CHECK();
  return 0;
}

int kdlae_ln_stats(const void* x, int C, long rows, float* rstd, float* mu, int precision, void* stream) {
  API_BEGIN();
  KD_CHECK(x && rstd, "kdlae_ln_stats: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  return precision == KDLAE_PREC_BF16 ? kd::ln_stats<bf16>(reinterpret_cast<const bf16*>(x), C, C, rows, rstd, mu, s)
                                      : kd::ln_stats<float>(reinterpret_cast<const float*>(x), C, C, rows, rstd, mu, s);
}

int kdlae_dwconv3x3(const void* x, void* out, const float* w9c, int nimg, int H, int W, int C, int gate, int precision,
                    void* stream) {
  API_BEGIN();
  KD_CHECK(x && out && w9c, "kdlae_dwconv3x3: NULL argument");
  CHECK_PREC(precision);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long ldo = gate ? C / 2 : C;
  return precision == KDLAE_PREC_BF16
             ? kd::dwconv3x3<bf16>(reinterpret_cast<const bf16*>(x), C, reinterpret_cast<bf16*>(out), ldo, w9c, nullptr, nimg, H, W,
                                   C, gate, s)
             : kd::dwconv3x3<float>(reinterpret_cast<const float*>(x), C, reinterpret_cast<float*>(out), ldo, w9c, nullptr, nimg, H,
                                    W, C, gate, s);
}

int kdlae_pwdw_f2(const void* x, const float* rstd, const void* w1, int Nt, const float* w9c, void* out, int nimg, int H, int W, int C,
                  int gate, void* stream) {
  API_BEGIN();
  KD_CHECK(x && rstd && w1 && w9c && out, "kdlae_pwdw_f2: NULL argument");
  return kd::pwdw_f2(reinterpret_cast<const bf16*>(x), C, rstd, reinterpret_cast<const bf16*>(w1), Nt, w9c,
                     reinterpret_cast<bf16*>(out), gate ? Nt / 2 : Nt, nimg, H, W, C, gate, reinterpret_cast<cudaStream_t>(stream));
}

int kdlae_pwdw_t(const void* x, const float* rstd, const void* w1, int Nt, const float* w9c, void* out, int nimg, int H, int W, int C,
                 int gate, void* stream) {
  API_BEGIN();
  KD_CHECK(x && rstd && w1 && w9c && out, "kdlae_pwdw_t: NULL argument");
  return kd::pwdw_t(reinterpret_cast<const bf16*>(x), C, rstd, reinterpret_cast<const bf16*>(w1), Nt, w9c,
                    reinterpret_cast<bf16*>(out), gate ? Nt / 2 : Nt, nimg, H, W, C, gate, reinterpret_cast<cudaStream_t>(stream));
}

int kdlae_preprocess_u8(const unsigned char* src_hwc, int B, int h, int w, int c, const float* rates, float* img_nchw, float* rate_map,
                        int H, int W, void* stream) {
  API_BEGIN();
  KD_CHECK(src_hwc && img_nchw, "kdlae_preprocess_u8: NULL argument");
  return kd::preprocess_u8(src_hwc, B, h, w, c, rates, img_nchw, rate_map, H, W, reinterpret_cast<cudaStream_t>(stream));
}

int kdlae_postprocess_u8(const float* pred_nchw, const unsigned char* src_hwc, int B, int h, int w, int c, int Hp, int Wp, int scale,
                         unsigned char* out_hwc, void* stream) {
  API_BEGIN();
  KD_CHECK(pred_nchw && src_hwc && out_hwc, "kdlae_postprocess_u8: NULL argument");
  return kd::postprocess_u8(pred_nchw, src_hwc, B, h, w, c, Hp, Wp, scale, out_hwc, reinterpret_cast<cudaStream_t>(stream));
}

size_t kdlae_psnr_scratch_bytes(int B) { return B > 0 ? kd::psnr_scratch_bytes(B) : 0; }

int kdlae_psnr(const float* img1, const float* img2, int B, int C, int H, int W, int crop_border, int as_uint8, double* mse_max,
               void* scratch, void* stream) {
  API_BEGIN();
  KD_CHECK(img1 && img2 && mse_max && scratch, "kdlae_psnr: NULL argument");
  return kd::psnr_mse(img1, img2, B, C, H, W, crop_border, as_uint8, mse_max, scratch, reinterpret_cast<cudaStream_t>(stream));
}

size_t kdlae_l1_sr_scratch_bytes(void) { return kd::l1_sr_scratch_bytes(); }

int kdlae_l1_sr_loss(const float* hq, const float* hq_gt, long n_hq, const float* sr, const float* sr_gt, long n_sr,
                     float loss_weight, float* loss, float* grad_hq, float* grad_sr, double* terms, void* scratch, void* stream) {
  API_BEGIN();
  KD_CHECK(hq && hq_gt && loss && scratch && n_hq > 0, "kdlae_l1_sr_loss: NULL argument");
  KD_CHECK((sr == nullptr) == (sr_gt == nullptr) && (sr == nullptr || n_sr > 0), "kdlae_l1_sr_loss: sr and sr_gt go together");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // 0.5 * lw * L1(hq) + 0.25 * lw * shadow(hq)   [+ 0.25 * lw * L1(sr) + 0.25 * lw * shadow(sr)]
  KD_TRY(kd::l1_shadow_term(hq, hq_gt, n_hq, 0.5f * loss_weight, 0.25f * loss_weight, 0, loss, grad_hq, terms, scratch, s));
  if (sr) {
    KD_TRY(kd::l1_shadow_term(sr, sr_gt, n_sr, 0.25f * loss_weight, 0.25f * loss_weight, 1, loss, grad_sr, terms ? terms + 2 : nullptr,
                              scratch, s));
  }
  return 0;
}

size_t kdlae_gdfn_train_ws_floats(int nimg, int H, int W, int C, int hp) {
  return (nimg > 0 && H > 0 && W > 0 && C > 0 && hp > 0) ? kd::gdfn_train_ws_floats(nimg, H, W, C, hp) : 0;
}
int kdlae_gdfn_forward_train(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, float* out,
                             int nimg, int H, int W, int C, int hp, float* ws, void* stream) {
  API_BEGIN();
  KD_CHECK(x && gamma && w_in && w_dw && w_out && out && ws, "kdlae_gdfn_forward_train: NULL argument");
  return kd::gdfn_forward_train(x, gamma, w_in, w_dw, w_out, out, nimg, H, W, C, hp, ws, reinterpret_cast<cudaStream_t>(stream));
}
int kdlae_gdfn_backward(const float* x, const float* gamma, const float* w_in, const float* w_dw, const float* w_out, const float* dout,
                        float* dx, float* dgamma, float* dw_in, float* dw_dw, float* dw_out, int nimg, int H, int W, int C, int hp,
                        float* ws, void* stream) {
  API_BEGIN();
  KD_CHECK(x && gamma && w_in && w_dw && w_out && dout && dx && dgamma && dw_in && dw_dw && dw_out && ws, "kdlae_gdfn_backward: NULL argument");
  return kd::gdfn_backward(x, gamma, w_in, w_dw, w_out, dout, dx, dgamma, dw_in, dw_dw, dw_out, nimg, H, W, C, hp, ws,
                           reinterpret_cast<cudaStream_t>(stream));
}
size_t kdlae_mdta_train_ws_floats(int nimg, int H, int W, int C, int heads) {
  return (nimg > 0 && H > 0 && W > 0 && C > 0 && heads > 0 && C % heads == 0) ? kd::mdta_train_ws_floats(nimg, H, W, C, heads) : 0;
}
int kdlae_mdta_forward_train(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj,
                             const float* temp, float* out, int nimg, int H, int W, int C, int heads, float* ws, void* stream) {
  API_BEGIN();
  KD_CHECK(x && gamma && w_qkv && w_dw && w_proj && temp && out && ws, "kdlae_mdta_forward_train: NULL argument");
  return kd::mdta_forward_train(x, gamma, w_qkv, w_dw, w_proj, temp, out, nimg, H, W, C, heads, ws, reinterpret_cast<cudaStream_t>(stream));
}
int kdlae_mdta_backward(const float* x, const float* gamma, const float* w_qkv, const float* w_dw, const float* w_proj, const float* temp,
                        const float* dout, float* dx, float* dgamma, float* dw_qkv, float* dw_dw, float* dw_proj, float* dtemp, int nimg,
                        int H, int W, int C, int heads, float* ws, void* stream) {
  API_BEGIN();
  KD_CHECK(x && gamma && w_qkv && w_dw && w_proj && temp && dout && dx && dgamma && dw_qkv && dw_dw && dw_proj && dtemp && ws,
           "kdlae_mdta_backward: NULL argument");
  return kd::mdta_backward(x, gamma, w_qkv, w_dw, w_proj, temp, dout, dx, dgamma, dw_qkv, dw_dw, dw_proj, dtemp, nimg, H, W, C, heads, ws,
                           reinterpret_cast<cudaStream_t>(stream));
}
int kdlae_set_train_matmul_tf32(int on) {
  kd::set_train_matmul_tf32(on);
  return 0;
}
int kdlae_train_matmul_tf32(void) { return kd::train_matmul_tf32(); }
size_t kdlae_conv_train_ws_floats(int nimg, int H, int W, int Cin, int Cout, int ksize) {
  return (nimg > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (ksize == 1 || ksize == 3)) ? kd::conv_train_ws_floats(nimg, H, W, Cin, Cout, ksize)
                                                                                         : 0;
}
int kdlae_conv_train_forward(const float* x, const float* w, float* out, int nimg, int H, int W, int Cin, int Cout, int ksize, int dilation,
                             void* stream) {
  API_BEGIN();
  KD_CHECK(x && w && out, "kdlae_conv_train_forward: NULL argument");
  return kd::conv_train_forward(x, w, out, nimg, H, W, Cin, Cout, ksize, dilation, reinterpret_cast<cudaStream_t>(stream));
}
int kdlae_conv_train_backward(const float* x, const float* w, const float* dout, float* dx, float* dw, int nimg, int H, int W, int Cin,
                              int Cout, int ksize, int dilation, float* ws, void* stream) {
  API_BEGIN();
  KD_CHECK(x && w && dout && dw && ws, "kdlae_conv_train_backward: NULL argument");
  return kd::conv_train_backward(x, w, dout, dx, dw, nimg, H, W, Cin, Cout, ksize, dilation, ws, reinterpret_cast<cudaStream_t>(stream));
}
int kdlae_grad_norm_sq(const float* grad, long n, double* norm_sq, double* scratch, void* stream) {
  API_BEGIN();
  return kd::grad_norm_sq(grad, n, norm_sq, scratch, reinterpret_cast<cudaStream_t>(stream));
}
int kdlae_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int step, float max_norm, const double* norm_sq, void* stream) {
  API_BEGIN();
  return kd::adamw_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, max_norm, norm_sq,
                        reinterpret_cast<cudaStream_t>(stream));
}

int kdlae_debug_trace_begin(void) {
  API_BEGIN();
  if (!kd::g_trace_dev) KD_CUDA(cudaMalloc(&kd::g_trace_dev, sizeof(unsigned long long) * 2 * kd::TRACE_MAX));
  KD_CUDA(cudaMemset(kd::g_trace_dev, 0, sizeof(unsigned long long) * 2 * kd::TRACE_MAX));
  kd::g_trace_tags.clear();
  kd::g_trace_on = true;
  return 0;
}

int kdlae_debug_trace_end(unsigned long long* sums, int max_points, char* tags, int tags_bytes) {
  API_BEGIN();
  kd::g_trace_on = false;
  KD_CUDA(cudaDeviceSynchronize());
  const int n = (int)std::min<size_t>(kd::g_trace_tags.size(), (size_t)std::max(0, max_points));
  if (n > 0 && sums) KD_CUDA(cudaMemcpy(sums, kd::g_trace_dev, sizeof(unsigned long long) * 2 * n, cudaMemcpyDeviceToHost));
  if (tags && tags_bytes > 0) {
    std::string all;
    for (int i = 0; i < n; ++i) { all += kd::g_trace_tags[i]; all += '\n'; }
    snprintf(tags, (size_t)tags_bytes, "%s", all.c_str());
  }
  return n;
}

}  // extern "C"
