// Model-level host orchestration: packed-weight layouts and forward schedules.
#pragma once
#include "ops.cuh"
#include "../../include/kdlae_b200.h"

namespace kd {

// Bump allocator over a caller-provided blob; with base == nullptr it only measures.
struct Bump {
  uint8_t* base = nullptr;
  size_t off = 0;
  template <typename U> U* take(size_t n) {
    off = align_up(off, 256);
    U* p = base ? reinterpret_cast<U*>(base + off) : nullptr;
    off += n * sizeof(U);
    return p;
  }
};

template <typename T> int teacher_pack(const kdlae_teacher_cfg& cfg, const float* const* tensors, int n_tensors, void* packed,
                                       size_t packed_bytes, cudaStream_t s);
template <typename T> size_t teacher_packed_bytes(const kdlae_teacher_cfg& cfg);
template <typename T> size_t teacher_workspace_bytes(const kdlae_teacher_cfg& cfg, int mb, int H, int W);
template <typename T> int teacher_forward(const kdlae_teacher_cfg& cfg, const void* packed, const float* img, const float* rate,
                                          int rate_per_image, float* hq, float* sr, int B, int H, int W, int micro_batch, void* ws,
                                          size_t ws_bytes, cudaStream_t s);
int teacher_num_tensors(const kdlae_teacher_cfg& cfg);

template <typename T> int student_pack(const kdlae_student_cfg& cfg, const float* const* tensors, int n_tensors, void* packed,
                                       size_t packed_bytes, cudaStream_t s);
template <typename T> size_t student_packed_bytes(const kdlae_student_cfg& cfg);
template <typename T> size_t student_workspace_bytes(const kdlae_student_cfg& cfg, int mb, int F, int H, int W);
template <typename T> int student_forward(const kdlae_student_cfg& cfg, const void* packed, const float* x, float* y, int B, int F,
                                          int H, int W, int micro_batch, void* ws, size_t ws_bytes, cudaStream_t s);

template <typename T> int asdqe_pack(const kdlae_asdqe_cfg& cfg, const float* const* tensors, int n_tensors, void* packed,
                                     size_t packed_bytes, cudaStream_t s);
template <typename T> size_t asdqe_packed_bytes(const kdlae_asdqe_cfg& cfg);
template <typename T> size_t asdqe_workspace_bytes(const kdlae_asdqe_cfg& cfg, int mb, int H, int W);
template <typename T> int asdqe_forward(const kdlae_asdqe_cfg& cfg, const void* packed, const float* lq, const float* gt,
                                        float* score, float* feat /*nullable*/, int B, int H, int W, int micro_batch, void* ws,
                                        size_t ws_bytes, cudaStream_t s);

template <typename T> int conv_few_in_sized(const SmallConv& op, int Hin, int Win, cudaStream_t s);

}  // namespace kd
