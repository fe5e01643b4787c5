// Shared helpers for libpeagnn_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/peagnn.h"

namespace peagnn {

void set_error(const char* fmt, ...);

extern unsigned long long g_launches;  // kernels launched by this library in this process

inline int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return PEAGNN_ERR_CUDA;
  }
  return PEAGNN_OK;
}

#define PEAGNN_REQUIRE(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      peagnn::set_error(__VA_ARGS__);    \
      return PEAGNN_ERR_ARG;             \
    }                                    \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200

__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// offset of row c in a table with leading dimension ld: one IMAD.WIDE.U32 instead of a 64x64 multiply
__device__ __forceinline__ size_t row_off(int c, unsigned ld) {
  return (size_t)((unsigned long long)(unsigned)c * (unsigned long long)ld);
}

__device__ __forceinline__ float leaky(float s, float slope) { return s > 0.f ? s : slope * s; }

// -log(sigmoid(z)) = softplus(-z), evaluated without overflow.  The reference spells it
// sigmoid().log() (models/base.py:48), which torch keeps finite down to z ~ -103 (denormal sigmoid);
// a literal 1/(1+expf(-z)) already overflows at z < -88.7 and would turn one outlier triple into an
// infinite summed loss.  This form agrees with the reference to fp32 rounding wherever the reference is
// finite and stays finite beyond.
__device__ __forceinline__ float neg_log_sigmoid(float z) {
  return fmaxf(-z, 0.f) + log1pf(expf(-fabsf(z)));
}
// d/dz of the above = sigmoid(z) - 1 = -1 / (1 + e^z); expf(z) -> inf gives -0, -> 0 gives -1
__device__ __forceinline__ float neg_log_sigmoid_grad(float z) { return -1.f / (1.f + expf(z)); }

// Grouped projection launches (peagnn_linear_grouped): the problem table travels by value in the kernel parameters
// (it is captured with the launch by a CUDA graph); CTA b works on the problem whose block range holds b.
struct LinearGroup {
  int count;
  int block_start[PEAGNN_MAX_GROUP + 1];
  peagnn_linear_problem_t p[PEAGNN_MAX_GROUP];
};
__device__ __forceinline__ int group_of_block(const LinearGroup& g, int b) {
  int k = 0;
  while (k + 1 < g.count && b >= g.block_start[k + 1]) ++k;
  return k;
}

struct WgradGroup {
  int count;
  int block_start[PEAGNN_MAX_GROUP + 1];
  int64_t rows_per_cta[PEAGNN_MAX_GROUP];
  float* partial[PEAGNN_MAX_GROUP];      // this problem's [parts][K*M + M] slice of the workspace
  peagnn_wgrad_problem_t p[PEAGNN_MAX_GROUP];
};
__device__ __forceinline__ int group_of_block(const WgradGroup& g, int b) {
  int k = 0;
  while (k + 1 < g.count && b >= g.block_start[k + 1]) ++k;
  return k;
}

}  // namespace peagnn
