// Shared host/device helpers: error plumbing, activation table, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/shifu_b200.h"

namespace sb {

// ---- thread-local last error (sb_last_error) ----
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);

#define SB_CUDA(call)                                                                                       \
  do {                                                                                                      \
    cudaError_t _e = (call);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return ::sb::set_error(SB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, \
                             __LINE__);                                                                     \
  } while (0)

#define SB_CHECK(cond, code, ...)                          \
  do {                                                     \
    if (!(cond)) return ::sb::set_error(code, __VA_ARGS__); \
  } while (0)

#define SB_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != SB_OK) return _s; \
  } while (0)

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- activations (get_activation_fun, res/ssgd_monitor.py:74-88; tf.nn.leaky_relu alpha = 0.2) ----
#define SB_LEAKY_ALPHA 0.2f

__device__ __forceinline__ float sigmoidf_stable(float z) {
  // same two-branch form as the oracle; expf (not __expf) to stay within 1e-6 of fp32 libm
  if (z >= 0.f) return 1.f / (1.f + expf(-z));
  float e = expf(z);
  return e / (1.f + e);
}

__device__ __forceinline__ float act_apply(float z, int act) {
  switch (act) {
    case SB_ACT_SIGMOID: return sigmoidf_stable(z);
    case SB_ACT_TANH: return tanhf(z);
    case SB_ACT_RELU: return fmaxf(z, 0.f);
    case SB_ACT_LEAKYRELU: return z > 0.f ? z : z * SB_LEAKY_ALPHA;
    default: return z;  // SB_ACT_NONE
  }
}
// d act/dz written in terms of the activation OUTPUT a (TF's SigmoidGrad/TanhGrad/ReluGrad convention)
__device__ __forceinline__ float act_grad_from_out(float a, int act) {
  switch (act) {
    case SB_ACT_SIGMOID: return a * (1.f - a);
    case SB_ACT_TANH: return 1.f - a * a;
    case SB_ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case SB_ACT_LEAKYRELU: return a > 0.f ? 1.f : SB_LEAKY_ALPHA;
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums of a 32(lanes = rows) x 32(registers = columns) fp32 fragment.
// On return lane j holds sum over the 32 rows of column j.  31 shuffles instead of 160.
__device__ __forceinline__ float warp_colsum_32x32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

}  // namespace sb
