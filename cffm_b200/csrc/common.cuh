// Shared device helpers for the CFFM hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/cffm.h"

namespace cffm {

constexpr float kSeluScale = 1.0507009873554805f;
constexpr float kSeluAlpha = 1.6732632423543772f;
constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kInvSqrt2Pi = 0.39894228040143267794f;

// ---- self.activation (CFFM.py:132-141, :149-155) -------------------------------------------
template <int ACT>
__device__ __forceinline__ float act_f(float x) {
  if constexpr (ACT == CFFM_ACT_RELU) return fmaxf(x, 0.f);
  if constexpr (ACT == CFFM_ACT_ELU) return x > 0.f ? x : expm1f(x);
  if constexpr (ACT == CFFM_ACT_SELU) return x > 0.f ? kSeluScale * x : kSeluScale * kSeluAlpha * expm1f(x);
  if constexpr (ACT == CFFM_ACT_PRELU) return x > 0.f ? x : 0.25f * x;
  if constexpr (ACT == CFFM_ACT_GELU) return x * (0.5f * (1.f + erff(x * kInvSqrt2)));
  return x;
}
// d act / dx  ([TF-1.14] gradient kernels: negative branch only for x < 0; relu'(0) = 0)
template <int ACT>
__device__ __forceinline__ float act_df(float x) {
  if constexpr (ACT == CFFM_ACT_RELU) return x > 0.f ? 1.f : 0.f;
  if constexpr (ACT == CFFM_ACT_ELU) return x < 0.f ? expf(x) : 1.f;
  if constexpr (ACT == CFFM_ACT_SELU) return x < 0.f ? kSeluScale * kSeluAlpha * expf(x) : kSeluScale;
  if constexpr (ACT == CFFM_ACT_PRELU) return x > 0.f ? 1.f : (x < 0.f ? 0.25f : 0.f);
  if constexpr (ACT == CFFM_ACT_GELU)
    return 0.5f * (1.f + erff(x * kInvSqrt2)) + x * kInvSqrt2Pi * expf(-0.5f * x * x);
  return 1.f;
}
// phi(y) = activation(relu(y)): conv_layer always applies relu, the caller then applies
// self.activation (CFFM.py:475-478 + :330 / :387).
template <int ACT>
__device__ __forceinline__ float phi_f(float y) {
  float r = fmaxf(y, 0.f);
  if constexpr (ACT == CFFM_ACT_SELU) return kSeluScale * r;
  if constexpr (ACT == CFFM_ACT_GELU) return r * (0.5f * (1.f + erff(r * kInvSqrt2)));
  return r;  // relu, elu, prelu are the identity on r >= 0
}
template <int ACT>
__device__ __forceinline__ float phi_df(float y) {
  if (!(y > 0.f)) return 0.f;
  if constexpr (ACT == CFFM_ACT_SELU) return kSeluScale;
  if constexpr (ACT == CFFM_ACT_GELU)
    return 0.5f * (1.f + erff(y * kInvSqrt2)) + y * kInvSqrt2Pi * expf(-0.5f * y * y);
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Dispatch a runtime activation id to a template instantiation.
#define CFFM_DISPATCH_ACT(act, ...)                                         \
  switch (act) {                                                            \
    case CFFM_ACT_RELU: { constexpr int ACT = CFFM_ACT_RELU; __VA_ARGS__; } break;   \
    case CFFM_ACT_ELU: { constexpr int ACT = CFFM_ACT_ELU; __VA_ARGS__; } break;     \
    case CFFM_ACT_SELU: { constexpr int ACT = CFFM_ACT_SELU; __VA_ARGS__; } break;   \
    case CFFM_ACT_PRELU: { constexpr int ACT = CFFM_ACT_PRELU; __VA_ARGS__; } break; \
    default: { constexpr int ACT = CFFM_ACT_GELU; __VA_ARGS__; } break;              \
  }

// The conv stack only sees the activation through phi = activation o relu, and relu / elu / prelu are
// the identity on relu's range: three distinct instantiations instead of five.
#define CFFM_DISPATCH_PHI(act, ...)                                         \
  switch (act) {                                                            \
    case CFFM_ACT_SELU: { constexpr int ACT = CFFM_ACT_SELU; __VA_ARGS__; } break;   \
    case CFFM_ACT_GELU: { constexpr int ACT = CFFM_ACT_GELU; __VA_ARGS__; } break;   \
    default: { constexpr int ACT = CFFM_ACT_RELU; __VA_ARGS__; } break;              \
  }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Device allocations of the library go through these two.  With CFFM_GUARD=1 in the environment (read once) every
// allocation gets a 4 KB band on either side filled with a pattern; cffm_debug_check_guards() verifies the bands, i.e.
// that no kernel wrote just outside one of its buffers (compute-sanitizer is closed on this GPU pool).
cudaError_t dev_malloc(void** p, size_t bytes);
cudaError_t dev_free(void* p);
// number of allocations whose bands are damaged (and a description of the first) -- synchronises the device
int dev_check_guards(char* msg, int cap);
int dev_guard_selftest();   // 0: an overrun by one element on either side is detected; -1: guard off

// cudaFuncSetAttribute is per device (context): a "done" flag has to be kept per device ordinal, or a second
// handle on another GPU of the same process launches without its shared-memory opt-in.
struct PerDeviceOnce {
  bool done[64] = {};
  bool& operator()() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) d = 0;
    return done[d];
  }
};

}  // namespace cffm
