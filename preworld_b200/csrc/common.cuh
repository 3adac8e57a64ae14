// Shared helpers for the preworld_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PW_API extern "C" __attribute__((visibility("default")))

#define PW_ERR_INVALID_ARGUMENT (-1)

#define PW_REQUIRE(cond)                         \
  do {                                           \
    if (!(cond)) return PW_ERR_INVALID_ARGUMENT; \
  } while (0)

// Launch check: kernels are asynchronous; this reports configuration errors.
#define PW_LAUNCH_CHECK()                        \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

// launch accounting (elementwise.cu); feeds pw_launch_count()
void pw_count_launch(int k);

static inline int pw_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

enum PwAct { PW_ACT_NONE = 0, PW_ACT_RELU = 1, PW_ACT_SOFTPLUS = 2, PW_ACT_SIGMOID = 3,
             PW_ACT_GELU = 4 };

__device__ __forceinline__ float pw_softplus(float v) {
  // torch.nn.Softplus(beta=1, threshold=20)
  return v > 20.f ? v : log1pf(expf(v));
}

__device__ __forceinline__ float pw_activate(float v, int act) {
  switch (act) {
    case PW_ACT_RELU: return fmaxf(v, 0.f);
    case PW_ACT_SOFTPLUS: return pw_softplus(v);
    case PW_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    // torch.nn.GELU() (approximate='none'): x * 0.5 * (1 + erf(x / sqrt(2)))
    case PW_ACT_GELU: return v * 0.5f * (1.f + erff(v * 0.70710678118654752440f));
    default: return v;
  }
}

// out-of-line variant for epilogues that are almost always ReLU / none: keeps
// the transcendental code out of the unrolled store loops
__device__ __noinline__ static float pw_activate_slow(float v, int act) {
  return pw_activate(v, act);
}

__device__ __forceinline__ float4 pw_ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
