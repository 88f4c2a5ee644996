// Shared helpers for libcodes_b200.so (error handling, activations, launch counting).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/codes_b200.h"

namespace cb {

// thread-local error string + launch counter (the only global state; per thread)
char* error_buffer();
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define CB_CHECK_ARG(cond, ...)                \
  do {                                         \
    if (!(cond)) {                             \
      cb::set_error(__VA_ARGS__);              \
      return CB_ERR_ARG;                       \
    }                                          \
  } while (0)

#define CB_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      cb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                    __LINE__);                                                          \
      return CB_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define CB_LAUNCH_CHECK()                                                               \
  do {                                                                                  \
    cb::count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      cb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),         \
                    __FILE__, __LINE__);                                                \
      return CB_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

inline int num_sms() {
  static thread_local int cached = 0;
  if (cached == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

// Opt the kernel in to the LARGEST dynamic shared-memory size the device allows for it (device opt-in limit minus the
// kernel's static shared memory).  The attribute is per-kernel, process-wide state: setting it to the exact size of
// each launch let one host thread LOWER it between another thread's set and launch ("invalid argument" launch failures
// with several worker threads per GPU, benchmarks/full_benchmark.py) and let a second plan with a smaller tile shrink
// the limit of an existing one.  Always the same maximal value => idempotent and race-free.
inline cudaError_t allow_max_dynamic_smem(const void* kernel) {
  int dev = 0, max_smem = 232448;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) max_smem -= (int)fa.sharedSizeBytes;
  else cudaGetLastError();
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

// ---------------------------------------------------------------- activations
// torch.nn default semantics, fp32.  Accurate variants (erff/tanhf/expf, not the
// fast-math intrinsics) because the fp32 path is the exactness mode.
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float act_fwd(int act, float x, float p) {
  switch (act) {
    case CB_ACT_RELU: return fmaxf(x, 0.f);
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: return x >= 0.f ? x : p * x;
    case CB_ACT_TANH: return tanhf(x);
    case CB_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    case CB_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case CB_ACT_SILU: return x / (1.f + expf(-x));
    case CB_ACT_SOFTPLUS: return softplus_f(x);
    case CB_ACT_MISH: return x * tanhf(softplus_f(x));
    case CB_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}

// derivative w.r.t. the pre-activation x
__device__ __forceinline__ float act_bwd(int act, float x, float p) {
  switch (act) {
    case CB_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: return x >= 0.f ? 1.f : p;
    case CB_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    case CB_ACT_GELU:
      return 0.5f * (1.f + erff(x * 0.70710678118654752440f)) +
             x * expf(-0.5f * x * x) * 0.39894228040143267794f;
    case CB_ACT_ELU: return x > 0.f ? 1.f : expf(x);
    case CB_ACT_SILU: { float s = 1.f / (1.f + expf(-x)); return s * (1.f + x * (1.f - s)); }
    case CB_ACT_SOFTPLUS: return x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
    case CB_ACT_MISH: {
      float sp = softplus_f(x), t = tanhf(sp);
      float s = x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
      return t + x * (1.f - t * t) * s;
    }
    case CB_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-x)); return s * (1.f - s); }
    default: return 1.f;
  }
}

inline bool valid_act(int a) { return a >= CB_ACT_IDENTITY && a <= CB_ACT_PRELU; }

inline int32_t check_desc(const cb_mlp_desc* d) {
  CB_CHECK_ARG(d != nullptr, "cb_mlp_desc is NULL");
  CB_CHECK_ARG(d->n_layers >= 1 && d->n_layers <= CB_MAX_LAYERS, "n_layers=%d out of range [1,%d]",
               d->n_layers, CB_MAX_LAYERS);
  for (int i = 0; i <= d->n_layers; ++i)
    CB_CHECK_ARG(d->dims[i] >= 1 && d->dims[i] <= 65536, "dims[%d]=%d out of range", i, d->dims[i]);
  CB_CHECK_ARG(valid_act(d->act) && valid_act(d->final_act), "unknown activation id");
  return CB_OK;
}

}  // namespace cb
