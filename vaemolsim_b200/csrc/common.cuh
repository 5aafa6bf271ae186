// common.cuh -- shared helpers for libvms_b200 (error channel, launch accounting, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/vms_b200.h"

namespace vms {

// thread-local error message returned by vms_last_error()
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(vms_stream s) { return reinterpret_cast<cudaStream_t>(s); }

// SM count of the current device (cached per device)
int sm_count();
int max_smem_optin();

}  // namespace vms

#define VMS_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      vms::set_error(__VA_ARGS__);        \
      return (code);                      \
    }                                     \
  } while (0)

#define VMS_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      vms::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VMS_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

// after a kernel launch: pick up launch-configuration errors synchronously
#define VMS_LAUNCH_CHECK(name)                                                     \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      vms::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
      return VMS_ERR_CUDA;                                                         \
    }                                                                              \
    vms::count_launch();                                                           \
  } while (0)

namespace vms {

// NVTX ranges around the C-ABI entry points of the hot path (timeline annotation for Nsight Systems / Compute, SURVEY 5
// "tracing"): active when VMS_NVTX=1 is set in the environment, otherwise one predictable branch.
bool nvtx_enabled();
void nvtx_push(const char* name);
void nvtx_pop();
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_enabled()) {
    if (on) nvtx_push(name);
  }
  ~NvtxRange() {
    if (on) nvtx_pop();
  }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define VMS_RANGE(name) vms::NvtxRange _vms_range(name)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// tf.math.softplus (Eigen): x > -thr -> x ; x < thr -> exp(x) ; else log1p(exp(x));  thr = log(eps32) + 2
__device__ __forceinline__ float softplus_tf(float x) {
  const float thr = -13.942385f;
  if (x > -thr) return x;
  float e = expf(x);
  if (x < thr) return e;
  return log1pf(e);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

#define VMS_EPS32 1.1920928955078125e-07f
#define VMS_HALF_LOG_2PI 0.91893853320467274178f
#define VMS_LOG_2PI 1.83787706640934548356f

// tfp Normal._log_prob: -0.5 (x/s - m/s)^2 - (0.5 log 2pi + log s)
__device__ __forceinline__ float normal_lp(float x, float loc, float scale) {
  float z = x / scale - loc / scale;
  return -0.5f * z * z - (VMS_HALF_LOG_2PI + logf(scale));
}

__device__ __forceinline__ float apply_scale(float raw, int mode) {
  if (mode == VMS_SCALE_IDENTITY) return raw;
  float s = softplus_tf(raw);
  return mode == VMS_SCALE_SOFTPLUS_EPS ? s + VMS_EPS32 : s;
}
// d scale / d raw
__device__ __forceinline__ float apply_scale_grad(float raw, int mode) {
  return mode == VMS_SCALE_IDENTITY ? 1.0f : sigmoidf_(raw);
}

}  // namespace vms
