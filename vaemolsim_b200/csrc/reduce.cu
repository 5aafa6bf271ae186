// reduce.cu -- K5: ELBO / KL batch reductions (deterministic two-level warp-shuffle sums) and small elementwise ops.
//
// Replaces losses.py:253 (`KLDivergenceEstimate.call`: reduce_mean(log a - log b)), losses.py:296 / :330 (the
// LogProbRegularizer / reverse-KL variants, expressed through the same kernel by argument order and sign),
// losses.py:58 + Keras' batch-mean reduction (`LogProbLoss`), and the tensor `+` of mcmc.py:103,109.
#include "common.cuh"

namespace vms {

constexpr int RT = 256;        // threads per block
constexpr int kChunk = 8192;   // elements per block: fixed partition => result independent of grid / SM count

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sh[RT / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = threadIdx.x < RT / 32 ? sh[threadIdx.x] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid in warp 0
}

// partial[blockIdx] = sum over the block's chunk of (a - b); single-block case writes scale * sum / B directly.
__global__ void __launch_bounds__(RT) diff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      int64_t B, float* __restrict__ partial, float* out, float scale) {
  const int64_t beg = (int64_t)blockIdx.x * kChunk;
  const int64_t end = min(B, beg + kChunk);
  float s = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += RT) s += b ? a[i] - b[i] : a[i];
  s = block_sum(s);
  if (threadIdx.x == 0) {
    if (gridDim.x == 1) out[0] = scale * (s / (float)B);
    else partial[blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(RT) final_sum_kernel(const float* __restrict__ partial, int n, int64_t B, float* out,
                                                       float scale) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += RT) s += partial[i];
  s = block_sum(s);
  if (threadIdx.x == 0) out[0] = scale * (s / (float)B);
}

vms_status mean_diff(const float* a, const float* b, int64_t B, float scale, float* out, cudaStream_t st) {
  VMS_REQUIRE(a && out, VMS_ERR_INVALID_ARG, "mean reduction: NULL pointer");
  VMS_REQUIRE(B >= 1, VMS_ERR_SHAPE, "mean reduction over an empty batch");
  const int64_t nb = (B + kChunk - 1) / kChunk;
  VMS_REQUIRE(nb < (1LL << 31), VMS_ERR_SHAPE, "batch too large");
  float* partial = nullptr;
  if (nb > 1) VMS_CUDA(cudaMallocAsync((void**)&partial, nb * sizeof(float), st));
  diff_sum_kernel<<<(unsigned)nb, RT, 0, st>>>(a, b, B, partial, out, scale);
  VMS_LAUNCH_CHECK("diff_sum_kernel");
  if (nb > 1) {
    final_sum_kernel<<<1, RT, 0, st>>>(partial, (int)nb, B, out, scale);
    VMS_LAUNCH_CHECK("final_sum_kernel");
    VMS_CUDA(cudaFreeAsync(partial, st));
  }
  return VMS_OK;
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b, int64_t n,
                             float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}

// out[b, d] = x[b, d] * scale[d] + shift[d]  (tfp.bijectors.Shift / Scale of flows.py:53-58); inverse handled by the
// host passing 1/scale and -shift/scale.
__global__ void affine_cols_kernel(const float* __restrict__ x, int64_t ld_x, int64_t B, int D,
                                   const float* __restrict__ scale, const float* __restrict__ shift, int shift_first,
                                   float* __restrict__ out, int64_t ld_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float v = x[b * ld_x + d];
  const float sc = scale ? scale[d] : 1.f, sh = shift ? shift[d] : 0.f;
  out[b * ld_out + d] = shift_first ? __fmul_rn(__fadd_rn(v, sh), sc) : __fadd_rn(__fmul_rn(v, sc), sh);
}

}  // namespace vms

using namespace vms;

extern "C" {

vms_status vms_kl_mean(const float* lq, const float* lp, int64_t B, float weight, float* out, vms_stream stream) {
  return mean_diff(lq, lp, B, weight, out, as_stream(stream));
}
vms_status vms_scaled_mean(const float* v, int64_t B, float scale, float* out, vms_stream stream) {
  return mean_diff(v, nullptr, B, scale, out, as_stream(stream));
}
vms_status vms_axpby(const float* x, const float* y, float a, float b, int64_t n, float* out, vms_stream stream) {
  VMS_REQUIRE(n >= 0 && (n == 0 || (x && out)), VMS_ERR_INVALID_ARG, "axpby: NULL pointer");
  if (n == 0) return VMS_OK;
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, a, b, n, out);
  VMS_LAUNCH_CHECK("axpby_kernel");
  return VMS_OK;
}

vms_status vms_affine_cols(const float* x, int64_t ld_x, int64_t B, int D, const float* scale, const float* shift,
                           int shift_first, float* out, int64_t ld_out, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "affine_cols: bad shape");
  VMS_REQUIRE(B == 0 || (x && out), VMS_ERR_INVALID_ARG, "affine_cols: NULL pointer");
  if (B == 0) return VMS_OK;
  affine_cols_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld_x, B, D, scale, shift,
                                                                                    shift_first, out, ld_out);
  VMS_LAUNCH_CHECK("affine_cols_kernel");
  return VMS_OK;
}

}  // extern "C"
