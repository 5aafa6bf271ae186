// adam.cu -- K8: Keras Adam on the flat parameter buffer, fused with the fixed-order sum of partial gradients.
//
// Replaces `tf.keras.optimizers.Adam(learning_rate=1e-3)` as invoked by Model.fit at tests/test_models.py:181-182:
//   lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);
//   theta -= lr_t m / (sqrt(v) + eps)
#include "common.cuh"
#include <math.h>

namespace vms {

__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ g, int n_partials, int64_t stride,
                            float grad_scale, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr_t,
                            float one_minus_b1, float one_minus_b2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = 0.f;
  for (int j = 0; j < n_partials; ++j) gi += g[(int64_t)j * stride + i];
  gi *= grad_scale;
  const float mi = m[i] + (gi - m[i]) * one_minus_b1;
  const float vi = v[i] + (gi * gi - v[i]) * one_minus_b2;
  m[i] = mi;
  v[i] = vi;
  theta[i] = theta[i] - lr_t * mi / (sqrtf(vi) + eps);
}

}  // namespace vms

using namespace vms;

extern "C" {

vms_status vms_adam_step(float* theta, const float* g, int n_partials, float grad_scale, float* m, float* v, int64_t n,
                         int64_t t, double lr, double beta1, double beta2, double eps, vms_stream stream) {
  VMS_REQUIRE(n >= 0 && t >= 1 && n_partials >= 1, VMS_ERR_INVALID_ARG, "adam_step: bad arguments");
  VMS_REQUIRE(n == 0 || (theta && g && m && v), VMS_ERR_INVALID_ARG, "adam_step: NULL pointer");
  if (n == 0) return VMS_OK;
  const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(
      theta, g, n_partials, n, grad_scale, m, v, n, (float)lr_t, (float)(1.0 - beta1),
      (float)(1.0 - beta2), (float)eps);
  VMS_LAUNCH_CHECK("adam_kernel");
  return VMS_OK;
}

}  // extern "C"
