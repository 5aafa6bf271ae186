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

// Every weight tensor of a model in ONE launch (the op-by-op training path is bound by its launch count: a Keras model has
// a kernel and a bias per layer).  blockIdx.y = tensor; an optional 0 / 1 mask multiplies the gradient first (the MADE
// constraint of tfp's AutoregressiveNetwork: masked kernel entries stay zero).
struct AdamMulti {
  vms_adam_tensor t[VMS_ADAM_MULTI_MAX];
};
__global__ void adam_multi_kernel(const AdamMulti a, float grad_scale, float lr_t, float one_minus_b1, float one_minus_b2,
                                  float eps) {
  const vms_adam_tensor& t = a.t[blockIdx.y];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = t.grad[i];
    if (t.mask) gi *= t.mask[i];
    gi *= grad_scale;
    const float mi = t.m[i] + (gi - t.m[i]) * one_minus_b1;
    const float vi = t.v[i] + (gi * gi - t.v[i]) * one_minus_b2;
    t.m[i] = mi;
    t.v[i] = vi;
    t.theta[i] = t.theta[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// the same with the step count and bias-corrected learning rate in DEVICE memory, so that a captured CUDA graph of a training
// step stays valid from step to step: adam_lr_kernel advances t and writes lr_t, the update kernel reads it
__global__ void adam_lr_kernel(long long* __restrict__ t, double lr, double b1, double b2, float* __restrict__ lr_t) {
  const long long tn = *t + 1;
  *t = tn;
  *lr_t = (float)(lr * sqrt(1.0 - pow(b2, (double)tn)) / (1.0 - pow(b1, (double)tn)));
}
__global__ void adam_multi_dev_kernel(const AdamMulti a, float grad_scale, const float* __restrict__ lr_t_dev, float one_minus_b1,
                                      float one_minus_b2, float eps) {
  const vms_adam_tensor& t = a.t[blockIdx.y];
  const float lr_t = *lr_t_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t.n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = t.grad[i];
    if (t.mask) gi *= t.mask[i];
    gi *= grad_scale;
    const float mi = t.m[i] + (gi - t.m[i]) * one_minus_b1;
    const float vi = t.v[i] + (gi * gi - t.v[i]) * one_minus_b2;
    t.m[i] = mi;
    t.v[i] = vi;
    t.theta[i] = t.theta[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
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

vms_status vms_adam_step_multi(const vms_adam_tensor* tensors, int n_tensors, float grad_scale, int64_t t, double lr,
                               double beta1, double beta2, double eps, vms_stream stream) {
  VMS_REQUIRE(n_tensors >= 0 && t >= 1, VMS_ERR_INVALID_ARG, "adam_step_multi: bad arguments");
  VMS_REQUIRE(n_tensors == 0 || tensors, VMS_ERR_INVALID_ARG, "adam_step_multi: NULL tensor table");
  const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t));
  for (int base = 0; base < n_tensors; base += VMS_ADAM_MULTI_MAX) {
    const int cnt = n_tensors - base < VMS_ADAM_MULTI_MAX ? n_tensors - base : VMS_ADAM_MULTI_MAX;
    AdamMulti a = {};
    int64_t n_max = 0;
    for (int k = 0; k < cnt; ++k) {
      const vms_adam_tensor& e = tensors[base + k];
      VMS_REQUIRE(e.n >= 0 && (e.n == 0 || (e.theta && e.grad && e.m && e.v)), VMS_ERR_INVALID_ARG,
                  "adam_step_multi: NULL pointer in tensor %d", base + k);
      a.t[k] = e;
      if (e.n > n_max) n_max = e.n;
    }
    if (n_max == 0) continue;
    int64_t bx = (n_max + 255) / 256;
    if (bx > 1024) bx = 1024;  // grid-stride beyond
    adam_multi_kernel<<<dim3((unsigned)bx, (unsigned)cnt), 256, 0, as_stream(stream)>>>(
        a, grad_scale, (float)lr_t, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps);
    VMS_LAUNCH_CHECK("adam_multi_kernel");
  }
  return VMS_OK;
}

vms_status vms_adam_step_multi_dev(const vms_adam_tensor* tensors, int n_tensors, float grad_scale, long long* t_dev,
                                   float* lr_t_dev, double lr, double beta1, double beta2, double eps, vms_stream stream) {
  VMS_REQUIRE(n_tensors >= 0 && t_dev && lr_t_dev, VMS_ERR_INVALID_ARG, "adam_step_multi_dev: bad arguments");
  VMS_REQUIRE(n_tensors == 0 || tensors, VMS_ERR_INVALID_ARG, "adam_step_multi_dev: NULL tensor table");
  adam_lr_kernel<<<1, 1, 0, as_stream(stream)>>>(t_dev, lr, beta1, beta2, lr_t_dev);
  VMS_LAUNCH_CHECK("adam_lr_kernel");
  for (int base = 0; base < n_tensors; base += VMS_ADAM_MULTI_MAX) {
    const int cnt = n_tensors - base < VMS_ADAM_MULTI_MAX ? n_tensors - base : VMS_ADAM_MULTI_MAX;
    AdamMulti a = {};
    int64_t n_max = 0;
    for (int k = 0; k < cnt; ++k) {
      const vms_adam_tensor& e = tensors[base + k];
      VMS_REQUIRE(e.n >= 0 && (e.n == 0 || (e.theta && e.grad && e.m && e.v)), VMS_ERR_INVALID_ARG,
                  "adam_step_multi_dev: NULL pointer in tensor %d", base + k);
      a.t[k] = e;
      if (e.n > n_max) n_max = e.n;
    }
    if (n_max == 0) continue;
    int64_t bx = (n_max + 255) / 256;
    if (bx > 1024) bx = 1024;
    adam_multi_dev_kernel<<<dim3((unsigned)bx, (unsigned)cnt), 256, 0, as_stream(stream)>>>(
        a, grad_scale, lr_t_dev, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps);
    VMS_LAUNCH_CHECK("adam_multi_dev_kernel");
  }
  return VMS_OK;
}

}  // extern "C"
