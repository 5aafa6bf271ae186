// batchnorm.cu -- batch statistics and the (de)normalisation coefficients of the two batch-normalisation variants the
// reference can switch on (both default off):
//   tf.keras.layers.BatchNormalization between the Dense layers of FCDeepNN            (mappings.py:113-114)
//   tfp.bijectors.BatchNormalization between the coupling blocks of the flows          (flows.py:308-309, :623-624)
// [TF/TFP-recalled]  training: mean, var = tf.nn.moments(x, axis=0) (two passes, biased variance);
//   normalise     out = x * inv + (beta - mean * inv),  inv = rsqrt(var + eps) * gamma      (tf.nn.batch_normalization)
//   de-normalise  out = x * r + (mean - beta * r),      r = sqrt(var + eps) / gamma         (tfp bijector forward)
//   log-det of the normalising direction (the bijector's INVERSE): sum_d log gamma_d - 0.5 log(var_d + eps)
//   moving statistics: moving = moving * momentum + batch * (1 - momentum)   (vms_axpby)
// The elementwise map itself is vms_affine_cols (reduce.cu) with the per-column scale / shift produced here, so a batch
// norm costs one streaming pass (plus two reduction passes over the batch in training mode).  All sums run in a fixed
// order (deterministic): per-split partials in double, then one thread per column adds the splits.
#include "common.cuh"
#include <math.h>

namespace vms {

namespace {

constexpr int kRowsPerSplit = 1024;

int n_splits(int64_t B) {
  int64_t s = (B + kRowsPerSplit - 1) / kRowsPerSplit;
  return (int)(s < 1 ? 1 : (s > 256 ? 256 : s));
}

// partial[split][d] = sum over the split's rows of (x - shift_d)^p, p = 1 or 2
__global__ void __launch_bounds__(256) col_partial_kernel(const float* __restrict__ x, int64_t ld_x, int64_t B, int D,
                                                          const float* __restrict__ shift, int square,
                                                          double* __restrict__ partial) {
  __shared__ double sh[8][33];
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  const int64_t rows = (B + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows, r1 = min(B, r0 + rows);
  const float sft = (shift && d < D) ? shift[d] : 0.f;
  double acc = 0.0;
  if (d < D)
    for (int64_t r = r0 + slot; r < r1; r += 8) {
      const float v = x[r * ld_x + d] - sft;
      acc += square ? (double)(v * v) : (double)v;
    }
  sh[slot][lane] = acc;
  __syncthreads();
  if (slot == 0 && d < D) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += sh[k][lane];
    partial[(size_t)blockIdx.y * D + d] = s;
  }
}

__global__ void col_final_kernel(const double* __restrict__ partial, int n_split, int D, double inv_n, float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double s = 0.0;
  for (int k = 0; k < n_split; ++k) s += partial[(size_t)k * D + d];
  out[d] = (float)(s * inv_n);
}

__global__ void bn_coeffs_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, int D, float eps,
                                 int denormalize, float* __restrict__ scale, float* __restrict__ shift,
                                 float* __restrict__ ldj) {
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float g = gamma ? gamma[d] : 1.f, b = beta ? beta[d] : 0.f;
    if (!denormalize) {
      const float inv = (1.0f / sqrtf(var[d] + eps)) * g;
      scale[d] = inv;
      shift[d] = b - mean[d] * inv;
    } else {
      const float r = sqrtf(var[d] + eps) / g;
      scale[d] = r;
      shift[d] = mean[d] - b * r;
    }
  }
  if (threadIdx.x == 0 && ldj) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) s += (double)logf(gamma ? gamma[d] : 1.f) - 0.5 * (double)logf(var[d] + eps);
    *ldj = (float)(denormalize ? -s : s);
  }
}

// ---- reverse mode of the NORMALISING direction  y = (x - mean) rsqrt(var + eps) gamma + beta  (Keras layer in training /
// inference mode; the tfp bijector's inverse, whose log-det sum_d log gamma_d - 0.5 log(var_d + eps) is added to every row).
// Column sums in a fixed order (one block per 32 columns, 8 row slots, doubles):
//   s1_d = sum_b g[b, d],   s2_d = sum_b g[b, d] xhat[b, d]      with xhat = (x - mean) rsqrt(var + eps)
__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ g,
                                                          int64_t ld_g, int64_t B, int D, const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps, double* __restrict__ s1,
                                                          double* __restrict__ s2) {
  __shared__ double sh1[8][33], sh2[8][33];
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  double a1 = 0.0, a2 = 0.0;
  if (d < D) {
    const float m = mean[d], r = 1.0f / sqrtf(var[d] + eps);
    for (int64_t b = slot; b < B; b += 8) {
      const float gv = g[b * ld_g + d];
      a1 += (double)gv;
      a2 += (double)(gv * ((x[b * ld_x + d] - m) * r));
    }
  }
  sh1[slot][lane] = a1;
  sh2[slot][lane] = a2;
  __syncthreads();
  if (slot == 0 && d < D) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < 8; ++k) { t1 += sh1[k][lane]; t2 += sh2[k][lane]; }
    s1[d] = t1;
    s2[d] = t2;
  }
}

// g_x += gamma r (g - [batch_stats] (s1 + xhat s2) / B) - [batch_stats] G (x - mean) r^2 / B;   one thread per element.
// Block (0, 0) row 0 also adds the parameter gradients: g_beta += s1, g_gamma += s2 + G / gamma.
__global__ void bn_bwd_apply_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ g, int64_t ld_g, int64_t B,
                                    int D, const float* __restrict__ mean, const float* __restrict__ var,
                                    const float* __restrict__ gamma, float eps, int batch_stats, const double* __restrict__ s1,
                                    const double* __restrict__ s2, const float* __restrict__ G_ptr, float* __restrict__ g_x,
                                    int64_t ld_gx, float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float gm = gamma ? gamma[d] : 1.f;
  const float r = 1.0f / sqrtf(var[d] + eps);
  const float xc = x[b * ld_x + d] - mean[d];
  const float G = G_ptr ? *G_ptr : 0.f;
  float dx = g[b * ld_g + d];
  if (batch_stats) dx -= (float)((s1[d] + (double)(xc * r) * s2[d]) / (double)B);
  dx *= gm * r;
  if (batch_stats) dx -= G * xc * r * r / (float)B;
  g_x[b * ld_gx + d] += dx;
  if (b == 0) {
    if (g_beta) g_beta[d] += (float)s1[d];
    if (g_gamma) g_gamma[d] += (float)s2[d] + G / gm;
  }
}

// ---- cross-replica statistics (data-parallel training: SURVEY 8f-3 "extra allreduce of per-feature mean / var")
// stage 1: buf[d] = n mean_r[d];  stage 2 (mean_g given): buf[d] = n (var_r[d] + (mean_r[d] - mean_g[d])^2);  buf[D] = n.
// After a sum over the ranks, buf[d] / buf[D] is the global mean (stage 1) / the global biased variance (stage 2: the
// parallel-variance combination, no E[x^2] - mean^2 cancellation).
__global__ void bn_sync_pack_kernel(const float* __restrict__ mean_r, const float* __restrict__ var_r,
                                    const float* __restrict__ mean_g, float n, int D, float* __restrict__ buf) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) {
    if (!mean_g) buf[d] = n * mean_r[d];
    else {
      const float dm = mean_r[d] - mean_g[d];
      buf[d] = n * (var_r[d] + dm * dm);
    }
  }
  if (d == D) buf[D] = n;
}
__global__ void bn_sync_unpack_kernel(const float* __restrict__ buf, int D, float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) out[d] = buf[d] / buf[D];
}
// local sums of the reverse mode as floats: [s1 (D) | s2 (D) | G | B]
__global__ void bn_bwd_pack_sums_kernel(const double* __restrict__ s1, const double* __restrict__ s2, const float* __restrict__ G,
                                        float B, int D, float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) { out[d] = (float)s1[d]; out[D + d] = (float)s2[d]; }
  if (d == D) { out[2 * D] = G ? *G : 0.f; out[2 * D + 1] = B; }
}
// g_x += gamma r (g - (S1 + xhat S2) / N) - G_all (x - mean) r^2 / N  with the GLOBAL sums S1, S2, G_all, N (summed over the
// ranks); the parameter gradients take the LOCAL sums (the trainer sums them over the ranks afterwards).
__global__ void bn_bwd_apply_sync_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ g, int64_t ld_g,
                                         int64_t B, int D, const float* __restrict__ mean, const float* __restrict__ var,
                                         const float* __restrict__ gamma, float eps, const float* __restrict__ loc,
                                         const float* __restrict__ glob, float* __restrict__ g_x, int64_t ld_gx,
                                         float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float gm = gamma ? gamma[d] : 1.f;
  const float r = 1.0f / sqrtf(var[d] + eps);
  const float xc = x[b * ld_x + d] - mean[d];
  const float N = glob[2 * D + 1], G_all = glob[2 * D];
  float dx = g[b * ld_g + d] - (glob[d] + (xc * r) * glob[D + d]) / N;
  dx *= gm * r;
  dx -= G_all * xc * r * r / N;
  g_x[b * ld_gx + d] += dx;
  if (b == 0) {
    if (g_beta) g_beta[d] += loc[d];
    if (g_gamma) g_gamma[d] += loc[D + d] + loc[2 * D] / gm;
  }
}

__global__ void broadcast_scalar_kernel(const float* __restrict__ s, int64_t n, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = *s;
}

}  // namespace

}  // namespace vms

using namespace vms;

extern "C" {

size_t vms_batch_moments_workspace(int64_t B, int D) { return (size_t)n_splits(B) * (size_t)(D > 0 ? D : 1) * sizeof(double); }

vms_status vms_batch_moments(const float* x, int64_t ld_x, int64_t B, int D, float* mean, float* var, void* workspace,
                             vms_stream stream) {
  VMS_REQUIRE(x && mean && var && workspace, VMS_ERR_INVALID_ARG, "batch_moments: NULL pointer");
  VMS_REQUIRE(B >= 1 && D >= 1 && ld_x >= D, VMS_ERR_SHAPE, "batch_moments: need B >= 1, D >= 1, ld_x >= D");
  const int ns = n_splits(B);
  cudaStream_t st = as_stream(stream);
  double* part = (double*)workspace;
  dim3 grid((D + 31) / 32, ns);
  col_partial_kernel<<<grid, 256, 0, st>>>(x, ld_x, B, D, nullptr, 0, part);
  VMS_LAUNCH_CHECK("col_partial_kernel");
  col_final_kernel<<<(D + 127) / 128, 128, 0, st>>>(part, ns, D, 1.0 / (double)B, mean);
  VMS_LAUNCH_CHECK("col_final_kernel");
  col_partial_kernel<<<grid, 256, 0, st>>>(x, ld_x, B, D, mean, 1, part);
  VMS_LAUNCH_CHECK("col_partial_kernel");
  col_final_kernel<<<(D + 127) / 128, 128, 0, st>>>(part, ns, D, 1.0 / (double)B, var);
  VMS_LAUNCH_CHECK("col_final_kernel");
  return VMS_OK;
}

vms_status vms_batchnorm_coeffs(const float* mean, const float* var, const float* gamma, const float* beta, int D, float eps,
                                int denormalize, float* scale, float* shift, float* ldj, vms_stream stream) {
  VMS_REQUIRE(mean && var && scale && shift && D >= 1 && eps >= 0.f, VMS_ERR_INVALID_ARG, "batchnorm_coeffs: bad arguments");
  bn_coeffs_kernel<<<1, 256, 0, as_stream(stream)>>>(mean, var, gamma, beta, D, eps, denormalize, scale, shift, ldj);
  VMS_LAUNCH_CHECK("bn_coeffs_kernel");
  return VMS_OK;
}

size_t vms_batchnorm_backward_workspace(int D) { return (size_t)2 * (size_t)(D > 0 ? D : 1) * sizeof(double); }

vms_status vms_batchnorm_backward(const float* x, int64_t ld_x, int64_t B, int D, const float* mean, const float* var,
                                  const float* gamma, float eps, int batch_stats, const float* g_out, int64_t ld_g,
                                  const float* g_ldj_total, float* g_x, int64_t ld_gx, float* g_gamma, float* g_beta,
                                  void* workspace, vms_stream stream) {
  VMS_REQUIRE(x && mean && var && g_out && g_x && workspace, VMS_ERR_INVALID_ARG, "batchnorm_backward: NULL pointer");
  VMS_REQUIRE(B >= 1 && D >= 1, VMS_ERR_SHAPE, "batchnorm_backward: need B >= 1, D >= 1");
  cudaStream_t st = as_stream(stream);
  double* s1 = (double*)workspace;
  double* s2 = s1 + D;
  bn_bwd_sums_kernel<<<(D + 31) / 32, 256, 0, st>>>(x, ld_x, g_out, ld_g, B, D, mean, var, eps, s1, s2);
  VMS_LAUNCH_CHECK("bn_bwd_sums_kernel");
  bn_bwd_apply_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, st>>>(x, ld_x, g_out, ld_g, B, D, mean, var, gamma, eps,
                                                                       batch_stats, s1, s2, g_ldj_total, g_x, ld_gx, g_gamma,
                                                                       g_beta);
  VMS_LAUNCH_CHECK("bn_bwd_apply_kernel");
  return VMS_OK;
}

vms_status vms_bn_sync_pack(const float* mean_r, const float* var_r, const float* mean_g, int64_t n_rows, int D, float* buf,
                            vms_stream stream) {
  VMS_REQUIRE(mean_r && buf && D >= 1 && n_rows >= 0 && (!mean_g || var_r), VMS_ERR_INVALID_ARG, "bn_sync_pack: bad arguments");
  bn_sync_pack_kernel<<<(D + 1 + 127) / 128, 128, 0, as_stream(stream)>>>(mean_r, var_r, mean_g, (float)n_rows, D, buf);
  VMS_LAUNCH_CHECK("bn_sync_pack_kernel");
  return VMS_OK;
}

vms_status vms_bn_sync_unpack(const float* buf, int D, float* out, vms_stream stream) {
  VMS_REQUIRE(buf && out && D >= 1, VMS_ERR_INVALID_ARG, "bn_sync_unpack: bad arguments");
  bn_sync_unpack_kernel<<<(D + 127) / 128, 128, 0, as_stream(stream)>>>(buf, D, out);
  VMS_LAUNCH_CHECK("bn_sync_unpack_kernel");
  return VMS_OK;
}

vms_status vms_batchnorm_backward_sums(const float* x, int64_t ld_x, int64_t B, int D, const float* mean, const float* var,
                                       float eps, const float* g_out, int64_t ld_g, const float* g_ldj_total, float* sums,
                                       void* workspace, vms_stream stream) {
  VMS_REQUIRE(x && mean && var && g_out && sums && workspace, VMS_ERR_INVALID_ARG, "batchnorm_backward_sums: NULL pointer");
  VMS_REQUIRE(B >= 1 && D >= 1, VMS_ERR_SHAPE, "batchnorm_backward_sums: need B >= 1, D >= 1");
  cudaStream_t st = as_stream(stream);
  double* s1 = (double*)workspace;
  double* s2 = s1 + D;
  bn_bwd_sums_kernel<<<(D + 31) / 32, 256, 0, st>>>(x, ld_x, g_out, ld_g, B, D, mean, var, eps, s1, s2);
  VMS_LAUNCH_CHECK("bn_bwd_sums_kernel");
  bn_bwd_pack_sums_kernel<<<(D + 1 + 127) / 128, 128, 0, st>>>(s1, s2, g_ldj_total, (float)B, D, sums);
  VMS_LAUNCH_CHECK("bn_bwd_pack_sums_kernel");
  return VMS_OK;
}

vms_status vms_batchnorm_backward_apply(const float* x, int64_t ld_x, int64_t B, int D, const float* mean, const float* var,
                                        const float* gamma, float eps, const float* g_out, int64_t ld_g, const float* local_sums,
                                        const float* global_sums, float* g_x, int64_t ld_gx, float* g_gamma, float* g_beta,
                                        vms_stream stream) {
  VMS_REQUIRE(x && mean && var && g_out && g_x && local_sums && global_sums, VMS_ERR_INVALID_ARG,
              "batchnorm_backward_apply: NULL pointer");
  VMS_REQUIRE(B >= 1 && D >= 1, VMS_ERR_SHAPE, "batchnorm_backward_apply: need B >= 1, D >= 1");
  bn_bwd_apply_sync_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, as_stream(stream)>>>(
      x, ld_x, g_out, ld_g, B, D, mean, var, gamma, eps, local_sums, global_sums, g_x, ld_gx, g_gamma, g_beta);
  VMS_LAUNCH_CHECK("bn_bwd_apply_sync_kernel");
  return VMS_OK;
}

vms_status vms_broadcast_scalar(const float* scalar, int64_t n, float* out, vms_stream stream) {
  VMS_REQUIRE(scalar && out && n >= 0, VMS_ERR_INVALID_ARG, "broadcast_scalar: bad arguments");
  if (n == 0) return VMS_OK;
  broadcast_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(scalar, n, out);
  VMS_LAUNCH_CHECK("broadcast_scalar_kernel");
  return VMS_OK;
}

}  // extern "C"
