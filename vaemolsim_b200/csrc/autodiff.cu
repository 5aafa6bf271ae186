// autodiff.cu -- reverse-mode kernels of the op-by-op path: what TF autodiff does, for compositions outside the fused
// ELBO family (von Mises / blockwise / autoregressive decoders, periodic featurisation, MADE conditioners), so that
// `VAE.fit`, `FlowModel.fit` and the reference's model tests (tests/test_models.py:189-262) train on the device.
//
// Replaces TF's gradients of:
//   tfp Normal._log_prob / VonMises._log_prob + parameter transforms       dists.py:56-78, :197-217, :589-610
//   tfp Normal._sample_n (reparameterised) and random_von_mises with its IMPLICIT reparameterisation gradient
//     (von_mises.py: d sample / d concentration = -dF/dconcentration / p(sample), F = von_mises_cdf by Hill's series
//     below concentration 10.5 and a corrected Normal approximation above)                     dists.py:602-610
//   FCDeepNN's periodic featurisation  [x | cos x | sin x]                                     mappings.py:144-149
//   small elementwise pieces of the tape (strided accumulate, scalar broadcast, mask multiply)
// All outputs ACCUMULATE (+=): a tensor may feed several consumers.  One thread per row (or element); these kernels are
// HBM-bound streaming passes over [B, D] tensors with D <= 64.
#include "common.cuh"
#include <math.h>
#include <string.h>

namespace vms {

constexpr int kMaxDofA = 64;
struct BlockwiseSpecA {
  int8_t kind[kMaxDofA];
  int16_t loc[kMaxDofA], loc2[kMaxDofA], scale[kMaxDofA];
};

__device__ __forceinline__ float chbevl(float y, const float* c, int n) {
  float b0 = c[0], b1 = 0.f, b2 = 0.f;
  for (int i = 1; i < n; ++i) { b2 = b1; b1 = b0; b0 = y * b1 - b2 + c[i]; }
  return 0.5f * (b0 - b2);
}

// Cephes single-precision i0ef / i1ef (the routines behind tf.math.bessel_i0e / bessel_i1e for float32)
__device__ float i0e_a(float x) {
  const float A[18] = {-1.30002500998624804212E-8f, 6.04699502254191894932E-8f,  -2.67079385394061173391E-7f,
                       1.11738753912010371815E-6f,  -4.41673835845875056359E-6f, 1.64484480707288970893E-5f,
                       -5.75419501008210370398E-5f, 1.88502885095841655729E-4f,  -5.76375574538582365885E-4f,
                       1.63947561694133579842E-3f,  -4.32430999505057594430E-3f, 1.05464603945949983183E-2f,
                       -2.37374148058994688156E-2f, 4.93052842396707084878E-2f,  -9.49010970480476444210E-2f,
                       1.71620901522208775349E-1f,  -3.04682672343198398683E-1f, 6.76795274409476084995E-1f};
  const float Bc[7] = {3.39623202570838634515E-9f, 2.26666899049817806459E-8f, 2.04891858946906374183E-7f,
                       2.89137052083475648297E-6f, 6.88975834691682398426E-5f, 3.36911647825569408990E-3f,
                       8.04490411014108831608E-1f};
  x = fabsf(x);
  if (x <= 8.0f) return chbevl(0.5f * x - 2.0f, A, 18);
  return chbevl(32.0f / x - 2.0f, Bc, 7) / sqrtf(x);
}
__device__ float i1e_a(float x) {
  const float A[17] = {9.38153738649577178388E-9f,  -4.44505912879632808065E-8f, 2.00329475355213526229E-7f,
                       -8.56872026469545474066E-7f, 3.47025130813767847674E-6f,  -1.32731636560394358279E-5f,
                       4.78156510755005422638E-5f,  -1.61760815825896745588E-4f, 5.12285956168575772895E-4f,
                       -1.51357245063125314899E-3f, 4.15642294431288815669E-3f,  -1.05640848946261981558E-2f,
                       2.47264490306265168283E-2f,  -5.29459812080949914269E-2f, 1.02643658689847095384E-1f,
                       -1.76416518357834055153E-1f, 2.52587186443633654823E-1f};
  const float Bc[7] = {-3.83538038596423702205E-9f, -2.63146884688951950684E-8f, -2.51223623787020892529E-7f,
                       -3.88256480887769039346E-6f, -1.10588938762623716291E-4f, -9.76109749136146840777E-3f,
                       7.78576235018280120474E-1f};
  const float z = fabsf(x);
  float r = z <= 8.0f ? chbevl(0.5f * z - 2.0f, A, 17) * z : chbevl(32.0f / z - 2.0f, Bc, 7) / sqrtf(z);
  return x < 0.f ? -r : r;
}

// loc = atan2(s, c):  d loc / d s = c / (s^2 + c^2),  d loc / d c = -s / (s^2 + c^2)
__device__ __forceinline__ void atan2_grad(float s, float c, float g, float& gs, float& gc) {
  const float r2 = s * s + c * c;
  const float inv = r2 > 0.f ? 1.0f / r2 : 0.f;
  gs = g * c * inv;
  gc = -g * s * inv;
}

// ------------------------------------------------------------------------------------------------ log_prob backward
// lp[b] = sum_d log_prob_d(x[b, d]; params[b, :]);  upstream g_lp[b].  g_x += d lp / d x, g_params += d lp / d params.
__global__ void blockwise_lp_bwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ params,
                                        int64_t ld_p, int64_t B, int D, const BlockwiseSpecA spec, int scale_mode,
                                        const float* __restrict__ g_lp, float* __restrict__ g_x, int64_t ld_gx,
                                        float* __restrict__ g_params, int64_t ld_gp) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* xr = x + b * ld_x;
  const float* pr = params + b * ld_p;
  const float g = g_lp[b];
  for (int d = 0; d < D; ++d) {
    const float raw = pr[spec.scale[d]];
    const float sc = apply_scale(raw, scale_mode);
    const float dsc = apply_scale_grad(raw, scale_mode);
    float gx, gloc, gsc;
    float loc;
    if (spec.kind[d] == VMS_DIST_NORMAL) {
      loc = pr[spec.loc[d]];
      const float z = xr[d] / sc - loc / sc;
      gx = -z / sc;
      gloc = z / sc;
      gsc = (z * z - 1.f) / sc;
    } else {
      loc = spec.loc2[d] >= 0 ? atan2f(pr[spec.loc[d]], pr[spec.loc2[d]]) : pr[spec.loc[d]];
      float sn, cs;
      sincosf(xr[d] - loc, &sn, &cs);
      gx = -sc * sn;
      gloc = sc * sn;
      gsc = cs - i1e_a(sc) / i0e_a(sc);  // d/dk [k (cos - 1) - log i0e(k)] = cos - 1 - (i1e / i0e - 1)
    }
    if (g_x) g_x[b * ld_gx + d] += g * gx;
    if (g_params) {
      float* gp = g_params + b * ld_gp;
      if (spec.kind[d] == VMS_DIST_VONMISES && spec.loc2[d] >= 0) {
        float gs, gc;
        atan2_grad(pr[spec.loc[d]], pr[spec.loc2[d]], g * gloc, gs, gc);
        gp[spec.loc[d]] += gs;
        gp[spec.loc2[d]] += gc;
      } else {
        gp[spec.loc[d]] += g * gloc;
      }
      gp[spec.scale[d]] += g * gsc * dsc;
    }
  }
}

__global__ void std_normal_lp_bwd_kernel(const float* __restrict__ x, int64_t ld_x, int64_t B, int D,
                                         const float* __restrict__ g_lp, float* __restrict__ g_x, int64_t ld_gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  g_x[b * ld_gx + d] += -x[b * ld_x + d] * g_lp[b];
}

// ------------------------------------------------------------------------------------------------ sample backward
// value + derivative with respect to the concentration (what tfp.math.value_and_gradient carries)
struct Dual { float v, d; };
__device__ __forceinline__ Dual dmul(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ Dual ddiv(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
__device__ __forceinline__ Dual dadd(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual dsub(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual dconst(float c) { return {c, 0.f}; }

// d F(x; k) / d k of the von Mises CDF (tfp von_mises_cdf), x in [-pi, pi]
__device__ float vonmises_dcdf_dconc(float x, float k) {
  if (k < 10.5f) {
    float rn = 0.f, drn = 0.f, vn = 0.f, dvn = 0.f;
    for (int n = 20; n > 0; --n) {
      const float fn = (float)n;
      const float den = 2.f * fn / k + rn;
      const float dden = -2.f * fn / (k * k) + drn;
      rn = 1.f / den;
      drn = -dden / (den * den);
      const float mult = sinf(fn * x) / fn + vn;
      dvn = drn * mult + rn * dvn;
      vn = rn * mult;
    }
    const float cdf = 0.5f + x / (2.f * 3.14159265358979323846f) + vn / 3.14159265358979323846f;
    return (cdf >= 0.f && cdf <= 1.f) ? dvn / 3.14159265358979323846f : 0.f;
  }
  const float i0 = i0e_a(k);
  const float dlog = i1e_a(k) / i0 - 1.f;  // d log i0e / d k
  Dual z = {0.7978845608028654f / i0 * sinf(0.5f * x), 0.f};
  z.d = -z.v * dlog;
  const Dual z2 = dmul(z, z), z3 = dmul(z2, z), z4 = dmul(z2, z2);
  const Dual c = {24.f * k, 24.f};
  const Dual a = ddiv(dsub(dsub(c, dmul(dconst(2.f), z2)), dconst(16.f)), dconst(3.f));
  const Dual bn = dadd(dadd(z4, dmul(dconst(1.75f), z2)), dconst(83.5f));
  const Dual bd = dadd(dsub(dsub(c, dconst(56.f)), z2), dconst(3.f));
  const Dual dd = dsub(a, ddiv(bn, bd));
  const Dual xi = dsub(z, ddiv(z3, dmul(dd, dd)));
  return 0.3989422804014327f * expf(-0.5f * xi.v * xi.v) * xi.d;  // d Phi(xi) / d k
}

__device__ __forceinline__ float wrap_pi(float a) {  // to [-pi, pi)
  const float two_pi = 6.283185307179586f;
  a = a - two_pi * floorf((a + 3.14159265358979323846f) / two_pi);
  return a;
}

// z[b, d] is a reparameterised sample of dof d (Normal: eps * scale + loc; von Mises: loc + centred sample, wrapped);
// upstream g_z.  g_params += d z / d params * g_z.
__global__ void blockwise_sample_bwd_kernel(const float* __restrict__ params, int64_t ld_p, int64_t B, int D,
                                            const BlockwiseSpecA spec, int scale_mode, const float* __restrict__ z,
                                            int64_t ld_z, const float* __restrict__ g_z, int64_t ld_gz,
                                            float* __restrict__ g_params, int64_t ld_gp) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* pr = params + b * ld_p;
  float* gp = g_params + b * ld_gp;
  for (int d = 0; d < D; ++d) {
    const float g = g_z[b * ld_gz + d];
    const float raw = pr[spec.scale[d]];
    const float sc = apply_scale(raw, scale_mode);
    const float dsc = apply_scale_grad(raw, scale_mode);
    if (spec.kind[d] == VMS_DIST_NORMAL) {
      const float loc = pr[spec.loc[d]];
      gp[spec.loc[d]] += g;
      gp[spec.scale[d]] += g * ((z[b * ld_z + d] - loc) / sc) * dsc;  // eps recovered from the sample
    } else {
      const bool two = spec.loc2[d] >= 0;
      const float loc = two ? atan2f(pr[spec.loc[d]], pr[spec.loc2[d]]) : pr[spec.loc[d]];
      if (two) {
        float gs, gc;
        atan2_grad(pr[spec.loc[d]], pr[spec.loc2[d]], g, gs, gc);
        gp[spec.loc[d]] += gs;
        gp[spec.loc2[d]] += gc;
      } else {
        gp[spec.loc[d]] += g;
      }
      const float s = wrap_pi(z[b * ld_z + d] - loc);  // the centred sample
      const float inv_prob = expf(-sc * (cosf(s) - 1.f)) * (6.283185307179586f * i0e_a(sc));
      const float ds_dk = -vonmises_dcdf_dconc(s, sc) * inv_prob;
      gp[spec.scale[d]] += g * ds_dk * dsc;
    }
  }
}

// ------------------------------------------------------------------------------------------------ featurisation
// out = [x[:, ~periodic] | cos(x[:, periodic]) | sin(x[:, periodic])]  (mappings.py:144-149);  g_x += d out / d x ^T g_out
__global__ void periodic_bwd_kernel(const float* __restrict__ x, int64_t B, int D, const uint8_t* __restrict__ periodic,
                                    const float* __restrict__ g_out, float* __restrict__ g_x, int n_per) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n_non = D - n_per;
  const float* go = g_out + b * (int64_t)(D + n_per);
  int i_non = 0, i_per = 0;
  for (int d = 0; d < D; ++d) {
    const float v = x[b * (int64_t)D + d];
    if (periodic[d]) {
      float sn, cs;
      sincosf(v, &sn, &cs);
      g_x[b * (int64_t)D + d] += -sn * go[n_non + i_per] + cs * go[n_non + n_per + i_per];
      ++i_per;
    } else {
      g_x[b * (int64_t)D + d] += go[i_non];
      ++i_non;
    }
  }
}

// ------------------------------------------------------------------------------------------------ tape plumbing
__global__ void add_cols_kernel(float* __restrict__ dst, int64_t ld_dst, const float* __restrict__ src, int64_t ld_src,
                                int64_t B, int D, float alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  dst[b * ld_dst + d] += alpha * src[b * ld_src + d];
}
__global__ void add_scalar_kernel(float* __restrict__ dst, int64_t n, const float* __restrict__ scalar, float alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += alpha * (scalar ? scalar[0] : 1.f);
}
__global__ void mul_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] *= src[i];
}
__global__ void sum_all_kernel(const float* __restrict__ src, int64_t n, float alpha, float* __restrict__ out) {
  // fixed-order two-level sum of a small tensor into out[0] (+=): one block
  __shared__ float sh[256];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += src[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += sh[i];
    out[0] += alpha * t;
  }
}

static vms_status fill_spec(BlockwiseSpecA& s, int D, const int32_t* kind, const int32_t* loc_off, const int32_t* loc2_off,
                            const int32_t* scale_off) {
  VMS_REQUIRE(D >= 1 && D <= kMaxDofA && kind && loc_off && scale_off, VMS_ERR_INVALID_ARG,
              "blockwise backward: 1 <= D <= %d and non-NULL layout arrays required", kMaxDofA);
  memset(&s, 0, sizeof(s));
  for (int d = 0; d < D; ++d) {
    s.kind[d] = (int8_t)kind[d];
    s.loc[d] = (int16_t)loc_off[d];
    s.loc2[d] = (int16_t)(loc2_off ? loc2_off[d] : -1);
    s.scale[d] = (int16_t)scale_off[d];
  }
  return VMS_OK;
}

}  // namespace vms

using namespace vms;

extern "C" {

vms_status vms_blockwise_log_prob_backward(const float* x, int64_t ld_x, const float* params, int64_t ld_p, int64_t B, int D,
                                           const int32_t* kind, const int32_t* loc_off, const int32_t* loc2_off,
                                           const int32_t* scale_off, int scale_mode, const float* g_lp, float* g_x,
                                           int64_t ld_gx, float* g_params, int64_t ld_gp, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && x && params && g_lp, VMS_ERR_INVALID_ARG, "blockwise_log_prob_backward: NULL pointer");
  BlockwiseSpecA s;
  vms_status st = fill_spec(s, D, kind, loc_off, loc2_off, scale_off);
  if (st) return st;
  if (B == 0) return VMS_OK;
  blockwise_lp_bwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, ld_x, params, ld_p, B, D, s, scale_mode,
                                                                                       g_lp, g_x, ld_gx, g_params, ld_gp);
  VMS_LAUNCH_CHECK("blockwise_lp_bwd_kernel");
  return VMS_OK;
}

vms_status vms_std_normal_log_prob_backward(const float* x, int64_t ld_x, int64_t B, int D, const float* g_lp, float* g_x,
                                            int64_t ld_gx, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1 && x && g_lp && g_x, VMS_ERR_INVALID_ARG, "std_normal_log_prob_backward: bad arguments");
  if (B == 0) return VMS_OK;
  std_normal_lp_bwd_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld_x, B, D, g_lp, g_x, ld_gx);
  VMS_LAUNCH_CHECK("std_normal_lp_bwd_kernel");
  return VMS_OK;
}

vms_status vms_blockwise_sample_backward(const float* params, int64_t ld_p, int64_t B, int D, const int32_t* kind,
                                         const int32_t* loc_off, const int32_t* loc2_off, const int32_t* scale_off,
                                         int scale_mode, const float* z, int64_t ld_z, const float* g_z, int64_t ld_gz,
                                         float* g_params, int64_t ld_gp, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && params && z && g_z && g_params, VMS_ERR_INVALID_ARG, "blockwise_sample_backward: NULL pointer");
  BlockwiseSpecA s;
  vms_status st = fill_spec(s, D, kind, loc_off, loc2_off, scale_off);
  if (st) return st;
  if (B == 0) return VMS_OK;
  blockwise_sample_bwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(params, ld_p, B, D, s, scale_mode, z,
                                                                                           ld_z, g_z, ld_gz, g_params, ld_gp);
  VMS_LAUNCH_CHECK("blockwise_sample_bwd_kernel");
  return VMS_OK;
}

vms_status vms_periodic_featurise_backward(const float* x, int64_t B, int D, const uint8_t* periodic, int n_periodic,
                                           const float* g_out, float* g_x, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1 && x && periodic && g_out && g_x && n_periodic >= 0 && n_periodic <= D, VMS_ERR_INVALID_ARG,
              "periodic_featurise_backward: bad arguments");
  if (B == 0) return VMS_OK;
  periodic_bwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, B, D, periodic, g_out, g_x, n_periodic);
  VMS_LAUNCH_CHECK("periodic_bwd_kernel");
  return VMS_OK;
}

vms_status vms_add_cols(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t B, int D, float alpha,
                        vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 0 && (B * D == 0 || (dst && src)), VMS_ERR_INVALID_ARG, "add_cols: bad arguments");
  if (B * D == 0) return VMS_OK;
  add_cols_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, as_stream(stream)>>>(dst, ld_dst, src, ld_src, B, D, alpha);
  VMS_LAUNCH_CHECK("add_cols_kernel");
  return VMS_OK;
}

vms_status vms_add_scalar(float* dst, int64_t n, const float* scalar, float alpha, vms_stream stream) {
  VMS_REQUIRE(n >= 0 && (n == 0 || dst), VMS_ERR_INVALID_ARG, "add_scalar: bad arguments");
  if (n == 0) return VMS_OK;
  add_scalar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dst, n, scalar, alpha);
  VMS_LAUNCH_CHECK("add_scalar_kernel");
  return VMS_OK;
}

vms_status vms_mul_inplace(float* dst, const float* src, int64_t n, vms_stream stream) {
  VMS_REQUIRE(n >= 0 && (n == 0 || (dst && src)), VMS_ERR_INVALID_ARG, "mul_inplace: bad arguments");
  if (n == 0) return VMS_OK;
  mul_inplace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dst, src, n);
  VMS_LAUNCH_CHECK("mul_inplace_kernel");
  return VMS_OK;
}

vms_status vms_sum_all(const float* src, int64_t n, float alpha, float* out, vms_stream stream) {
  VMS_REQUIRE(n >= 0 && out && (n == 0 || src), VMS_ERR_INVALID_ARG, "sum_all: bad arguments");
  if (n == 0) return VMS_OK;
  sum_all_kernel<<<1, 256, 0, as_stream(stream)>>>(src, n, alpha, out);
  VMS_LAUNCH_CHECK("sum_all_kernel");
  return VMS_OK;
}

}  // extern "C"
