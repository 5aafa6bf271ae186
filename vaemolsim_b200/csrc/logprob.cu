// logprob.cu -- K4: distribution log_prob / reparameterised sample kernels and their reverse mode.
//
// Replaces (third-party arithmetic reached from the reference at the cited call sites):
//   tfp Normal._log_prob      = -0.5 (x/s - m/s)^2 - (0.5 log 2pi + log s)        dists.py:213-217, tests/test_models.py:167-170
//   tfp VonMises._log_prob    = k (cos(x - m) - 1) - log 2pi - log i0e(k)         dists.py:602-610, :64-71
//   tfp Normal._sample_n      = eps * s + m                                        models.py:310, mcmc.py:100-102
//   parameter transforms      softplus / softplus + eps32 / atan2                  dists.py:56-78, :603-607
// One thread per row: the event size is 1..64 and the op is HBM-bound (3*D*4 + 4 bytes per row).
#include "common.cuh"
#include <math.h>

namespace vms {

constexpr int kMaxDof = 64;
struct BlockwiseSpec {
  int8_t kind[kMaxDof];
  int16_t loc[kMaxDof], loc2[kMaxDof], scale[kMaxDof];
};

// Cephes single-precision exponentially scaled modified Bessel function I0 (i0ef), the routine behind
// tf.math.bessel_i0e for float32 (Eigen generic_i0e<float>): Chebyshev series on [0,8] and (8,inf).
__device__ __forceinline__ float i0e_f(float x) {
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
  if (x <= 8.0f) {
    float y = 0.5f * x - 2.0f, b0 = A[0], b1 = 0.f, b2 = 0.f;
#pragma unroll
    for (int i = 1; i < 18; ++i) { b2 = b1; b1 = b0; b0 = y * b1 - b2 + A[i]; }
    return 0.5f * (b0 - b2);
  }
  float y = 32.0f / x - 2.0f, b0 = Bc[0], b1 = 0.f, b2 = 0.f;
#pragma unroll
  for (int i = 1; i < 7; ++i) { b2 = b1; b1 = b0; b0 = y * b1 - b2 + Bc[i]; }
  return 0.5f * (b0 - b2) / sqrtf(x);
}

__global__ void blockwise_lp_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ params,
                                    int64_t ld_p, int64_t B, int D, const BlockwiseSpec spec, int scale_mode,
                                    float* __restrict__ lp, int accumulate) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* xr = x + b * ld_x;
  const float* pr = params + b * ld_p;
  float s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float sc = apply_scale(pr[spec.scale[d]], scale_mode);
    if (spec.kind[d] == VMS_DIST_NORMAL) {
      s += normal_lp(xr[d], pr[spec.loc[d]], sc);
    } else {
      const float loc = spec.loc2[d] >= 0 ? atan2f(pr[spec.loc[d]], pr[spec.loc2[d]]) : pr[spec.loc[d]];
      s += sc * (cosf(xr[d] - loc) - 1.0f) - VMS_LOG_2PI - logf(i0e_f(sc));
    }
  }
  lp[b] = accumulate ? lp[b] + s : s;
}

// Fast path of blockwise_lp_kernel for the layout every Normal head of the path uses (tfp.layers.IndependentNormal,
// tests/test_models.py:167-170): all dofs Normal, params = [loc_0..loc_{D-1} | raw_0..raw_{D-1}], contiguous rows.
// ncu on the generic kernel: it is ISSUE-bound (about 90 instructions per dof: two IEEE divisions, logf, expf and
// log1pf behind softplus), 41 % of the HBM copy peak.  Here the two divisions by the scale become one correctly
// rounded reciprocal and two multiplies (<= 1.5 ulp on z), log(scale) uses the fast hardware log2 (|error| < 1e-6
// absolute; scale = softplus(raw) > 0), D is a template parameter so the dof loop is straight-line code, and the row's
// x and params are fetched with 8- / 16-byte loads.  (A shared-memory-staged "fully coalesced" variant measured
// SLOWER, 22 %: the per-element index arithmetic of the staging cost more than the sector inefficiency it removed.)
template <int D>
__global__ void __launch_bounds__(128) normal_rows_lp_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                                             int64_t B, int scale_mode, float* __restrict__ lp,
                                                             int accumulate) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float xr[D], pr[2 * D];
  if (D % 2 == 0) {
    const float2* x2 = reinterpret_cast<const float2*>(x + b * D);
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      const float2 t = __ldg(x2 + i);
      xr[2 * i] = t.x; xr[2 * i + 1] = t.y;
    }
    const float4* p4 = reinterpret_cast<const float4*>(params + b * 2 * D);
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      const float4 t = __ldg(p4 + i);
      pr[4 * i] = t.x; pr[4 * i + 1] = t.y; pr[4 * i + 2] = t.z; pr[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < D; ++i) xr[i] = __ldg(x + b * D + i);
#pragma unroll
    for (int i = 0; i < 2 * D; ++i) pr[i] = __ldg(params + b * 2 * D + i);
  }
  float s = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float sc = apply_scale(pr[D + d], scale_mode);
    const float inv = __frcp_rn(sc);
    const float z = fmaf(xr[d], inv, -(pr[d] * inv));
    s += -0.5f * z * z - (VMS_HALF_LOG_2PI + __logf(sc));
  }
  lp[b] = accumulate ? lp[b] + s : s;
}

// constrained parameters of every dof: loc [B, D] and scale / concentration [B, D] (make_param_transform, dists.py:28-87)
__global__ void blockwise_params_kernel(const float* __restrict__ params, int64_t ld_p, int64_t B, int D,
                                        const BlockwiseSpec spec, int scale_mode, float* __restrict__ loc,
                                        float* __restrict__ scale) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float* pr = params + b * ld_p;
  scale[i] = apply_scale(pr[spec.scale[d]], scale_mode);
  loc[i] = (spec.kind[d] == VMS_DIST_VONMISES && spec.loc2[d] >= 0) ? atan2f(pr[spec.loc[d]], pr[spec.loc2[d]])
                                                                   : pr[spec.loc[d]];
}

__global__ void std_normal_lp_kernel(const float* __restrict__ x, int64_t ld_x, int64_t B, int D,
                                     float* __restrict__ lp, int accumulate) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* xr = x + b * ld_x;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += normal_lp(xr[d], 0.f, 1.f);
  lp[b] = accumulate ? lp[b] + s : s;
}

__global__ void normal_sample_lp_kernel(const float* __restrict__ params, int64_t ld_p, int loc_off, int scale_off,
                                        int scale_mode, const float* __restrict__ eps, int64_t B, int D,
                                        float* __restrict__ z, int64_t ld_z, float* __restrict__ lp) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* pr = params + b * ld_p;
  float s = 0.f;
  for (int d = 0; d < D; ++d) {
    const float loc = pr[loc_off + d], sc = apply_scale(pr[scale_off + d], scale_mode);
    const float zz = __fadd_rn(__fmul_rn(eps[b * D + d], sc), loc);  // separate TF mul and add ops: no FMA contraction
    if (z) z[b * ld_z + d] = zz;
    s += normal_lp(zz, loc, sc);
  }
  if (lp) lp[b] = s;
}

__global__ void normal_lp_bwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ params,
                                     int64_t ld_p, int loc_off, int scale_off, int scale_mode,
                                     const float* __restrict__ g_lp, int64_t B, int D, float* g_x, int64_t ld_gx,
                                     int accumulate_x, float* __restrict__ g_params, int64_t ld_gp) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float g = g_lp[b];
  const float* pr = params + b * ld_p;
  for (int d = 0; d < D; ++d) {
    const float raw = pr[scale_off + d];
    const float loc = pr[loc_off + d], sc = apply_scale(raw, scale_mode);
    const float u = x[b * ld_x + d] / sc - loc / sc;
    if (g_x) {
      float* dst = g_x + b * ld_gx + d;
      const float v = g * (-u / sc);
      *dst = accumulate_x ? *dst + v : v;
    }
    if (g_params) {
      g_params[b * ld_gp + loc_off + d] = g * (u / sc);
      g_params[b * ld_gp + scale_off + d] = g * ((u * u - 1.0f) / sc) * apply_scale_grad(raw, scale_mode);
    }
  }
}

// ------------------------------------------------------------------------------------------------ sampling
// Philox4x32-10 (Salmon et al. 2011), counter = (row, dof, attempt), key = seed: samples do not depend on the launch
// geometry or the rank count.  (TF's own RNG streams cannot be reproduced outside TF; sampling parity is distributional.)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(unsigned a) { return ((float)a + 0.5f) * 2.3283064365386963e-10f; }  // (0, 1)

// One sample per (row, dof) of a Blockwise distribution (dists.py:210-217, :326-336, :602-610 -> tfp sample):
//   Normal     z = eps * scale + loc   (tfp Normal._sample_n; eps given, or Philox + Box-Muller)
//   von Mises  tfp random_von_mises: Best & Fisher (1979) rejection sampler with a wrapped-Cauchy envelope,
//              s = (1 + rho^2) / (2 rho) for concentration > 2e-2 (float32 cut-off) else 1 / concentration;
//              sample = sign(u) acos(w) + loc, wrapped to [-pi, pi) by x - 2 pi round(x / 2 pi)   [TFP-recalled]
__global__ void blockwise_sample_kernel(const float* __restrict__ params, int64_t ld_p, int64_t B, int D,
                                        const BlockwiseSpec spec, int scale_mode, const float* __restrict__ eps,
                                        int64_t ld_e, unsigned long long seed, float* __restrict__ out, int64_t ld_o) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float* pr = params + b * ld_p;
  const float sc = apply_scale(pr[spec.scale[d]], scale_mode);
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  if (spec.kind[d] == VMS_DIST_NORMAL) {
    float e;
    if (eps) {
      e = eps[b * ld_e + d];
    } else {
      const uint4 r = philox4x32_10(make_uint4((unsigned)b, (unsigned)(b >> 32), (unsigned)d, 0u), key);
      float sn, cs;
      sincospif(2.f * u01(r.y), &sn, &cs);
      e = sqrtf(-2.f * logf(u01(r.x))) * cs;
    }
    out[b * ld_o + d] = fmaf(e, sc, pr[spec.loc[d]]);
    return;
  }
  const float loc = spec.loc2[d] >= 0 ? atan2f(pr[spec.loc[d]], pr[spec.loc2[d]]) : pr[spec.loc[d]];
  const float k = sc;
  const float r = 1.f + sqrtf(1.f + 4.f * k * k);
  const float rho = (r - sqrtf(2.f * r)) / (2.f * k);
  const float s = k > 2e-2f ? (1.f + rho * rho) / (2.f * rho) : 1.f / k;
  float w = 0.f, u = 0.f;
  for (unsigned attempt = 1; attempt <= 64; ++attempt) {  // acceptance is >= 0.66 for every concentration: 64 is never reached
    const uint4 rnd = philox4x32_10(make_uint4((unsigned)b, (unsigned)(b >> 32), (unsigned)d, attempt), key);
    u = 2.f * u01(rnd.x) - 1.f;
    const float z = cospif(u);
    w = (1.f + s * z) / (s + z);
    const float y = k * (s - w);
    const float v = u01(rnd.y);
    if (y * (2.f - y) >= v || logf(y / v) + 1.f >= y) break;
  }
  w = fminf(fmaxf(w, -1.f), 1.f);
  float x = copysignf(acosf(w), u) + loc;
  x -= 6.283185307179586f * rintf(x * 0.15915494309189535f);
  out[b * ld_o + d] = x;
}

// four standard normals per Philox call; element i of the stream comes from counter (offset + i) / 4, lane (offset + i) % 4
__global__ void std_normal_fill_kernel(unsigned long long seed, unsigned long long offset, int64_t n, float* __restrict__ out) {
  const unsigned long long first = offset >> 2, last = (offset + (unsigned long long)n + 3ull) >> 2;
  const unsigned long long q = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= last) return;
  const uint4 r = philox4x32_10(make_uint4((unsigned)q, (unsigned)(q >> 32), 0x4e6f726du, 0u),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  float v[4], sn, cs;
  float rad = sqrtf(-2.f * logf(u01(r.x)));
  sincospif(2.f * u01(r.y), &sn, &cs);
  v[0] = rad * cs; v[1] = rad * sn;
  rad = sqrtf(-2.f * logf(u01(r.z)));
  sincospif(2.f * u01(r.w), &sn, &cs);
  v[2] = rad * cs; v[3] = rad * sn;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned long long e = 4ull * q + k;
    if (e >= offset && e < offset + (unsigned long long)n) out[e - offset] = v[k];
  }
}

// Independent(Deterministic(loc)).log_prob (dists.py:701-704): 0 where every coordinate equals loc, -inf elsewhere
__global__ void deterministic_lp_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ loc,
                                        int64_t ld_l, int64_t B, int D, float* __restrict__ lp) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  bool eq = true;
  for (int d = 0; d < D; ++d) eq = eq && (x[b * ld_x + d] == loc[b * ld_l + d]);
  lp[b] = eq ? 0.f : -INFINITY;
}

}  // namespace vms

using namespace vms;

extern "C" {

static vms_status make_spec(BlockwiseSpec& spec, int64_t B, int D, const int32_t* kind, const int32_t* loc_off,
                            const int32_t* loc2_off, const int32_t* scale_off, int scale_mode) {
  VMS_REQUIRE(B >= 0 && D >= 1 && D <= kMaxDof, VMS_ERR_SHAPE, "blockwise: D must be in [1, %d]", kMaxDof);
  VMS_REQUIRE(kind && loc_off && scale_off, VMS_ERR_INVALID_ARG, "blockwise: NULL offset table");
  VMS_REQUIRE(scale_mode >= 0 && scale_mode <= 2, VMS_ERR_INVALID_ARG, "blockwise: bad scale_mode");
  for (int d = 0; d < D; ++d) {
    VMS_REQUIRE(kind[d] == VMS_DIST_NORMAL || kind[d] == VMS_DIST_VONMISES, VMS_ERR_UNSUPPORTED,
                "blockwise: unsupported distribution kind %d", kind[d]);
    spec.kind[d] = (int8_t)kind[d];
    spec.loc[d] = (int16_t)loc_off[d];
    spec.loc2[d] = (int16_t)(loc2_off ? loc2_off[d] : -1);
    spec.scale[d] = (int16_t)scale_off[d];
  }
  return VMS_OK;
}

vms_status vms_blockwise_params(const float* params, int64_t ld_p, int64_t B, int D, const int32_t* kind,
                                const int32_t* loc_off, const int32_t* loc2_off, const int32_t* scale_off,
                                int scale_mode, float* loc, float* scale, vms_stream stream) {
  BlockwiseSpec spec = {};
  vms_status s = make_spec(spec, B, D, kind, loc_off, loc2_off, scale_off, scale_mode);
  if (s) return s;
  VMS_REQUIRE(params && loc && scale, VMS_ERR_INVALID_ARG, "blockwise_params: NULL pointer");
  if (B == 0) return VMS_OK;
  blockwise_params_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, as_stream(stream)>>>(params, ld_p, B, D, spec,
                                                                                         scale_mode, loc, scale);
  VMS_LAUNCH_CHECK("blockwise_params_kernel");
  return VMS_OK;
}

vms_status vms_blockwise_log_prob(const float* x, int64_t ld_x, const float* params, int64_t ld_p, int64_t B, int D,
                                  const int32_t* kind, const int32_t* loc_off, const int32_t* loc2_off,
                                  const int32_t* scale_off, int scale_mode, float* lp, int accumulate,
                                  vms_stream stream) {
  BlockwiseSpec spec = {};
  vms_status s = make_spec(spec, B, D, kind, loc_off, loc2_off, scale_off, scale_mode);
  if (s) return s;
  VMS_REQUIRE(x && params && lp, VMS_ERR_INVALID_ARG, "blockwise_log_prob: NULL pointer");
  if (B == 0) return VMS_OK;
  bool planar = ld_x == D && ld_p == 2 * D && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                (reinterpret_cast<uintptr_t>(params) & 15u) == 0;
  for (int d = 0; d < D && planar; ++d) planar = kind[d] == VMS_DIST_NORMAL && loc_off[d] == d && scale_off[d] == D + d;
  if (planar && D <= 8) {
    const unsigned grid = (unsigned)((B + 127) / 128);
    cudaStream_t st = as_stream(stream);
    switch (D) {
#define VMS_NR(DD) case DD: normal_rows_lp_kernel<DD><<<grid, 128, 0, st>>>(x, params, B, scale_mode, lp, accumulate); break;
      VMS_NR(1) VMS_NR(2) VMS_NR(3) VMS_NR(4) VMS_NR(5) VMS_NR(6) VMS_NR(7) VMS_NR(8)
#undef VMS_NR
    }
    VMS_LAUNCH_CHECK("normal_rows_lp_kernel");
    return VMS_OK;
  }
  blockwise_lp_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, ld_x, params, ld_p, B, D, spec,
                                                                                 scale_mode, lp, accumulate);
  VMS_LAUNCH_CHECK("blockwise_lp_kernel");
  return VMS_OK;
}

vms_status vms_blockwise_sample(const float* params, int64_t ld_p, int64_t B, int D, const int32_t* kind,
                                const int32_t* loc_off, const int32_t* loc2_off, const int32_t* scale_off, int scale_mode,
                                const float* eps, int64_t ld_eps, unsigned long long seed, float* out, int64_t ld_out,
                                vms_stream stream) {
  BlockwiseSpec spec = {};
  vms_status s = make_spec(spec, B, D, kind, loc_off, loc2_off, scale_off, scale_mode);
  if (s) return s;
  VMS_REQUIRE(params && out, VMS_ERR_INVALID_ARG, "blockwise_sample: NULL pointer");
  if (B == 0) return VMS_OK;
  blockwise_sample_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, as_stream(stream)>>>(params, ld_p, B, D, spec,
                                                                                         scale_mode, eps, ld_eps, seed,
                                                                                         out, ld_out);
  VMS_LAUNCH_CHECK("blockwise_sample_kernel");
  return VMS_OK;
}

vms_status vms_standard_normal(unsigned long long seed, unsigned long long offset, int64_t n, float* out, vms_stream stream) {
  VMS_REQUIRE(out && n >= 0, VMS_ERR_INVALID_ARG, "standard_normal: bad arguments");
  if (n == 0) return VMS_OK;
  const unsigned long long quads = ((offset + (unsigned long long)n + 3ull) >> 2) - (offset >> 2);
  std_normal_fill_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, as_stream(stream)>>>(seed, offset, n, out);
  VMS_LAUNCH_CHECK("std_normal_fill_kernel");
  return VMS_OK;
}

vms_status vms_deterministic_log_prob(const float* x, int64_t ld_x, const float* loc, int64_t ld_loc, int64_t B, int D,
                                      float* lp, vms_stream stream) {
  VMS_REQUIRE(x && loc && lp && B >= 0 && D >= 1, VMS_ERR_INVALID_ARG, "deterministic_log_prob: bad arguments");
  if (B == 0) return VMS_OK;
  deterministic_lp_kernel<<<(unsigned)((B + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld_x, loc, ld_loc, B, D, lp);
  VMS_LAUNCH_CHECK("deterministic_lp_kernel");
  return VMS_OK;
}

vms_status vms_std_normal_log_prob(const float* x, int64_t ld_x, int64_t B, int D, float* lp, int accumulate,
                                   vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "std_normal_log_prob: bad shape");
  VMS_REQUIRE(x && lp, VMS_ERR_INVALID_ARG, "std_normal_log_prob: NULL pointer");
  if (B == 0) return VMS_OK;
  std_normal_lp_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, ld_x, B, D, lp, accumulate);
  VMS_LAUNCH_CHECK("std_normal_lp_kernel");
  return VMS_OK;
}

vms_status vms_normal_sample_log_prob(const float* params, int64_t ld_p, int loc_off, int scale_off, int scale_mode,
                                      const float* eps, int64_t B, int D, float* z, int64_t ld_z, float* lp,
                                      vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "normal_sample_log_prob: bad shape");
  VMS_REQUIRE(params && eps, VMS_ERR_INVALID_ARG, "normal_sample_log_prob: NULL pointer");
  if (B == 0) return VMS_OK;
  normal_sample_lp_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(
      params, ld_p, loc_off, scale_off, scale_mode, eps, B, D, z, ld_z, lp);
  VMS_LAUNCH_CHECK("normal_sample_lp_kernel");
  return VMS_OK;
}

vms_status vms_normal_log_prob_backward(const float* x, int64_t ld_x, const float* params, int64_t ld_p, int loc_off,
                                        int scale_off, int scale_mode, const float* g_lp, int64_t B, int D,
                                        float* g_x, int64_t ld_gx, int accumulate_x, float* g_params, int64_t ld_gp,
                                        vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "normal_log_prob_backward: bad shape");
  VMS_REQUIRE(x && params && g_lp, VMS_ERR_INVALID_ARG, "normal_log_prob_backward: NULL pointer");
  if (B == 0) return VMS_OK;
  normal_lp_bwd_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(
      x, ld_x, params, ld_p, loc_off, scale_off, scale_mode, g_lp, B, D, g_x, ld_gx, accumulate_x, g_params, ld_gp);
  VMS_LAUNCH_CHECK("normal_lp_bwd_kernel");
  return VMS_OK;
}

}  // extern "C"
