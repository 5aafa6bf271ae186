// mcmc.cu -- K7: batched Metropolis acceptance for VAE-proposal MC moves, and device-resident test energies.
//
// Replaces mcmc.py:116-128 (NumPy on the host in the reference):
//   log_acc = new_energies + reverse_log_p - energies - forward_log_p      (float64; float32 log-probs promoted)
//   acc     = log_acc >= log_rand
//   counters, and restoring the old configuration / energy of rejected chains.
// The uniform stream stays on the host (NumPy PCG64 + glibc log, mcmc.py:119) so decisions are bit-exact; only
// log_u (8 bytes per chain) is uploaded.  Chains are independent: one thread per chain, no communication.
#include "common.cuh"

namespace vms {

// NumPy evaluates left to right, ((E_new + rev) - E_old) - fwd, in the promoted type of its operands: float64 when the energy
// callback returns float64 (tests/test_mcmc.py:28-32), float32 when it returns float32 like a tfp log_prob (MC notebook cell
// 38); only additions, so no contraction is possible.  The comparison with the float64 log_rand promotes a float32 log_acc.
__device__ __forceinline__ double log_acc(double en, double eo, float f, float r) {
  return __dsub_rn(__dsub_rn(__dadd_rn(en, (double)r), eo), (double)f);
}
__device__ __forceinline__ double log_acc(float en, float eo, float f, float r) {
  return (double)__fsub_rn(__fsub_rn(__fadd_rn(en, r), eo), f);
}

template <typename ET>
__global__ void mc_accept_kernel(const ET* __restrict__ E_new, const ET* __restrict__ E_old,
                                 const float* __restrict__ fwd, const float* __restrict__ rev,
                                 const double* __restrict__ log_u, int64_t B, int D, const float* __restrict__ x_old,
                                 float* __restrict__ x_new, ET* __restrict__ E_out, uint8_t* __restrict__ acc,
                                 unsigned long long* n_acc) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool a = false;
  if (b < B) {
    a = log_acc(E_new[b], E_old[b], fwd[b], rev[b]) >= log_u[b];
    if (!a)
      for (int d = 0; d < D; ++d) x_new[b * D + d] = x_old[b * D + d];
    E_out[b] = a ? E_new[b] : E_old[b];
    if (acc) acc[b] = a ? 1 : 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, a);
  if ((threadIdx.x & 31) == 0 && m && n_acc) atomicAdd(n_acc, (unsigned long long)__popc(m));
}

// tests/test_mcmc.py:28-32: np.sum((configs - means)**2, axis=-1) with float32 configs and float64 means
__global__ void energy_quadratic_kernel(const float* __restrict__ x, int64_t B, int D, const double* __restrict__ means,
                                        double* __restrict__ E) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    const double t = __dsub_rn((double)x[b * D + d], means[d]);
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  E[b] = s;
}

}  // namespace vms

using namespace vms;

namespace vms {
// MC_Moves_with_VAEs.ipynb cell 5 / 38: `data_dist.log_prob(configs)` of a tfp Mixture of Independent Normals, evaluated the
// way tfp does in float32: per component sum_d Normal log_prob (x / s - m / s form) + log cat prob, logsumexp over the
// components (max-shifted); float32 out like `log_prob(...).numpy()`.  One thread per configuration.
__global__ void energy_gmm_kernel(const float* __restrict__ x, int64_t B, int D, int n_comp, const float* __restrict__ log_w,
                                  const float* __restrict__ loc, const float* __restrict__ scale, float* __restrict__ E) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float lp[16];
  float mx = -INFINITY;
  for (int k = 0; k < n_comp; ++k) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) s += normal_lp(x[b * D + d], loc[k * D + d], scale[k * D + d]);
    lp[k] = s + log_w[k];
    mx = fmaxf(mx, lp[k]);
  }
  float acc = 0.f;
  for (int k = 0; k < n_comp; ++k) acc += expf(lp[k] - mx);
  E[b] = mx + logf(acc);
}
}  // namespace vms

extern "C" {

vms_status vms_mc_accept(const double* E_new, const double* E_old, const float* fwd, const float* rev,
                         const double* log_u, int64_t B, int D, const float* x_old, float* x_new_inout, double* E_out,
                         uint8_t* acc, unsigned long long* n_acc, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "mc_accept: bad shape");
  VMS_REQUIRE(B == 0 || (E_new && E_old && fwd && rev && log_u && x_old && x_new_inout && E_out), VMS_ERR_INVALID_ARG,
              "mc_accept: NULL pointer");
  if (B == 0) return VMS_OK;
  mc_accept_kernel<double><<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(E_new, E_old, fwd, rev, log_u, B, D,
                                                                                      x_old, x_new_inout, E_out, acc, n_acc);
  VMS_LAUNCH_CHECK("mc_accept_kernel");
  return VMS_OK;
}

vms_status vms_mc_accept_f32(const float* E_new, const float* E_old, const float* fwd, const float* rev,
                             const double* log_u, int64_t B, int D, const float* x_old, float* x_new_inout, float* E_out,
                             uint8_t* acc, unsigned long long* n_acc, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "mc_accept_f32: bad shape");
  VMS_REQUIRE(B == 0 || (E_new && E_old && fwd && rev && log_u && x_old && x_new_inout && E_out), VMS_ERR_INVALID_ARG,
              "mc_accept_f32: NULL pointer");
  if (B == 0) return VMS_OK;
  mc_accept_kernel<float><<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(E_new, E_old, fwd, rev, log_u, B, D,
                                                                                     x_old, x_new_inout, E_out, acc, n_acc);
  VMS_LAUNCH_CHECK("mc_accept_kernel<float>");
  return VMS_OK;
}

vms_status vms_energy_quadratic(const float* x, int64_t B, int D, const double* means, double* E, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "energy_quadratic: bad shape");
  VMS_REQUIRE(B == 0 || (x && means && E), VMS_ERR_INVALID_ARG, "energy_quadratic: NULL pointer");
  if (B == 0) return VMS_OK;
  energy_quadratic_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, B, D, means, E);
  VMS_LAUNCH_CHECK("energy_quadratic_kernel");
  return VMS_OK;
}

vms_status vms_energy_gmm(const float* x, int64_t B, int D, int n_comp, const float* log_w, const float* loc,
                          const float* scale, float* E, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1 && n_comp >= 1 && n_comp <= 16, VMS_ERR_SHAPE, "energy_gmm: need D >= 1 and 1..16 components");
  VMS_REQUIRE(B == 0 || (x && log_w && loc && scale && E), VMS_ERR_INVALID_ARG, "energy_gmm: NULL pointer");
  if (B == 0) return VMS_OK;
  energy_gmm_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, B, D, n_comp, log_w, loc, scale, E);
  VMS_LAUNCH_CHECK("energy_gmm_kernel");
  return VMS_OK;
}

}  // extern "C"
