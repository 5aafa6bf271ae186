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

__global__ void mc_accept_kernel(const double* __restrict__ E_new, const double* __restrict__ E_old,
                                 const float* __restrict__ fwd, const float* __restrict__ rev,
                                 const double* __restrict__ log_u, int64_t B, int D, const float* __restrict__ x_old,
                                 float* __restrict__ x_new, double* __restrict__ E_out, uint8_t* __restrict__ acc,
                                 unsigned long long* n_acc) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool a = false;
  if (b < B) {
    // NumPy evaluates left to right: ((E_new + rev) - E_old) - fwd, all in float64, no contraction possible
    const double la = __dsub_rn(__dsub_rn(__dadd_rn(E_new[b], (double)rev[b]), E_old[b]), (double)fwd[b]);
    a = la >= log_u[b];
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

extern "C" {

vms_status vms_mc_accept(const double* E_new, const double* E_old, const float* fwd, const float* rev,
                         const double* log_u, int64_t B, int D, const float* x_old, float* x_new_inout, double* E_out,
                         uint8_t* acc, unsigned long long* n_acc, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "mc_accept: bad shape");
  VMS_REQUIRE(B == 0 || (E_new && E_old && fwd && rev && log_u && x_old && x_new_inout && E_out), VMS_ERR_INVALID_ARG,
              "mc_accept: NULL pointer");
  if (B == 0) return VMS_OK;
  mc_accept_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(E_new, E_old, fwd, rev, log_u, B, D,
                                                                              x_old, x_new_inout, E_out, acc, n_acc);
  VMS_LAUNCH_CHECK("mc_accept_kernel");
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

}  // extern "C"
