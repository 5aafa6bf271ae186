// mc_fused.cu -- VAE-proposal Monte Carlo: n_steps full MC steps of B independent chains in ONE kernel launch.
//
// Replaces, for the Gaussian-VAE family of tests/test_mcmc.py:14-26 (C4a of SURVEY 8d: FCDeepNN encoder / decoder
// with tfp.layers.IndependentNormal heads, N(0, I) prior), the whole body of `MCMC.single_step` (mcmc.py:68-130) and
// the loop of `MCMC.run` (mcmc.py:133-159):
//   :100-103  z1, log q(z1|x1) = encoder(x1).experimental_sample_and_log_prob();  z2, log p(z2) = prior sample;
//             x2, log p(x2|z2) = decoder(z2) sample;        forward_log_p = (l1 + l2) + l3          (float32)
//   :106-109  reverse_log_p = (log q(z2|x2) + log p(z1)) + log p(x1|z1)                              (float32)
//   :113      new energies = energy_func(x2)        (device-resident quadratic energy of tests/test_mcmc.py:28-32, float64)
//   :116-120  log_acc = E_new + rev - E_old - fwd (float64, left to right);  acc = log_acc >= log(u)
//   :123-128  counters; rejected chains keep their configuration and energy
// The op-by-op host path (vaemolsim_b200/mcmc.py over the per-layer kernels) stays as the generic path -- any VAE,
// any energy callback -- and as the cross-check of this kernel.
//
// Design (B200).  Chains are independent (mcmc.py:86-88): a CTA owns tiles of 32 chains and runs ALL steps of a tile
// before moving on, with the chain state, the 5,216 encoder / decoder weights and every activation in shared memory:
// per step a chain moves 8 bytes from HBM (its log u; the uniform stream stays NumPy's PCG64 on the host so decisions
// are bit-identical to the reference under the same seed) and nothing else.  One step is four small MLP evaluations
// arranged as two independent chains [enc(x1) -> z1 -> dec(z1)] and [z2 -> dec(z2) -> x2 -> enc(x2)], evaluated in
// lock-step with the tile GEMM routines of tile_gemm.cuh.  Sampling noise is Philox4x32-10 +
// Box-Muller keyed by (seed, global chain, step), so results do not depend on the grid or the number of GPUs; in
// parity mode the noise is an input instead.
#include "tile_gemm.cuh"
#include <math.h>
#include <string.h>

namespace vms {

struct McParams {
  long long chain0;  // global index of chain 0 (Philox key): results do not depend on how chains are sharded
  int dx, dz, hidden, n_mlp;
  int enc0W, enc0b, enc1W, enc1b, dec0W, dec0b, dec1W, dec1b;
  int64_t B;
  int n_tiles, n_steps;
  const float* theta;
  float* x;            // [B, dx] in / out
  double* E;           // [B] in / out
  int energies_valid;
  const float* noise;  // [n_steps, B, 2 dz + dx] or NULL (device RNG)
  unsigned long long seed, step0;
  const double* log_u; // [n_steps, B]
  const double* means; // [dx]
  unsigned long long* n_acc;
  uint8_t* acc_trace;  // [n_steps, B] or NULL
  float *fwd_trace, *rev_trace;  // [n_steps, B] or NULL
  double* e_new_trace;           // [n_steps, B] or NULL
  // shared-memory pitches / offsets (floats)
  int ldx, ldz, ldh, ldpe, ldpd, ldn;
  int o_Wp, o_x1, o_x1T, o_x2, o_x2T, o_z1, o_z1T, o_z2, o_z2T, o_he1, o_he2, o_hd1, o_hd2, o_pe1, o_pe2, o_pd1, o_pd2,
      o_nz, o_lp, o_E, o_tq, o_tx, o_te, o_sc, o_sc2;
};

// Philox4x32-10 (Salmon et al. 2011): counter-based, so chain / step / slot index the stream directly.
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
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
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, float& n0, float& n1) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0, 1)
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.f * logf(u1));
  float s, c;
  sincospif(2.f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

__global__ void __launch_bounds__(FT, 1) mc_fused_kernel(const __grid_constant__ McParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int dx = p.dx, dz = p.dz, H = p.hidden, ldh = p.ldh;
  const int nn = 2 * dz + dx;  // noise values per chain-step: eps(z1) | eps(z2) | eps(x2)
  float* x1 = sm + p.o_x1;
  float* x2 = sm + p.o_x2;
  float* z1 = sm + p.o_z1;
  float* z2 = sm + p.o_z2;
  float* nz = sm + p.o_nz;
  float* lp = sm + p.o_lp;   // [6][FR]: lq1, lz1, lz2, lx2, lx1, lq2
  double* Es = reinterpret_cast<double*>(sm + p.o_E);  // [2][FR]: E_old, E_new
  float* tq = sm + p.o_tq;   // [3][FR][dz] per-dof log-prob terms (latent side)
  float* tx = sm + p.o_tx;   // [FR][dx] per-dof log-prob terms (configuration side)
  double* te = reinterpret_cast<double*>(sm + p.o_te);  // [FR][dx] per-dof energy terms
  for (int i = tid; i < p.n_mlp; i += FT) cp_async4(sm + p.o_Wp + i, p.theta + i);
  cp_async_commit_wait_all();
  unsigned long long cta_acc = 0;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * FR;
    const int nr = (int)min((int64_t)FR, p.B - row0);
    __syncthreads();
    // chain state of the tile: configurations (row-major + feature-major) and energies
    for (int i = tid; i < FR * dx; i += FT) {
      const int r = i / dx, c = i - r * dx;
      const float v = r < nr ? p.x[(row0 + r) * dx + c] : 0.f;
      x1[r * p.ldx + c] = v;
      sm[p.o_x1T + c * FR + r] = v;
    }
    __syncthreads();
    if (tid < FR) {
      double e = 0.0;
      if (tid < nr) {
        if (p.energies_valid) {
          e = p.E[row0 + tid];
        } else {
          // tests/test_mcmc.py:28-32: sum((x - means)^2) with float64 means => float64 arithmetic
          for (int d = 0; d < dx; ++d) {
            const double t = __dsub_rn((double)x1[tid * p.ldx + d], p.means[d]);
            e = __dadd_rn(e, __dmul_rn(t, t));  // NumPy squares, then sums: no FMA contraction
          }
        }
      }
      Es[tid] = e;
    }
#pragma unroll 1
    for (int step = 0; step < p.n_steps; ++step) {
      // ---------------------------------------------------------------- noise of this step
      if (p.noise) {
        for (int i = tid; i < FR * nn; i += FT) {
          const int r = i / nn, c = i - r * nn;
          nz[r * p.ldn + c] = r < nr ? __ldg(p.noise + ((int64_t)step * p.B + row0 + r) * nn + c) : 0.f;
        }
      } else {
        const int quads = (nn + 3) / 4;
        for (int i = tid; i < FR * quads; i += FT) {
          const int r = i / quads, q = i - r * quads;
          const unsigned long long chain = (unsigned long long)(p.chain0 + row0 + r), st = p.step0 + (unsigned long long)step;
          const uint4 rnd = philox4x32(make_uint4((unsigned)chain, (unsigned)(chain >> 32), (unsigned)st, (unsigned)q),
                                       make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32) ^ (unsigned)(st >> 32)));
          float n[4];
          box_muller(rnd.x, rnd.y, n[0], n[1]);
          box_muller(rnd.z, rnd.w, n[2], n[3]);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (4 * q + k < nn) nz[r * p.ldn + 4 * q + k] = n[k];
        }
      }
      __syncthreads();
      // z2 ~ N(0, I) prior sample (tfp Normal._sample_n with loc 0, scale 1), both layouts
      for (int i = tid; i < FR * dz; i += FT) {
        const int r = i / dz, d = i - r * dz;
        const float v = nz[r * p.ldn + dz + d];
        z2[r * p.ldz + d] = v;
        sm[p.o_z2T + d * FR + r] = v;
      }
      __syncthreads();
      // ---------------------------------------------------------------- A: enc hidden(x1)  ||  dec hidden(z2)
      outer_gemm<8, 2, false>(p.o_x1T, FR, p.o_Wp + p.enc0W, H, 1, FR, H, dx, 0, 0,
                              epi_store(p.o_he1, ldh, p.o_Wp + p.enc0b, 1));
      outer_gemm<8, 2, false>(p.o_z2T, FR, p.o_Wp + p.dec0W, H, 1, FR, H, dz, 0, 0,
                              epi_store(p.o_hd2, ldh, p.o_Wp + p.dec0b, 1));
      __syncthreads();
      // ---------------------------------------------------------------- B: enc head(x1)  ||  dec head(z2)
      thin_gemm(p.o_he1, ldh, p.o_Wp + p.enc1W, 2 * dz, 1, H, 2 * dz, p.o_pe1, p.ldpe, p.o_Wp + p.enc1b, 0, p.o_sc);
      thin_gemm(p.o_hd2, ldh, p.o_Wp + p.dec1W, 2 * dx, 1, H, 2 * dx, p.o_pd2, p.ldpd, p.o_Wp + p.dec1b, 0, p.o_sc2);
      __syncthreads();
      // ---------------------------------------------------------------- C: samples z1, x2 and their log-prob terms
      // one thread per (chain, dof): the per-dof terms go to shared memory and are summed in dof order afterwards
      // (same summation order as the op-by-op kernels; a single warp doing all dofs was 8k cycles of dependent math)
      for (int i = tid; i < FR * (dz + dx); i += FT) {
        if (i < FR * dz) {
          const int r = i / dz, d = i - r * dz;
          const float* pe = sm + p.o_pe1 + r * p.ldpe;
          const float loc = pe[d], sc = softplus_tf(pe[dz + d]);
          const float zz = __fadd_rn(__fmul_rn(nz[r * p.ldn + d], sc), loc);  // separate TF mul and add ops
          z1[r * p.ldz + d] = zz;
          sm[p.o_z1T + d * FR + r] = zz;
          tq[r * dz + d] = normal_lp(zz, loc, sc);
          tq[(FR + r) * dz + d] = normal_lp(zz, 0.f, 1.f);
          tq[(2 * FR + r) * dz + d] = normal_lp(z2[r * p.ldz + d], 0.f, 1.f);
        } else {
          const int j = i - FR * dz, r = j / dx, d = j - r * dx;
          const float* pd = sm + p.o_pd2 + r * p.ldpd;
          const float loc = pd[d], sc = softplus_tf(pd[dx + d]);
          const float xx = __fadd_rn(__fmul_rn(nz[r * p.ldn + 2 * dz + d], sc), loc);
          x2[r * p.ldx + d] = xx;
          sm[p.o_x2T + d * FR + r] = xx;
          tx[r * dx + d] = normal_lp(xx, loc, sc);
          const double t = __dsub_rn((double)xx, p.means[d]);
          te[r * dx + d] = __dmul_rn(t, t);
        }
      }
      __syncthreads();
      if (tid < FR) {
        const int r = tid;
        float lq1 = 0.f, lz1 = 0.f, lz2 = 0.f, lx2 = 0.f;
        for (int d = 0; d < dz; ++d) {
          lq1 += tq[r * dz + d];
          lz1 += tq[(FR + r) * dz + d];
          lz2 += tq[(2 * FR + r) * dz + d];
        }
        double e_new = 0.0;
        for (int d = 0; d < dx; ++d) {
          lx2 += tx[r * dx + d];
          e_new = __dadd_rn(e_new, te[r * dx + d]);  // NumPy squares, then sums: no FMA contraction
        }
        lp[r] = lq1; lp[FR + r] = lz1; lp[2 * FR + r] = lz2; lp[3 * FR + r] = lx2;
        Es[FR + r] = e_new;
      }
      // ---------------------------------------------------------------- D: dec hidden(z1)  ||  enc hidden(x2)
      outer_gemm<8, 2, false>(p.o_z1T, FR, p.o_Wp + p.dec0W, H, 1, FR, H, dz, 0, 0,
                              epi_store(p.o_hd1, ldh, p.o_Wp + p.dec0b, 1));
      outer_gemm<8, 2, false>(p.o_x2T, FR, p.o_Wp + p.enc0W, H, 1, FR, H, dx, 0, 0,
                              epi_store(p.o_he2, ldh, p.o_Wp + p.enc0b, 1));
      __syncthreads();
      // ---------------------------------------------------------------- E: dec head(z1)  ||  enc head(x2)
      thin_gemm(p.o_hd1, ldh, p.o_Wp + p.dec1W, 2 * dx, 1, H, 2 * dx, p.o_pd1, p.ldpd, p.o_Wp + p.dec1b, 0, p.o_sc2);
      thin_gemm(p.o_he2, ldh, p.o_Wp + p.enc1W, 2 * dz, 1, H, 2 * dz, p.o_pe2, p.ldpe, p.o_Wp + p.enc1b, 0, p.o_sc);
      __syncthreads();
      // ---------------------------------------------------------------- F: reverse log-probs, accept / reject
      for (int i = tid; i < FR * (dz + dx); i += FT) {
        if (i < FR * dz) {
          const int r = i / dz, d = i - r * dz;
          const float* pe = sm + p.o_pe2 + r * p.ldpe;
          tq[r * dz + d] = normal_lp(z2[r * p.ldz + d], pe[d], softplus_tf(pe[dz + d]));
        } else {
          const int j = i - FR * dz, r = j / dx, d = j - r * dx;
          const float* pd = sm + p.o_pd1 + r * p.ldpd;
          tx[r * dx + d] = normal_lp(x1[r * p.ldx + d], pd[d], softplus_tf(pd[dx + d]));
        }
      }
      __syncthreads();
      if (tid < FR) {
        const int r = tid;
        float lq2 = 0.f, lx1 = 0.f;
        for (int d = 0; d < dz; ++d) lq2 += tq[r * dz + d];
        for (int d = 0; d < dx; ++d) lx1 += tx[r * dx + d];
        // mcmc.py:103, :109: float32 sums, left to right
        const float fwd = __fadd_rn(__fadd_rn(lp[r], lp[2 * FR + r]), lp[3 * FR + r]);
        const float rev = __fadd_rn(__fadd_rn(lq2, lp[FR + r]), lx1);
        bool a = false;
        if (r < nr) {
          const int64_t g = (int64_t)step * p.B + row0 + r;
          const double e_old = Es[r], e_new = Es[FR + r];
          // mcmc.py:116: ((E_new + rev) - E_old) - fwd in float64, no contraction possible
          const double la = __dsub_rn(__dsub_rn(__dadd_rn(e_new, (double)rev), e_old), (double)fwd);
          a = la >= __ldg(p.log_u + g);
          if (p.acc_trace) p.acc_trace[g] = a ? 1 : 0;
          if (p.fwd_trace) p.fwd_trace[g] = fwd;
          if (p.rev_trace) p.rev_trace[g] = rev;
          if (p.e_new_trace) p.e_new_trace[g] = e_new;
          if (a) {
            Es[r] = e_new;
            for (int d = 0; d < dx; ++d) {
              const float v = x2[r * p.ldx + d];
              x1[r * p.ldx + d] = v;
              sm[p.o_x1T + d * FR + r] = v;
            }
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, a);
        if (r == 0) cta_acc += (unsigned long long)__popc(m);
      }
      __syncthreads();
    }
    // write the tile's chain state back
    for (int i = tid; i < nr * dx; i += FT) {
      const int r = i / dx, c = i - r * dx;
      p.x[(row0 + r) * dx + c] = x1[r * p.ldx + c];
    }
    if (tid < nr) p.E[row0 + tid] = Es[tid];
  }
  if (tid == 0 && cta_acc) atomicAdd(p.n_acc, cta_acc);
}

}  // namespace vms

namespace vms {
// mc_chain.cu: four-lanes-per-chain formulation, the default for the C4a shape (VMS_MC_KERNEL=tile forces this file's kernel)
bool mc_chain_enabled(int dx, int dz);
vms_status mc_chain_run(int dx, int dz, int hidden, const float* theta, float* x, double* E, int energies_valid,
                        const float* noise, unsigned long long seed, unsigned long long step0, const double* log_u,
                        const double* means, int64_t B, int n_steps, unsigned long long* n_acc, uint8_t* acc_trace,
                        float* fwd_trace, float* rev_trace, double* e_new_trace, cudaStream_t st,
                        const vms_pcg64_stream* rng, unsigned long long* n_uncertain, double* log_u_trace, long long chain0);
}  // namespace vms

using namespace vms;

struct vms_mc_plan_s {
  long long chain0 = 0;
  vms_mc_desc d;
  McParams p;
  size_t smem_bytes;
  int max_grid;
};

static inline int r4i(int v) { return (v + 3) & ~3; }

extern "C" {

int64_t vms_mc_param_count(const vms_mc_desc* d) {
  if (!d) return -1;
  return (int64_t)d->dx * d->hidden + d->hidden + (int64_t)d->hidden * 2 * d->dz + 2 * d->dz + (int64_t)d->dz * d->hidden +
         d->hidden + (int64_t)d->hidden * 2 * d->dx + 2 * d->dx;
}

vms_status vms_mc_plan_create(const vms_mc_desc* desc, vms_mc_plan* plan) {
  VMS_REQUIRE(desc && plan, VMS_ERR_INVALID_ARG, "mc_plan_create: NULL argument");
  const vms_mc_desc& d = *desc;
  VMS_REQUIRE(d.dx >= 1 && d.dz >= 1 && d.hidden >= 1, VMS_ERR_SHAPE, "mc_plan_create: dx, dz, hidden must be >= 1");
  VMS_REQUIRE(d.dx <= kMaxThin / 2 && d.dz <= kMaxThin / 2, VMS_ERR_UNSUPPORTED,
              "mc_plan_create: the fused MC kernel is built for dx, dz <= %d; use the op-by-op MCMC path", kMaxThin / 2);
  vms_mc_plan_s* pl = new vms_mc_plan_s();
  pl->d = d;
  McParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.dx = d.dx; p.dz = d.dz; p.hidden = d.hidden;
  int o = 0;
  p.enc0W = o; o += d.dx * d.hidden;
  p.enc0b = o; o += d.hidden;
  p.enc1W = o; o += d.hidden * 2 * d.dz;
  p.enc1b = o; o += 2 * d.dz;
  p.dec0W = o; o += d.dz * d.hidden;
  p.dec0b = o; o += d.hidden;
  p.dec1W = o; o += d.hidden * 2 * d.dx;
  p.dec1b = o; o += 2 * d.dx;
  p.n_mlp = o;
  p.ldx = r4i(d.dx); p.ldz = r4i(d.dz); p.ldh = (d.hidden + 1) | 1; p.ldpe = r4i(2 * d.dz); p.ldpd = r4i(2 * d.dx);
  p.ldn = r4i(2 * d.dz + d.dx);
  int off = 0;
  auto take = [&](int n) { int o0 = off; off += r4i(n); return o0; };
  p.o_Wp = take(p.n_mlp);
  p.o_x1 = take(FR * p.ldx); p.o_x1T = take(d.dx * FR); p.o_x2 = take(FR * p.ldx); p.o_x2T = take(d.dx * FR);
  p.o_z1 = take(FR * p.ldz); p.o_z1T = take(d.dz * FR); p.o_z2 = take(FR * p.ldz); p.o_z2T = take(d.dz * FR);
  p.o_he1 = take(FR * p.ldh); p.o_he2 = take(FR * p.ldh); p.o_hd1 = take(FR * p.ldh); p.o_hd2 = take(FR * p.ldh);
  p.o_pe1 = take(FR * p.ldpe); p.o_pe2 = take(FR * p.ldpe); p.o_pd1 = take(FR * p.ldpd); p.o_pd2 = take(FR * p.ldpd);
  p.o_nz = take(FR * p.ldn); p.o_lp = take(6 * FR);
  p.o_E = take(4 * FR);  // 2 x FR doubles (offset is a multiple of 4 floats => 16-byte aligned)
  p.o_sc = take(FW * r4i(2 * d.dz) * FR); p.o_sc2 = take(FW * r4i(2 * d.dx) * FR);
  p.o_tq = take(3 * FR * d.dz); p.o_tx = take(FR * d.dx); p.o_te = take(2 * FR * d.dx);
  off += 64;
  pl->smem_bytes = (size_t)off * sizeof(float);
  if (pl->smem_bytes > (size_t)max_smem_optin()) {
    delete pl;
    set_error("mc_plan_create: hidden = %d needs %zu bytes of shared memory per CTA; use the op-by-op MCMC path",
              d.hidden, (size_t)off * sizeof(float));
    return VMS_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(mc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_bytes);
  if (e != cudaSuccess) {
    delete pl;
    set_error("mc_plan_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return VMS_ERR_CUDA;
  }
  pl->max_grid = sm_count();
  *plan = pl;
  return VMS_OK;
}

vms_status vms_mc_plan_destroy(vms_mc_plan plan) {
  delete plan;
  return VMS_OK;
}

vms_status vms_mc_run(vms_mc_plan pl, const float* theta, float* x, double* E, int energies_valid, const float* noise,
                      unsigned long long seed, unsigned long long step0, const double* log_u, const double* means,
                      int64_t B, int n_steps, unsigned long long* n_acc, uint8_t* acc_trace, float* fwd_trace,
                      float* rev_trace, double* e_new_trace, vms_stream stream) {
  VMS_RANGE("vms_mc_run");
  VMS_REQUIRE(pl, VMS_ERR_INVALID_ARG, "mc_run: NULL plan");
  VMS_REQUIRE(theta && x && E && log_u && means && n_acc, VMS_ERR_INVALID_ARG, "mc_run: NULL pointer");
  VMS_REQUIRE(B >= 0 && n_steps >= 0, VMS_ERR_SHAPE, "mc_run: negative size");
  if (B == 0 || n_steps == 0) return VMS_OK;
  if (mc_chain_enabled(pl->d.dx, pl->d.dz))
    return mc_chain_run(pl->d.dx, pl->d.dz, pl->d.hidden, theta, x, E, energies_valid, noise, seed, step0, log_u, means, B,
                        n_steps, n_acc, acc_trace, fwd_trace, rev_trace, e_new_trace, as_stream(stream), nullptr, nullptr,
                        nullptr, pl->chain0);
  McParams p = pl->p;
  p.chain0 = pl->chain0;
  p.B = B;
  p.n_tiles = (int)((B + FR - 1) / FR);
  p.n_steps = n_steps;
  p.theta = theta; p.x = x; p.E = E; p.energies_valid = energies_valid;
  p.noise = noise; p.seed = seed; p.step0 = step0; p.log_u = log_u; p.means = means;
  p.n_acc = n_acc; p.acc_trace = acc_trace; p.fwd_trace = fwd_trace; p.rev_trace = rev_trace; p.e_new_trace = e_new_trace;
  const int grid = p.n_tiles < pl->max_grid ? p.n_tiles : pl->max_grid;
  mc_fused_kernel<<<grid, FT, pl->smem_bytes, as_stream(stream)>>>(p);
  VMS_LAUNCH_CHECK("mc_fused_kernel");
  return VMS_OK;
}

vms_status vms_mc_plan_set_chain_offset(vms_mc_plan pl, int64_t chain0) {
  VMS_REQUIRE(pl && chain0 >= 0, VMS_ERR_INVALID_ARG, "mc_plan_set_chain_offset: bad arguments");
  pl->chain0 = chain0;
  return VMS_OK;
}

int vms_mc_plan_has_device_rng(vms_mc_plan pl) { return pl && mc_chain_enabled(pl->d.dx, pl->d.dz) ? 1 : 0; }

vms_status vms_mc_run_pcg64(vms_mc_plan pl, const float* theta, float* x, double* E, int energies_valid, const float* noise,
                            unsigned long long seed, unsigned long long step0, const vms_pcg64_stream* rng,
                            const double* means, int64_t B, int n_steps, unsigned long long* n_acc,
                            unsigned long long* n_uncertain, uint8_t* acc_trace, float* fwd_trace, float* rev_trace,
                            double* e_new_trace, double* log_u_trace, vms_stream stream) {
  VMS_RANGE("vms_mc_run_pcg64");
  VMS_REQUIRE(pl, VMS_ERR_INVALID_ARG, "mc_run_pcg64: NULL plan");
  VMS_REQUIRE(theta && x && E && rng && means && n_acc && n_uncertain, VMS_ERR_INVALID_ARG, "mc_run_pcg64: NULL pointer");
  VMS_REQUIRE(B >= 0 && n_steps >= 0 && rng->chain0 >= 0, VMS_ERR_SHAPE, "mc_run_pcg64: negative size");
  VMS_REQUIRE(mc_chain_enabled(pl->d.dx, pl->d.dz), VMS_ERR_UNSUPPORTED,
              "mc_run_pcg64: the device uniform stream is built into the chain kernel (dx = 6, dz = 2); use vms_mc_run");
  if (B == 0 || n_steps == 0) return VMS_OK;
  return mc_chain_run(pl->d.dx, pl->d.dz, pl->d.hidden, theta, x, E, energies_valid, noise, seed, step0, nullptr, means, B,
                      n_steps, n_acc, acc_trace, fwd_trace, rev_trace, e_new_trace, as_stream(stream), rng, n_uncertain,
                      log_u_trace, rng->chain0);
}

}  // extern "C"
