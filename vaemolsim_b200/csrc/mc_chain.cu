// mc_chain.cu -- the production fused MC kernel for the C4a shape (dx = 6, dz = 2): a few lanes per chain, chain state in
// registers.  mc_fused_kernel (mc_fused.cu: 32-chain tiles, tile GEMMs) serves every other dx, dz <= 8 and stays
// selectable with VMS_MC_KERNEL=tile as an on-device cross-check.  Measured on a B200 (65,536 chains x 100 steps):
// 1,469 M proposals/s = 28.2 TFLOP/s with one lane per chain and two chains per lane (1,270 M with two lanes per chain,
// 799 M with four, 417 M for the tile kernel); bit-exact
// against the goldens of the reference's own mcmc.py (tests/test_gpu_models.py).
//
// Same contract as mc_fused.cu (mcmc.py:68-130 `MCMC.single_step` and the loop of `MCMC.run` :133-159 for the
// Gaussian-VAE family of tests/test_mcmc.py:14-26): n_steps VAE-proposal MC steps of B chains in one launch, noise from
// the same Philox4x32-10 stream (or given), accept uniforms from the host's PCG64 stream, float64 acceptance arithmetic
// in the reference's order.
//
// Why a second formulation.  mc_fused_kernel runs a 32-chain tile through tile-GEMM phases separated by CTA barriers
// (6 per step) and reaches 8 TFLOP/s (ncu: issue-active 55 %, barrier 1.7 stalled warps per issue).  The layers are
// thin (contractions of 6 and 2, outputs of 4 and 12): mlp_stream.cu's "thread per row, weights as 16-byte shared-memory
// broadcasts" forward reaches 19-21 TFLOP/s on the same shapes.  Here the lanes that own a chain keep its state in
// registers for all its steps and split the hidden units of the two MLP evaluations that run side by side (encoder(x1)
// with decoder(z2), then decoder(z1) with encoder(x2)) as FOUR interleaved unit streams whose partial head outputs meet as
// (s0 + s1) + (s2 + s3) -- in registers or through a butterfly, depending on the lane count (see `combine`).  No
// shared-memory activations, no CTA barrier inside a step; per step a chain reads its log u (8 bytes).  The four streams
// are the same for EVERY lane count and batch size: the summation order, hence every decision, is independent of the
// number of chains per GPU and of how many lanes the launcher gives a chain.
#include "common.cuh"
#include "mc_rng.cuh"
#include <math.h>
#include <stdlib.h>

namespace vms {

using namespace mcdev;

namespace {

constexpr int CT = 128;   // threads per CTA
constexpr int WROW = 28;  // floats per hidden unit in shared memory (7 x 16 bytes), see stage below
constexpr int kMaxDx = 6, kMaxDz = 2;  // compile-time register arrays (the C4a shape: dx = 6, dz = 2)

struct ChainParams {
  int dx, dz, hidden;
  int enc0W, enc0b, enc1W, enc1b, dec0W, dec0b, dec1W, dec1b;
  int64_t B;
  int n_steps;
  const float* theta;
  float* x;
  double* E;
  int energies_valid;
  const float* noise;
  unsigned long long seed, step0;
  const double* log_u;
  const double* means;
  unsigned long long* n_acc;
  uint8_t* acc_trace;
  float *fwd_trace, *rev_trace;
  double* e_new_trace;
  // device PCG64 uniform stream (log_u == NULL): see pcg64 below
  int use_pcg;
  unsigned long long s0_hi, s0_lo, inc_hi, inc_lo, jm_hi, jm_lo, ja_hi, ja_lo;
  int64_t chain0;
  unsigned long long* n_uncertain;
  double* log_u_trace;
};

// Encoder and decoder evaluated side by side on this lane's hidden units j = sub, sub + 4, ...:
//   pe [2 dz] += relu(xe W0e + b0e)_j W1e[j, :],   pd [2 dx] += relu(zd W0d + b0d)_j W1d[j, :]
// shared-memory row of hidden unit j (WROW floats): W0e[0..5][j] | b0e[j] | W1e[j][0..3] | W0d[0..1][j] | b0d[j] |
// W1d[j][0..11] | pad  (dx = 6, dz = 2)
// The hidden units are walked as FOUR interleaved streams (units j = s, s + 4, ...), each summed in ascending j, and the
// four partial sums meet as (s0 + s1) + (s2 + s3).  A chain is owned by TPC lanes (4, 2 or 1), each taking 4 / TPC streams:
// the summation order, hence every bit of every result and every decision, is the same for every TPC, so the launcher picks
// the lane count by the number of chains (one lane per chain when the chains alone fill the GPU: no replicated scalar work
// and a quarter of the shared-memory traffic; four lanes when chains are scarce, e.g. one shard of an 8-GPU job).
template <int TPC>
__device__ __forceinline__ float combine(const float (&a)[4 / TPC]) {
  if constexpr (TPC == 4) {
    return quad_sum(a[0]);
  } else if constexpr (TPC == 2) {
    const float t0 = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 1);
    const float t1 = a[1] + __shfl_xor_sync(0xffffffffu, a[1], 1);
    return t0 + t1;
  } else {
    return (a[0] + a[1]) + (a[2] + a[3]);
  }
}

// CPL chains per lane group share every weight load (the kernel is bound by shared-memory wavefronts -- a 16-byte
// broadcast costs up to four -- not by FFMA issue: two chains per load halve the wavefronts per FFMA).
template <int TPC, int CPL>
__device__ __forceinline__ void mlp_pair(const float* __restrict__ wsm, int Hp, int sub, const float (&xe)[CPL][kMaxDx],
                                         const float (&zd)[CPL][kMaxDz], float (&pe)[CPL][2 * kMaxDz],
                                         float (&pd)[CPL][2 * kMaxDx]) {
  constexpr int NS = 4 / TPC;
  float ae[CPL][2 * kMaxDz][NS], ad[CPL][2 * kMaxDx][NS];
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int q = 0; q < NS; ++q) {
#pragma unroll
      for (int n = 0; n < 2 * kMaxDz; ++n) ae[k][n][q] = 0.f;
#pragma unroll
      for (int n = 0; n < 2 * kMaxDx; ++n) ad[k][n][q] = 0.f;
    }
#pragma unroll 2
  for (int j0 = 0; j0 < Hp; j0 += 4) {  // Hp: hidden units padded with zero rows to a multiple of 4 (a zero row adds +0)
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const float4* row = reinterpret_cast<const float4*>(wsm + (j0 + sub + TPC * q) * WROW);
      const float4 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4], f = row[5], g = row[6];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        // encoder hidden unit: (sum_i x_i W0[i][j]) + b0[j]
        float he = 0.f;
        he = fmaf(xe[k][0], a.x, he); he = fmaf(xe[k][1], a.y, he); he = fmaf(xe[k][2], a.z, he);
        he = fmaf(xe[k][3], a.w, he); he = fmaf(xe[k][4], b.x, he); he = fmaf(xe[k][5], b.y, he);
        he = fmaxf(he + b.z, 0.f);
        ae[k][0][q] = fmaf(he, b.w, ae[k][0][q]); ae[k][1][q] = fmaf(he, c.x, ae[k][1][q]);
        ae[k][2][q] = fmaf(he, c.y, ae[k][2][q]); ae[k][3][q] = fmaf(he, c.z, ae[k][3][q]);
        // decoder hidden unit
        float hd = 0.f;
        hd = fmaf(zd[k][0], c.w, hd); hd = fmaf(zd[k][1], d.x, hd);
        hd = fmaxf(hd + d.y, 0.f);
        ad[k][0][q] = fmaf(hd, d.z, ad[k][0][q]); ad[k][1][q] = fmaf(hd, d.w, ad[k][1][q]);
        ad[k][2][q] = fmaf(hd, e.x, ad[k][2][q]); ad[k][3][q] = fmaf(hd, e.y, ad[k][3][q]);
        ad[k][4][q] = fmaf(hd, e.z, ad[k][4][q]); ad[k][5][q] = fmaf(hd, e.w, ad[k][5][q]);
        ad[k][6][q] = fmaf(hd, f.x, ad[k][6][q]); ad[k][7][q] = fmaf(hd, f.y, ad[k][7][q]);
        ad[k][8][q] = fmaf(hd, f.z, ad[k][8][q]); ad[k][9][q] = fmaf(hd, f.w, ad[k][9][q]);
        ad[k][10][q] = fmaf(hd, g.x, ad[k][10][q]); ad[k][11][q] = fmaf(hd, g.y, ad[k][11][q]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
#pragma unroll
    for (int n = 0; n < 2 * kMaxDz; ++n) pe[k][n] = combine<TPC>(ae[k][n]);
#pragma unroll
    for (int n = 0; n < 2 * kMaxDx; ++n) pd[k][n] = combine<TPC>(ad[k][n]);
  }
}

template <int TPC, int CPL, int MINB>
__global__ void __launch_bounds__(CT, MINB) mc_chain_kernel(const ChainParams p) {
  extern __shared__ __align__(16) float wsm[];  // [Hp][WROW] + enc b1 [4] + dec b1 [12]
  const int tid = threadIdx.x;
  const int H = p.hidden, Hp = (H + 3) & ~3;
  constexpr int dx = kMaxDx, dz = kMaxDz, nn = 2 * kMaxDz + kMaxDx;
  for (int e = tid; e < Hp * WROW; e += CT) {
    const int j = e / WROW, c = e - j * WROW;
    float v = 0.f;
    if (j >= H) v = 0.f;
    else if (c < 6) v = __ldg(p.theta + p.enc0W + c * H + j);
    else if (c == 6) v = __ldg(p.theta + p.enc0b + j);
    else if (c < 11) v = __ldg(p.theta + p.enc1W + j * 4 + (c - 7));
    else if (c < 13) v = __ldg(p.theta + p.dec0W + (c - 11) * H + j);
    else if (c == 13) v = __ldg(p.theta + p.dec0b + j);
    else if (c < 26) v = __ldg(p.theta + p.dec1W + j * 12 + (c - 14));
    wsm[e] = v;
  }
  float* b1e = wsm + Hp * WROW;
  float* b1d = b1e + 4;
  if (tid < 4) b1e[tid] = __ldg(p.theta + p.enc1b + tid);
  if (tid < 12) b1d[tid] = __ldg(p.theta + p.dec1b + tid);
  __syncthreads();

  // lane group g owns chains g * CPL + k
  const int64_t group = ((int64_t)blockIdx.x * CT + tid) / TPC;
  const int sub = tid & (TPC - 1);
  bool live[CPL];
  int64_t cc[CPL];  // idle slots shadow the last chain (full-warp shuffles), writes predicated off
  float x1[CPL][dx];
  double e_old[CPL];
  U128 rs[CPL];
  const U128 jm = {p.jm_hi, p.jm_lo}, ja = {p.ja_hi, p.ja_lo};
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int64_t chain = group * CPL + k;
    live[k] = chain < p.B;
    cc[k] = live[k] ? chain : p.B - 1;
#pragma unroll
    for (int d = 0; d < dx; ++d) x1[k][d] = __ldg(p.x + cc[k] * dx + d);
    if (p.energies_valid) {
      e_old[k] = p.E[cc[k]];
    } else {
      e_old[k] = 0.0;
#pragma unroll
      for (int d = 0; d < dx; ++d) {
        const double t = __dsub_rn((double)x1[k][d], p.means[d]);
        e_old[k] = __dadd_rn(e_old[k], __dmul_rn(t, t));
      }
    }
    rs[k] = U128{0ull, 0ull};
    if (p.use_pcg)
      rs[k] = pcg_advance(U128{p.s0_hi, p.s0_lo}, U128{p.inc_hi, p.inc_lo}, (unsigned long long)(p.chain0 + cc[k]) + 1ull);
  }
  unsigned n_accept = 0, n_unc = 0;

#pragma unroll 1
  for (int step = 0; step < p.n_steps; ++step) {
    // ---- noise of this step: eps(z1) [dz] | eps(z2) [dz] | eps(x2) [dx]
    float nz[CPL][12];
    float z2[CPL][dz];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      if (p.noise) {
#pragma unroll
        for (int i = 0; i < nn; ++i) nz[k][i] = __ldg(p.noise + ((int64_t)step * p.B + cc[k]) * nn + i);
      } else {
        const unsigned long long st = p.step0 + (unsigned long long)step;
        const uint2 key = make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32) ^ (unsigned)(st >> 32));
#pragma unroll
        for (int q = 0; q < (nn + 3) / 4; ++q) {
          const unsigned long long gc = (unsigned long long)(p.chain0 + cc[k]);  // GLOBAL chain index: sharding-invariant noise
          const uint4 rnd = philox4x32(make_uint4((unsigned)gc, (unsigned)(gc >> 32), (unsigned)st, (unsigned)q), key);
          box_muller(rnd.x, rnd.y, nz[k][4 * q], nz[k][4 * q + 1]);
          box_muller(rnd.z, rnd.w, nz[k][4 * q + 2], nz[k][4 * q + 3]);
        }
      }
#pragma unroll
      for (int d = 0; d < dz; ++d) z2[k][d] = nz[k][dz + d];
    }
    // ---- encoder(x1) || decoder(z2)
    float pe1[CPL][2 * dz], pd2[CPL][2 * dx];
    mlp_pair<TPC, CPL>(wsm, Hp, sub, x1, z2, pe1, pd2);
    // ---- samples z1, x2 and the forward log-probabilities (per-dof terms summed in dof order)
    float z1[CPL][dz], x2[CPL][dx];
    float lq1[CPL], lz1[CPL], lz2[CPL], lx2[CPL];
    double e_new[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      lq1[k] = lz1[k] = lz2[k] = lx2[k] = 0.f;
#pragma unroll
      for (int d = 0; d < dz; ++d) {
        const float loc = pe1[k][d] + b1e[d], sc = softplus_tf(pe1[k][dz + d] + b1e[dz + d]);
        z1[k][d] = __fadd_rn(__fmul_rn(nz[k][d], sc), loc);
        lq1[k] += normal_lp(z1[k][d], loc, sc);
        lz1[k] += normal_lp(z1[k][d], 0.f, 1.f);
        lz2[k] += normal_lp(z2[k][d], 0.f, 1.f);
      }
      e_new[k] = 0.0;
#pragma unroll
      for (int d = 0; d < dx; ++d) {
        const float loc = pd2[k][d] + b1d[d], sc = softplus_tf(pd2[k][dx + d] + b1d[dx + d]);
        x2[k][d] = __fadd_rn(__fmul_rn(nz[k][2 * dz + d], sc), loc);
        lx2[k] += normal_lp(x2[k][d], loc, sc);
        const double t = __dsub_rn((double)x2[k][d], p.means[d]);
        e_new[k] = __dadd_rn(e_new[k], __dmul_rn(t, t));
      }
    }
    // ---- encoder(x2) || decoder(z1)
    float pe2[CPL][2 * dz], pd1[CPL][2 * dx];
    mlp_pair<TPC, CPL>(wsm, Hp, sub, x2, z1, pe2, pd1);
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      float lq2 = 0.f, lx1 = 0.f;
#pragma unroll
      for (int d = 0; d < dz; ++d)
        lq2 += normal_lp(z2[k][d], pe2[k][d] + b1e[d], softplus_tf(pe2[k][dz + d] + b1e[dz + d]));
#pragma unroll
      for (int d = 0; d < dx; ++d)
        lx1 += normal_lp(x1[k][d], pd1[k][d] + b1d[d], softplus_tf(pd1[k][dx + d] + b1d[dx + d]));
      // ---- accept / reject (mcmc.py:103, :109, :116-120): every lane of the chain computes the same decision
      const float fwd = __fadd_rn(__fadd_rn(lq1[k], lz2[k]), lx2[k]);
      const float rev = __fadd_rn(__fadd_rn(lq2, lz1[k]), lx1);
      const int64_t g = (int64_t)step * p.B + cc[k];
      const double la = __dsub_rn(__dsub_rn(__dadd_rn(e_new[k], (double)rev), e_old[k]), (double)fwd);
      const bool mine = live[k] && sub == 0;
      double lu;
      if (p.use_pcg) {
        lu = log(pcg_uniform(rs[k]));
        rs[k] = add128(mul128(jm, rs[k]), ja);  // this chain's draw of the next MC step: B_global draws further down the stream
        if (fabs(la - lu) <= 1e-13 * fmax(1.0, fabs(lu))) n_unc += mine ? 1u : 0u;
      } else {
        lu = __ldg(p.log_u + g);
      }
      const bool a = la >= lu;
      if (mine) {
        if (p.log_u_trace) p.log_u_trace[g] = lu;
        if (p.acc_trace) p.acc_trace[g] = a ? 1 : 0;
        if (p.fwd_trace) p.fwd_trace[g] = fwd;
        if (p.rev_trace) p.rev_trace[g] = rev;
        if (p.e_new_trace) p.e_new_trace[g] = e_new[k];
        n_accept += a ? 1u : 0u;
      }
      if (a) {
        e_old[k] = e_new[k];
#pragma unroll
        for (int d = 0; d < dx; ++d) x1[k][d] = x2[k][d];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    if (live[k] && sub == 0) {
      const int64_t chain = group * CPL + k;
#pragma unroll
      for (int d = 0; d < dx; ++d) p.x[chain * dx + d] = x1[k][d];
      p.E[chain] = e_old[k];
    }
  }
  // accepted moves of the CTA -> one atomic per warp
  unsigned w = n_accept;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  if ((tid & 31) == 0 && w) atomicAdd(p.n_acc, (unsigned long long)w);
  if (p.use_pcg && p.n_uncertain) {
    unsigned q = n_unc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((tid & 31) == 0 && q) atomicAdd(p.n_uncertain, (unsigned long long)q);
  }
}

}  // namespace

bool mc_chain_enabled(int dx, int dz) {
  const char* e = getenv("VMS_MC_KERNEL");
  if (e && e[0] == 't') return false;       // "tile": force mc_fused_kernel (cross-check)
  return dx == kMaxDx && dz == kMaxDz;      // built for the C4a shape
}

vms_status mc_chain_run(int dx, int dz, int hidden, const float* theta, float* x, double* E, int energies_valid,
                        const float* noise, unsigned long long seed, unsigned long long step0, const double* log_u,
                        const double* means, int64_t B, int n_steps, unsigned long long* n_acc, uint8_t* acc_trace,
                        float* fwd_trace, float* rev_trace, double* e_new_trace, cudaStream_t st,
                        const vms_pcg64_stream* rng, unsigned long long* n_uncertain, double* log_u_trace, long long chain0) {
  VMS_REQUIRE(dx == kMaxDx && dz == kMaxDz, VMS_ERR_UNSUPPORTED, "mc_chain: built for dx = 6, dz = 2");
  VMS_REQUIRE((log_u != nullptr) != (rng != nullptr), VMS_ERR_INVALID_ARG, "mc_chain: exactly one of log_u / rng");
  ChainParams p = {};
  if (rng) {
    p.use_pcg = 1;
    p.s0_hi = rng->state_hi; p.s0_lo = rng->state_lo; p.inc_hi = rng->inc_hi; p.inc_lo = rng->inc_lo;
    p.jm_hi = rng->stride_mul_hi; p.jm_lo = rng->stride_mul_lo; p.ja_hi = rng->stride_add_hi; p.ja_lo = rng->stride_add_lo;
  }
  p.chain0 = chain0;
  p.n_uncertain = n_uncertain; p.log_u_trace = log_u_trace;
  p.dx = dx; p.dz = dz; p.hidden = hidden;
  int o = 0;
  p.enc0W = o; o += dx * hidden;
  p.enc0b = o; o += hidden;
  p.enc1W = o; o += hidden * 2 * dz;
  p.enc1b = o; o += 2 * dz;
  p.dec0W = o; o += dz * hidden;
  p.dec0b = o; o += hidden;
  p.dec1W = o; o += hidden * 2 * dx;
  p.dec1b = o;
  p.B = B; p.n_steps = n_steps; p.theta = theta; p.x = x; p.E = E; p.energies_valid = energies_valid;
  p.noise = noise; p.seed = seed; p.step0 = step0; p.log_u = log_u; p.means = means; p.n_acc = n_acc;
  p.acc_trace = acc_trace; p.fwd_trace = fwd_trace; p.rev_trace = rev_trace; p.e_new_trace = e_new_trace;
  const size_t smem = (size_t)(((hidden + 3) & ~3) * WROW + 16) * sizeof(float);
  VMS_REQUIRE(smem <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "mc_chain: hidden too large for shared memory");
  // lanes per chain by the number of chains (identical results for every choice, see `combine`)
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // (B200, 65,536 chains x 100 steps: 4 lanes 799 M proposals/s, 2 lanes 1,270 M, 1 lane 1,241 M at 128 registers, 1 lane
  //  with two chains per lane 1,469 M; 32,768 chains: 2 lanes 1,106 M; 8,192: 2 lanes 678 M, 4 lanes 510 M)
  const bool full = B >= (int64_t)sms * 256;  // enough chains for one lane per chain and two chains per lane
  int tpc = full ? 1 : (B >= (int64_t)sms * 40 ? 2 : 4);
  if (const char* e = getenv("VMS_MC_TPC")) {  // cross-checks: force a lane count
    const int t = atoi(e);
    if (t == 1 || t == 2 || t == 4) tpc = t;
  }
  int cpl = full ? 2 : 1;
  if (const char* e = getenv("VMS_MC_CPL")) {  // cross-checks: chains per lane group
    const int t = atoi(e);
    if (t == 1 || t == 2) cpl = t;
  }
  const int64_t groups = (B + cpl - 1) / cpl;
  const unsigned grid = (unsigned)((groups * tpc + CT - 1) / CT);
#define VMS_CHAIN_LAUNCH(T, C, M)                                                                                      \
  do {                                                                                                                 \
    VMS_CUDA(cudaFuncSetAttribute(mc_chain_kernel<T, C, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mc_chain_kernel<T, C, M><<<grid, CT, smem, st>>>(p);                                                              \
  } while (0)
  if (cpl == 1) {
    if (tpc == 1) VMS_CHAIN_LAUNCH(1, 1, 4);
    else if (tpc == 2) VMS_CHAIN_LAUNCH(2, 1, 4);
    else VMS_CHAIN_LAUNCH(4, 1, 4);
  } else {
    if (tpc == 1) VMS_CHAIN_LAUNCH(1, 2, 2);
    else if (tpc == 2) VMS_CHAIN_LAUNCH(2, 2, 2);
    else VMS_CHAIN_LAUNCH(4, 2, 3);
  }
#undef VMS_CHAIN_LAUNCH
  VMS_LAUNCH_CHECK("mc_chain_kernel");
  return VMS_OK;
}

}  // namespace vms
