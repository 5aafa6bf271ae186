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


// ------------------------------------------------------------------------------------------------ a warp per four chains
// One shard of an 8-GPU job holds 8,192 chains: 55 per SM.  With 1 - 4 lanes per chain that is 2 - 7 warps per SM, every
// warp a latency-bound chain of ~10,000 instructions per MC step (ncu at 8,192 chains, two lanes: 4 warps per SM, IPC 0.43
// per warp; four lanes: MORE warp instructions, since every lane repeats the per-chain scalar work -- 0.47 strong-scaling
// efficiency at N = 8).  A first wide variant (8 lanes per chain, unit j -> lane j % 8) cut the warp instructions per
// step to 2,900 but ran at the same 11.4 us per step: the four chains of a warp read the same weight rows, a 16-byte
// shared-memory load is served quarter-warp by quarter-warp (3.7 wavefronts per LDS.128 measured), and the kernel sat at
// 87 % of the shared-memory wavefront peak.  This variant removes that redundancy through registers:
//   * a warp owns FOUR chains; lane l walks the hidden units j = l, l + 32, ... (ascending) and applies each weight row it
//     loads to all four chains (112 FFMAs per 7 loads, every wavefront carries distinct data); a remainder of up to
//     eight units (H = 200) is shared out one unit per lane of each chain instead of a mostly padded round;
//   * the 4 x 16 head outputs are reduce-scattered over the 32 lanes (xor 16, 8, 4, 2, 1: 62 shuffles), which leaves lane
//     l = 8 k + s with (loc, raw scale) of ONE degree of freedom of chain k: s = 0, 1 the encoder's two latents, s = 2..7
//     the decoder's six coordinates -- softplus, the sample and the per-dof log-probability run once per dof;
//   * each lane runs one Philox call (its own eps) instead of three;
//   * samples and log-probability terms are gathered with indexed shuffles and summed in dof order by the eight lanes of
//     a chain, so the decision (float64, mcmc.py:116-120) is the same in all of them; the current configurations of the
//     four chains live in shared memory (24 floats per warp), read back as broadcasts for the next encoder pass.
// Summation order of a hidden layer here: 32 streams (unit j -> stream j % 32) met by the xor-16-8-4-2-1 tree -- NOT the
// four-stream order of the 1 / 2 / 4-lane kernel: log-probabilities agree to float32 rounding, decisions wherever the
// margin exceeds that rounding (both orders are tested against the reference's goldens; VMS_MC_TPC pins one of them when
// bit-identical chains across different shard sizes matter more than speed).
// Measured (B200, 100 steps; two-lane kernel -> this one): 4,096 chains 348 -> 749 M proposals/s, 8,192: 686 -> 983 M,
// 16,384: 956 -> 1,115 M; from 20,480 chains on the two-lane kernel is ahead again (918 vs 893 M; 24,576: 1,094 vs 900 M;
// 32,768: 1,122 vs 1,027 M), so the launcher takes this variant below 128 chains per SM.  ncu at 8,192 chains, first
// version (858 M): 3,440 warp instructions per warp and MC step (45 % of them the FFMAs of the two passes, 18 % the
// reduce-scatter), issue-active 70 % with 3.5 warps per scheduler; the slot permutation and the shared-out remainder round
// then removed ~10 % of the instructions (952 M), the accept uniforms drawn by lane d for step 8 r + d another 3 % (983 M).  Tried and dropped: packed FFMA2 head accumulation (801 vs 858 M: fewer
// issue slots, longer dependent chains), two units in flight per lane (832 M), a 144-register budget without the residual
// spills (789 M).
constexpr int WC = 4;        // chains per warp
constexpr int WL = 8;        // lanes that finish a chain (one per degree of freedom)
constexpr int kWideMaxChainsPerSm = 128;  // chains per SM below which the launcher takes this variant

template <int N>
__device__ __forceinline__ void scatter_step(const float (&v)[N], float (&w)[N / 2], bool upper, int lane_mask) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float snd = upper ? v[i] : v[i + N / 2];
    const float kp = upper ? v[i + N / 2] : v[i];
    w[i] = kp + __shfl_xor_sync(0xffffffffu, snd, lane_mask);
  }
}

// enc(xe[s]) | dec(zd[s]) for the warp's four chains; lane 8 k + d returns (loc, raw scale) of dof d of chain k.
// SLOT s of lane l holds chain s ^ (l >> 3) (slot 0 = the lane's own chain): the partner of the xor-16 / xor-8 exchange then
// needs exactly the upper slots, whatever the lane -- those two levels of the reduce-scatter are a shuffle and an add per
// value, no selects.  Hidden units: `n_full` rounds of 32 units (lane l: unit 32 i + l) for all four chains; when 1..8 units
// remain (H = 200: 6 x 32 + 8) the eight lanes of a chain take one each for their OWN chain only, and those terms join
// after the chain levels of the reduction (a full round for them would be 3/4 padding).
__device__ __forceinline__ void mlp_pair_warp(const float* __restrict__ wsm, int n_full, int tail_base, int lane,
                                              const float (&xe)[WC][kMaxDx], const float (&zd)[WC][kMaxDz], float& out_loc,
                                              float& out_raw) {
  float ae[WC][2 * kMaxDz], ad[WC][2 * kMaxDx];
#pragma unroll
  for (int k = 0; k < WC; ++k) {
#pragma unroll
    for (int n = 0; n < 2 * kMaxDz; ++n) ae[k][n] = 0.f;
#pragma unroll
    for (int n = 0; n < 2 * kMaxDx; ++n) ad[k][n] = 0.f;
  }
  const float* rowp = wsm + lane * WROW;
#pragma unroll 1
  for (int it = 0; it < n_full; ++it, rowp += 32 * WROW) {
    const float4* row = reinterpret_cast<const float4*>(rowp);
    {
      const float4 a = row[0], b = row[1], c = row[2];
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        float he = 0.f;
        he = fmaf(xe[k][0], a.x, he); he = fmaf(xe[k][1], a.y, he); he = fmaf(xe[k][2], a.z, he);
        he = fmaf(xe[k][3], a.w, he); he = fmaf(xe[k][4], b.x, he); he = fmaf(xe[k][5], b.y, he);
        he = fmaxf(he + b.z, 0.f);
        ae[k][0] = fmaf(he, b.w, ae[k][0]); ae[k][1] = fmaf(he, c.x, ae[k][1]);
        ae[k][2] = fmaf(he, c.y, ae[k][2]); ae[k][3] = fmaf(he, c.z, ae[k][3]);
      }
    }
    {
      const float w0d0 = rowp[11];
      const float4 d = row[3], e = row[4], f = row[5], g = row[6];
#pragma unroll
      for (int k = 0; k < WC; ++k) {
        float hd = 0.f;
        hd = fmaf(zd[k][0], w0d0, hd); hd = fmaf(zd[k][1], d.x, hd);
        hd = fmaxf(hd + d.y, 0.f);
        ad[k][0] = fmaf(hd, d.z, ad[k][0]); ad[k][1] = fmaf(hd, d.w, ad[k][1]);
        ad[k][2] = fmaf(hd, e.x, ad[k][2]); ad[k][3] = fmaf(hd, e.y, ad[k][3]);
        ad[k][4] = fmaf(hd, e.z, ad[k][4]); ad[k][5] = fmaf(hd, e.w, ad[k][5]);
        ad[k][6] = fmaf(hd, f.x, ad[k][6]); ad[k][7] = fmaf(hd, f.y, ad[k][7]);
        ad[k][8] = fmaf(hd, f.z, ad[k][8]); ad[k][9] = fmaf(hd, f.w, ad[k][9]);
        ad[k][10] = fmaf(hd, g.x, ad[k][10]); ad[k][11] = fmaf(hd, g.y, ad[k][11]);
      }
    }
  }
  // entries 16 s + i of v: slot s, then the (loc, raw) pairs of the chain's eight dofs
  float v[64];
#pragma unroll
  for (int k = 0; k < WC; ++k) {
    v[16 * k + 0] = ae[k][0]; v[16 * k + 1] = ae[k][2]; v[16 * k + 2] = ae[k][1]; v[16 * k + 3] = ae[k][3];
#pragma unroll
    for (int d = 0; d < kMaxDx; ++d) {
      v[16 * k + 4 + 2 * d] = ad[k][d];
      v[16 * k + 5 + 2 * d] = ad[k][kMaxDx + d];
    }
  }
  // chain levels: slots 2, 3 go to lane ^ 16 (whose slots 2, 3 are this lane's slots 0, 1), then slot 1 to lane ^ 8
  float w32[32], w16[16], w8[8], w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 32; ++i) w32[i] = v[i] + __shfl_xor_sync(0xffffffffu, v[32 + i], 16);
#pragma unroll
  for (int i = 0; i < 16; ++i) w16[i] = w32[i] + __shfl_xor_sync(0xffffffffu, w32[16 + i], 8);
  if (tail_base >= 0) {  // own chain (slot 0), unit tail_base + (lane & 7); rows beyond H are zero
    const float* rp = wsm + (tail_base + (lane & (WL - 1))) * WROW;
    const float4* row = reinterpret_cast<const float4*>(rp);
    const float4 a = row[0], b = row[1], c = row[2], d = row[3], e = row[4], f = row[5], g = row[6];
    float he = 0.f;
    he = fmaf(xe[0][0], a.x, he); he = fmaf(xe[0][1], a.y, he); he = fmaf(xe[0][2], a.z, he);
    he = fmaf(xe[0][3], a.w, he); he = fmaf(xe[0][4], b.x, he); he = fmaf(xe[0][5], b.y, he);
    he = fmaxf(he + b.z, 0.f);
    float hd = 0.f;
    hd = fmaf(zd[0][0], c.w, hd); hd = fmaf(zd[0][1], d.x, hd);
    hd = fmaxf(hd + d.y, 0.f);
    w16[0] = fmaf(he, b.w, w16[0]); w16[1] = fmaf(he, c.y, w16[1]); w16[2] = fmaf(he, c.x, w16[2]); w16[3] = fmaf(he, c.z, w16[3]);
    w16[4] = fmaf(hd, d.z, w16[4]);   w16[5] = fmaf(hd, f.x, w16[5]);
    w16[6] = fmaf(hd, d.w, w16[6]);   w16[7] = fmaf(hd, f.y, w16[7]);
    w16[8] = fmaf(hd, e.x, w16[8]);   w16[9] = fmaf(hd, f.z, w16[9]);
    w16[10] = fmaf(hd, e.y, w16[10]); w16[11] = fmaf(hd, f.w, w16[11]);
    w16[12] = fmaf(hd, e.z, w16[12]); w16[13] = fmaf(hd, g.x, w16[13]);
    w16[14] = fmaf(hd, e.w, w16[14]); w16[15] = fmaf(hd, g.y, w16[15]);
  }
  // dof levels: lane 8 k + d ends with entries 2 d, 2 d + 1
  scatter_step<16>(w16, w8, (lane & 4) != 0, 4);
  scatter_step<8>(w8, w4, (lane & 2) != 0, 2);
  scatter_step<4>(w4, w2, (lane & 1) != 0, 1);
  out_loc = w2[0];
  out_raw = w2[1];
}

template <int CTW, int MINB>
__global__ void __launch_bounds__(CTW, MINB) mc_chain_warp_kernel(const ChainParams p) {
  extern __shared__ __align__(16) float wsm[];  // [Hp32][WROW] + enc b1 [4] + dec b1 [12] + per warp x [WC][8]
  const int tid = threadIdx.x;
  const int H = p.hidden, Hp32 = (H + 31) & ~31;
  constexpr int dx = kMaxDx, dz = kMaxDz, nn = 2 * kMaxDz + kMaxDx;
  for (int e = tid; e < Hp32 * WROW; e += CTW) {
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
  float* b1e = wsm + Hp32 * WROW;
  float* b1d = b1e + 4;
  if (tid < 4) b1e[tid] = __ldg(p.theta + p.enc1b + tid);
  if (tid < 12) b1d[tid] = __ldg(p.theta + p.dec1b + tid);
  __syncthreads();

  const int lane = tid & 31;
  const int rem = H & 31;
  const int n_full = (H >> 5) + (rem > WL ? 1 : 0);
  const int tail_base = (rem > 0 && rem <= WL) ? (H & ~31) : -1;
  float* xs = b1d + 12 + (tid >> 5) * (WC * 8);  // this warp's current configurations: [WC][8] (6 used), 16-byte rows
  const int64_t chain = ((int64_t)blockIdx.x * CTW + tid) / WL;
  const int sub = lane & (WL - 1), kc = lane >> 3;
  const int base = lane & ~(WL - 1);         // first lane of this chain
  const bool enc_lane = sub < dz;            // lanes 0-1: encoder dof `sub`; lanes 2-7: decoder dof `sub - 2`
  const int dof = enc_lane ? sub : sub - dz;
  const bool live = chain < p.B;
  const int64_t cc = live ? chain : p.B - 1;  // idle slots shadow the last chain (full-warp shuffles), writes predicated off
  const float bias_loc = enc_lane ? b1e[dof] : b1d[dof];
  const float bias_raw = enc_lane ? b1e[dz + dof] : b1d[dx + dof];
  float x1_mine = enc_lane ? 0.f : __ldg(p.x + cc * dx + dof);
  xs[kc * 8 + sub] = 0.f;
  __syncwarp();
  if (!enc_lane) xs[kc * 8 + dof] = x1_mine;
  __syncwarp();
  double e_old;
  if (p.energies_valid) {
    e_old = p.E[cc];
  } else {
    e_old = 0.0;
#pragma unroll
    for (int d = 0; d < dx; ++d) {
      const double t = __dsub_rn((double)xs[kc * 8 + d], p.means[d]);
      e_old = __dadd_rn(e_old, __dmul_rn(t, t));
    }
  }
  // accept uniforms: lane d of a chain draws for the MC steps 8 r + d, once per eight steps -- the PCG64 step, the output
  // function and the double log run once per eight steps per warp instead of every step
  U128 rs = {0ull, 0ull};
  U128 jm8 = {p.jm_hi, p.jm_lo}, ja8 = {p.ja_hi, p.ja_lo};
  if (p.use_pcg) {
    rs = pcg_advance(U128{p.s0_hi, p.s0_lo}, U128{p.inc_hi, p.inc_lo}, (unsigned long long)(p.chain0 + cc) + 1ull);
    for (int q = 0; q < sub; ++q) rs = add128(mul128(jm8, rs), ja8);  // this lane's first step: `sub` strides further
#pragma unroll 1
    for (int q = 0; q < 3; ++q) {  // (m, a) -> (m^2, a (m + 1)): the stride of eight MC steps
      ja8 = mul128(add128(jm8, U128{0ull, 1ull}), ja8);
      jm8 = mul128(jm8, jm8);
    }
  }
  double lu_lane = 0.0;
  unsigned n_accept = 0, n_unc = 0;
  // noise layout of a step: eps(z1) [dz] | eps(z2) [dz] | eps(x2) [dx]; this lane's own eps is entry `eps_idx`, which is
  // component `eps_idx & 3` of Philox call `eps_idx >> 2`
  const int eps_idx = enc_lane ? sub : sub + dz;

#pragma unroll 1
  for (int step = 0; step < p.n_steps; ++step) {
    float eps_mine, z2a, z2b;  // z2 of this lane's chain
    if (p.noise) {
      const float* nrow = p.noise + ((int64_t)step * p.B + cc) * nn;
      eps_mine = __ldg(nrow + eps_idx);
      z2a = __ldg(nrow + dz);
      z2b = __ldg(nrow + dz + 1);
    } else {
      const unsigned long long st = p.step0 + (unsigned long long)step;
      const uint2 key = make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32) ^ (unsigned)(st >> 32));
      const unsigned long long gc = (unsigned long long)(p.chain0 + cc);  // GLOBAL chain index: sharding-invariant noise
      const uint4 rnd = philox4x32(make_uint4((unsigned)gc, (unsigned)(gc >> 32), (unsigned)st, (unsigned)(eps_idx >> 2)), key);
      float n0, n1, n2, n3;
      box_muller(rnd.x, rnd.y, n0, n1);
      box_muller(rnd.z, rnd.w, n2, n3);
      const int comp = eps_idx & 3;
      eps_mine = comp == 0 ? n0 : (comp == 1 ? n1 : (comp == 2 ? n2 : n3));
      z2a = __shfl_sync(0xffffffffu, n2, base);  // lane 0 of the chain ran call 0: entries 2, 3 are eps(z2)
      z2b = __shfl_sync(0xffffffffu, n3, base);
    }
    float loc, raw;
    {
      // ---- encoder(x1) || decoder(z2) of the four chains
      float xe[WC][dx], zd[WC][dz];
#pragma unroll
      for (int k = 0; k < WC; ++k) {  // slot k = chain k ^ kc
        const int ch = k ^ kc;
        const float4 lo = *reinterpret_cast<const float4*>(xs + ch * 8);
        const float2 hi = *reinterpret_cast<const float2*>(xs + ch * 8 + 4);
        xe[k][0] = lo.x; xe[k][1] = lo.y; xe[k][2] = lo.z; xe[k][3] = lo.w; xe[k][4] = hi.x; xe[k][5] = hi.y;
        zd[k][0] = __shfl_sync(0xffffffffu, z2a, 8 * ch);
        zd[k][1] = __shfl_sync(0xffffffffu, z2b, 8 * ch);
      }
      mlp_pair_warp(wsm, n_full, tail_base, lane, xe, zd, loc, raw);
    }
    loc += bias_loc;
    float sc = softplus_tf(raw + bias_raw);
    const float smp = __fadd_rn(__fmul_rn(eps_mine, sc), loc);  // z1_d (encoder lanes) or x2_d (decoder lanes)
    const float t_fwd = normal_lp(smp, loc, sc);                 // term of log q(z1 | x1) or of log p(x2 | z2)
    const float z2_mine = sub == 0 ? z2a : z2b;
    const float t_z1 = normal_lp(smp, 0.f, 1.f);                 // encoder lanes: term of log p(z1)
    const float t_z2 = normal_lp(z2_mine, 0.f, 1.f);             // encoder lanes: term of log p(z2)
    float lq1 = 0.f, lz1 = 0.f, lz2 = 0.f, lx2 = 0.f;
#pragma unroll
    for (int d = 0; d < dz; ++d) {
      lq1 += __shfl_sync(0xffffffffu, t_fwd, base + d);
      lz1 += __shfl_sync(0xffffffffu, t_z1, base + d);
      lz2 += __shfl_sync(0xffffffffu, t_z2, base + d);
    }
#pragma unroll
    for (int d = 0; d < dx; ++d) lx2 += __shfl_sync(0xffffffffu, t_fwd, base + dz + d);
    double e_new = 0.0;
#pragma unroll
    for (int d = 0; d < dx; ++d) {
      const double t = __dsub_rn((double)__shfl_sync(0xffffffffu, smp, base + dz + d), p.means[d]);
      e_new = __dadd_rn(e_new, __dmul_rn(t, t));
    }
    {
      // ---- encoder(x2) || decoder(z1) of the four chains
      float xe[WC][dx], zd[WC][dz];
#pragma unroll
      for (int k = 0; k < WC; ++k) {  // slot k = chain k ^ kc
        const int src = 8 * (k ^ kc);
#pragma unroll
        for (int d = 0; d < dz; ++d) zd[k][d] = __shfl_sync(0xffffffffu, smp, src + d);
#pragma unroll
        for (int d = 0; d < dx; ++d) xe[k][d] = __shfl_sync(0xffffffffu, smp, src + dz + d);
      }
      mlp_pair_warp(wsm, n_full, tail_base, lane, xe, zd, loc, raw);
    }
    loc += bias_loc;
    sc = softplus_tf(raw + bias_raw);
    const float t_rev = normal_lp(enc_lane ? z2_mine : x1_mine, loc, sc);  // term of log q(z2 | x2) or of log p(x1 | z1)
    float lq2 = 0.f, lx1 = 0.f;
#pragma unroll
    for (int d = 0; d < dz; ++d) lq2 += __shfl_sync(0xffffffffu, t_rev, base + d);
#pragma unroll
    for (int d = 0; d < dx; ++d) lx1 += __shfl_sync(0xffffffffu, t_rev, base + dz + d);
    // ---- accept / reject (mcmc.py:103, :109, :116-120): every lane of the chain computes the same decision
    const float fwd = __fadd_rn(__fadd_rn(lq1, lz2), lx2);
    const float rev = __fadd_rn(__fadd_rn(lq2, lz1), lx1);
    const int64_t g = (int64_t)step * p.B + cc;
    const double la = __dsub_rn(__dsub_rn(__dadd_rn(e_new, (double)rev), e_old), (double)fwd);
    const bool mine = live && sub == 0;
    double lu;
    if (p.use_pcg) {
      if ((step & (WL - 1)) == 0) {
        lu_lane = log(pcg_uniform(rs));
        rs = add128(mul128(jm8, rs), ja8);  // this lane's draw eight MC steps on: 8 B_global draws further down the stream
      }
      lu = __shfl_sync(0xffffffffu, lu_lane, base + (step & (WL - 1)));
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
      if (p.e_new_trace) p.e_new_trace[g] = e_new;
      n_accept += a ? 1u : 0u;
    }
    if (a) {
      e_old = e_new;
      if (!enc_lane) {
        x1_mine = smp;
        xs[kc * 8 + dof] = smp;
      }
    }
    __syncwarp();
  }
  if (live && !enc_lane) p.x[chain * dx + dof] = x1_mine;
  if (live && sub == 0) p.E[chain] = e_old;
  unsigned w = n_accept;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  if (lane == 0 && w) atomicAdd(p.n_acc, (unsigned long long)w);
  if (p.use_pcg && p.n_uncertain) {
    unsigned q = n_unc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0 && q) atomicAdd(p.n_uncertain, (unsigned long long)q);
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
  // rows padded for the warp-per-four-chains variant too, plus its per-warp configuration slots (up to 8 warps)
  const size_t smem = (size_t)(((hidden + 31) & ~31) * WROW + 16 + 8 * WC * 8) * sizeof(float);
  VMS_REQUIRE(smem <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "mc_chain: hidden too large for shared memory");
  // lanes per chain by the number of chains (identical results for every choice, see `combine`)
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // (B200, 65,536 chains x 100 steps: 4 lanes 799 M proposals/s, 2 lanes 1,270 M, 1 lane 1,241 M at 128 registers, 1 lane
  //  with two chains per lane 1,469 M; 32,768 chains: 2 lanes 1,106 M; 8,192: 2 lanes 678 M, 4 lanes 510 M)
  const bool full = B >= (int64_t)sms * 256;  // enough chains for one lane per chain and two chains per lane
  //  8 lanes (mc_chain_warp_kernel, its own summation order): see the numbers at the launch below
  int tpc = full ? 1 : (B >= (int64_t)sms * kWideMaxChainsPerSm ? 2 : 8);
  if (const char* e = getenv("VMS_MC_TPC")) {  // cross-checks / pinning one summation order: force a lane count
    const int t = atoi(e);
    if (t == 1 || t == 2 || t == 4 || t == 8) tpc = t;
  }
  if (tpc == 8) {
    // CTAs of 64 threads (8 chains) while they are all co-resident (finest balance over the SMs), else 128 / 256
    const int64_t lanes = B * WL;
    const int ctw = lanes <= (int64_t)sms * 8 * 64 ? 64 : (lanes <= (int64_t)sms * 8 * 128 ? 128 : 256);
    const unsigned gridw = (unsigned)((lanes + ctw - 1) / ctw);
#define VMS_WIDE_LAUNCH(C, M)                                                                                          \
  do {                                                                                                                 \
    VMS_CUDA(cudaFuncSetAttribute(mc_chain_warp_kernel<C, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mc_chain_warp_kernel<C, M><<<gridw, C, smem, st>>>(p);                                                             \
  } while (0)
    if (ctw == 64) VMS_WIDE_LAUNCH(64, 7);
    else if (ctw == 128) VMS_WIDE_LAUNCH(128, 3);
    else VMS_WIDE_LAUNCH(256, 1);
#undef VMS_WIDE_LAUNCH
    VMS_LAUNCH_CHECK("mc_chain_warp_kernel");
    return VMS_OK;
  }
  int cpl = full ? 2 : 1;
  if (!full && tpc == 2) {
    // two chains per lane pair halve the weight loads per FFMA (32,768 chains: 1,118 -> 1,207 M proposals/s) but double the
    // chains per CTA: taken when the fullest SM then holds no more chains than before (24,576 chains: 3 CTAs of 64 chains
    // against 2 of 128 on the fullest SM -- 1,094 vs 905 M)
    auto fullest = [&](int c) {
      const int64_t per_cta = (int64_t)(CT / 2) * c, n_cta = (B + per_cta - 1) / per_cta;
      return (n_cta + sms - 1) / sms * per_cta;
    };
    if (fullest(2) <= fullest(1)) cpl = 2;
  }
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
