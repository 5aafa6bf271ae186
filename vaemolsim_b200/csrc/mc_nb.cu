// mc_nb.cu -- fused MC kernel for the model family of examples/MC_Moves_with_VAEs.ipynb (C4b): whole MC steps of
// B chains in one launch, four lanes per chain, chain state in registers, every weight resident in shared memory.
//
// The step is mcmc.py:68-130 (`MCMC.single_step`) for
//   encoder  FCDeepNN(2 -> He -> 2, relu) into tfp.layers.IndependentNormal(1)                          (notebook cell 11)
//   prior    FlowedDistribution(RQSSplineMAF over a ONE-dimensional latent, Independent N(0, 1))         (cell 14)
//   decoder  FCDeepNN(1 -> Hd -> (2, 2), relu) into AutoregressiveBlockwise(2, [Normal] * 2, conditional on z,
//            MADE hidden_units [h0, h1, h2])                                                              (cell 17)
//   energy   log-density of a mixture of independent Normals (cell 5, 38), float32 like `log_prob(...).numpy()`
// evaluated as the op-by-op kernels evaluate it (same formulas, float32, sums in a different order), so that the two paths
// agree to float32 rounding and both reproduce the decisions of the reference driver (tests/golden/mcmc_reference_c4b.npz).
//
// What makes one launch possible:
//  * a masked autoregressive flow over ONE dimension has input-independent spline parameters (the MADE mask of a
//    one-dimensional event is empty: only the biases reach the output).  The host evaluates the conditioner networks once per
//    call on one row (the ordinary dense kernels, any depth / activation) and vms_rqs_knot_table turns the raw parameters
//    into a knot table per block (float64 knot positions as in rqs_device.cuh `find_bin`, float32 derivatives); a chain
//    applies a block with a binary search over 21 shared-memory knots instead of 3 networks + softmax per element.
//  * tfp's Autoregressive sampling is D + 1 = 3 sequential MADE passes with the SAME noise (dists.py:338-340), log_prob one
//    more; with the reverse move's pass that is five passes of a [3 -> h0 -> h1 -> h2 -> 4] network per proposal.  A lane
//    computes the thin outer layers whole and every fourth unit of the wide middle layer, streaming its weights as 16-byte
//    shared-memory broadcasts (row of unit j: W1[:, j] | W2[j, :] | b1[j], Wc1[j]); the h2 partial sums meet in a two-step
//    butterfly.  The two FCDeepNN evaluations of each half step (encoder(x1) with decoder-mapping(z2), then encoder(x2)
//    with decoder-mapping(z1)) run side by side the same way (mc_chain.cu).
//  * accept uniforms: the host's PCG64 stream, uploaded (log_u) or regenerated on the device by LCG jump-ahead (mc_rng.cuh);
//    the acceptance arithmetic is NumPy's float32 evaluation of mcmc.py:116 for a float32 energy callback.
#include "common.cuh"
#include "mc_rng.cuh"
#include "rqs_device.cuh"
#include <math.h>
#include <stdlib.h>

namespace vms {

using namespace mcdev;

namespace {

constexpr int CT = 128;   // threads per CTA
constexpr int MROW = 12;  // floats per FCDeepNN hidden unit: enc W0[0..1][j] b0[j] W1[j][0..1] | dec W0[0][j] b0[j] W1[j][0..3] | pad
constexpr int DX = 2, DZ = 1, NN = 2 * DZ + DX;

struct NbParams {
  vms_mc_nb_model m;
  int64_t B;
  int n_steps;
  float* x;
  float* E;
  int energies_valid;
  const float* noise;
  unsigned long long seed, step0;
  const double* log_u;
  int use_pcg;
  unsigned long long s0_hi, s0_lo, inc_hi, inc_lo, jm_hi, jm_lo, ja_hi, ja_lo;
  int64_t chain0;
  unsigned long long *n_acc, *n_uncertain;
  uint8_t* acc_trace;
  float *fwd_trace, *rev_trace, *e_new_trace;
  double* log_u_trace;
};

__host__ __device__ inline int table_stride(int K) { return 2 * (K + 1) + (K + 2) / 2; }  // in doubles
__host__ __device__ inline int r4(int v) { return (v + 3) & ~3; }

// shared-memory image (offsets in floats; every region starts on a 16-byte boundary, the knot tables come first for their
// 8-byte alignment).  Hidden-unit rows are padded with zero rows to a multiple of 4 units: every lane count walks the
// same four interleaved unit streams (below), a zero row adds +0 to its stream.
struct Layout {
  int tables, mlp, eb1, db1, l0, mid, tail, gmm, total;
  int Hp, H1p, RS;
};
__host__ __device__ inline Layout make_layout(const vms_mc_nb_model& m, int H0P, int H2P) {
  Layout L;
  L.Hp = r4(m.enc_hidden > m.dec_hidden ? m.enc_hidden : m.dec_hidden);
  L.H1p = r4(m.made_hidden[1]);
  L.RS = r4(H0P + H2P + 2);                // row of middle unit j: W1[:, j] [H0P] | W2[j, :] [H2P] | b1[j] | Wc1[j] | pad
  int o = 0;
  L.tables = o; o = r4(o + 2 * m.n_blocks * table_stride(m.n_bins));
  L.mlp = o;    o += L.Hp * MROW;
  L.eb1 = o;    o += 4;
  L.db1 = o;    o += 4;
  L.l0 = o;     o += 4 * H0P;              // per unit i: W0[0][i], W0[1][i], Wc0[i], b0[i]
  L.mid = o;    o += L.H1p * L.RS;
  L.tail = o;   o += 6 * H2P + 8;          // b2 [H2P] | Wc2 [H2P] | W3 [H2P][4] | b3 [4] | Wc3 [4]
  L.gmm = o;    o += r4(7 * m.n_comp);     // log_w [n] | per (component, dim): scale, loc / scale, 0.5 log 2pi + log scale
  L.total = r4(o);
  return L;
}

// tfp Normal._log_prob for N(0, 1): normal_lp(x, 0, 1) with its exact simplifications (x / 1 - 0 / 1 = x, log 1 = 0)
__device__ __forceinline__ float std_normal_lp(float x) { return -0.5f * x * x - VMS_HALF_LOG_2PI; }

// One block of the MAF prior from its knot table (the formulas of rqs_device.cuh `octet_apply`).
__device__ __noinline__ float2 spline_apply(const double* __restrict__ tb, int K, float bin_min, float v, bool inv) {
  const double* kx = tb;
  const double* ky = tb + (K + 1);
  const float* dks = reinterpret_cast<const float*>(tb + 2 * (K + 1));
  const double* ks = inv ? ky : kx;
  const double vd = (double)v;
  // bin k covers [knot k, knot k+1); the range edges themselves are outside (TFP: identity)
  if (!(vd > (double)bin_min && vd < ks[K])) return make_float2(v, 0.f);
  int lo = 0, hi = K;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ks[mid] <= vd) lo = mid; else hi = mid;
  }
  rqsdev::Bin b;
  b.lo_x = kx[lo]; b.lo_y = ky[lo];
  b.wk = (float)(kx[lo + 1] - b.lo_x);
  b.hk = (float)(ky[lo + 1] - b.lo_y);
  const float dk = dks[lo], dk1 = dks[lo + 1];
  const float sk = b.hk / b.wk;
  const float rr = rqsdev::rel_pos(b, vd, sk, dk, dk1, inv);
  const float omr = 1.f - rr, u = rr * omr;
  const float den = sk + (dk1 + dk - 2.f * sk) * u;
  float out;
  if (!inv) {
    const float num = b.hk * (sk * rr * rr + dk * u);
    out = (float)(b.lo_y + (double)(num / den));
  } else {
    out = (float)(b.lo_x + (double)(rr * b.wk));
  }
  const float Pq = dk1 * rr * rr + 2.f * sk * u + dk * omr * omr;
  const float ldj = logf((sk * sk) * Pq / (den * den));
  return make_float2(out, inv ? -ldj : ldj);
}

// The hidden units of every wide layer are walked as FOUR interleaved streams (units j = s, s + 4, ... for s = 0..3), each
// summed in ascending j, and the four partial sums meet as (s0 + s1) + (s2 + s3).  A chain is owned by TPC lanes (4, 2 or 1),
// each taking 4 / TPC streams: the summation order, hence every bit of every result and every decision, is the same for
// every TPC, so the launcher may pick the lane count by the number of chains (one lane per chain when there are enough chains
// to fill the GPU: no replicated scalar work, a quarter of the shared-memory traffic; four lanes when chains are scarce,
// e.g. a shard of a multi-GPU job).
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

// encoder(x0, x1) and decoder-mapping(zd) side by side
template <int TPC>
__device__ __forceinline__ void mlp_pair(const float* __restrict__ rows, int Hp, int sub, float x0, float x1, float zd,
                                         float (&pe)[2], float (&pd)[4]) {
  constexpr int NS = 4 / TPC;
  float ae[2][NS], ad[4][NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) {
    ae[0][q] = ae[1][q] = 0.f;
    ad[0][q] = ad[1][q] = ad[2][q] = ad[3][q] = 0.f;
  }
#pragma unroll 2
  for (int j0 = 0; j0 < Hp; j0 += 4) {
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const float4* row = reinterpret_cast<const float4*>(rows + (j0 + sub + TPC * q) * MROW);
      const float4 a = row[0], b = row[1], c = row[2];
      float he = fmaf(x1, a.y, x0 * a.x);
      he = fmaxf(he + a.z, 0.f);
      ae[0][q] = fmaf(he, a.w, ae[0][q]);
      ae[1][q] = fmaf(he, b.x, ae[1][q]);
      const float hd = fmaxf(fmaf(zd, b.y, b.z), 0.f);
      ad[0][q] = fmaf(hd, b.w, ad[0][q]);
      ad[1][q] = fmaf(hd, c.x, ad[1][q]);
      ad[2][q] = fmaf(hd, c.y, ad[2][q]);
      ad[3][q] = fmaf(hd, c.z, ad[3][q]);
    }
  }
  pe[0] = combine<TPC>(ae[0]); pe[1] = combine<TPC>(ae[1]);
#pragma unroll
  for (int n = 0; n < 4; ++n) pd[n] = combine<TPC>(ad[n]);
}

// activation of tfp's AutoregressiveNetwork hidden layers: none / relu as one max (lo = -inf / 0), tanh on a uniform flag
__device__ __forceinline__ float act_fn(float v, float lo, bool is_tanh) {
  v = fmaxf(v, lo);
  return is_tanh ? tanhf(v) : v;
}

// AutoregressiveNetwork(params = 2, event 2, conditional 1, hidden [h0, h1, h2]) on (s0, s1 | cond) -> out [dof][param]
template <int TPC, int H0P, int H2P>
__device__ __forceinline__ float4 made_pass(const float* __restrict__ sm, const Layout& L, float act_lo, bool is_tanh, int sub,
                                            float s0, float s1, float cond) {
  constexpr int NS = 4 / TPC;
  constexpr int RS = (H0P + H2P + 2 + 3) & ~3;
  float a0[H0P];
  const float4* l0 = reinterpret_cast<const float4*>(sm + L.l0);
#pragma unroll
  for (int i = 0; i < H0P; ++i) {
    const float4 w = l0[i];
    a0[i] = act_fn(fmaf(cond, w.z, fmaf(s1, w.y, s0 * w.x)) + w.w, act_lo, is_tanh);  // (padding units: act(0) = 0)
  }
  float a2[H2P][NS];
#pragma unroll
  for (int k = 0; k < H2P; ++k)
#pragma unroll
    for (int q = 0; q < NS; ++q) a2[k][q] = 0.f;
#pragma unroll 1
  for (int j0 = 0; j0 < L.H1p; j0 += 4) {
#pragma unroll
    for (int q = 0; q < NS; ++q) {
      const float4* row = reinterpret_cast<const float4*>(sm + L.mid + (j0 + sub + TPC * q) * RS);
      float w[RS];
#pragma unroll
      for (int t = 0; t < RS / 4; ++t) {
        const float4 r = row[t];
        w[4 * t] = r.x; w[4 * t + 1] = r.y; w[4 * t + 2] = r.z; w[4 * t + 3] = r.w;
      }
      float h = cond * w[H0P + H2P + 1];
#pragma unroll
      for (int i = 0; i < H0P; ++i) h = fmaf(a0[i], w[i], h);
      h = act_fn(h + w[H0P + H2P], act_lo, is_tanh);
#pragma unroll
      for (int k = 0; k < H2P; ++k) a2[k][q] = fmaf(h, w[H0P + k], a2[k][q]);
    }
  }
  const float* tl = sm + L.tail;
  const float4 bb = reinterpret_cast<const float4*>(tl + 6 * H2P)[0], wc = reinterpret_cast<const float4*>(tl + 6 * H2P)[1];
  float o0 = fmaf(cond, wc.x, bb.x), o1 = fmaf(cond, wc.y, bb.y), o2 = fmaf(cond, wc.z, bb.z), o3 = fmaf(cond, wc.w, bb.w);
  const float4* W3 = reinterpret_cast<const float4*>(tl + 2 * H2P);
#pragma unroll
  for (int k = 0; k < H2P; ++k) {
    const float v = act_fn(fmaf(cond, tl[H2P + k], combine<TPC>(a2[k])) + tl[k], act_lo, is_tanh);
    const float4 w = W3[k];
    o0 = fmaf(v, w.x, o0); o1 = fmaf(v, w.y, o1); o2 = fmaf(v, w.z, o2); o3 = fmaf(v, w.w, o3);
  }
  return make_float4(o0, o1, o2, o3);
}

// tfp Autoregressive sampling without mask knowledge (cold path): sample0 = ones, D + 1 passes with the same noise
template <int TPC, int H0P, int H2P>
__device__ __noinline__ float2 ar_sample_generic(const float* __restrict__ sm, const Layout L, float act_lo, bool is_tanh,
                                                 int sub, float4 in, float e0, float e1, float cond) {
  float s0 = 1.f, s1 = 1.f;
#pragma unroll 1
  for (int pass = 0; pass < DX + 1; ++pass) {
    const float4 mo = made_pass<TPC, H0P, H2P>(sm, L, act_lo, is_tanh, sub, s0, s1, cond);
    const float l0 = in.x + mo.x, c0 = softplus_tf(in.y + mo.y) + VMS_EPS32;
    const float l1 = in.z + mo.z, c1 = softplus_tf(in.w + mo.w) + VMS_EPS32;
    s0 = __fadd_rn(__fmul_rn(e0, c0), l0);
    s1 = __fadd_rn(__fmul_rn(e1, c1), l1);
  }
  return make_float2(s0, s1);
}

// Mixture log-density as tfp evaluates it in float32 (mcmc.cu energy_gmm_kernel): per component sum_d Normal log_prob
// (x / s - m / s form; m / s and 0.5 log 2pi + log s staged) + log cat prob, max-shifted logsumexp over the components.
__device__ __forceinline__ float gmm_component(const float* __restrict__ g, int n, int k, float x0, float x1) {
  const float* c = g + n + 6 * k;
  const float z0 = x0 / c[0] - c[1], z1 = x1 / c[3] - c[4];
  float s = 0.f;
  s += -0.5f * z0 * z0 - c[2];
  s += -0.5f * z1 * z1 - c[5];
  return s + g[k];
}
__device__ __forceinline__ float gmm_energy(const float* __restrict__ g, int n, float x0, float x1) {
  if (n <= 4) {  // the notebook's three components: everything in registers
    float lp[4];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lp[k] = k < n ? gmm_component(g, n, k, x0, x1) : -INFINITY;
      mx = fmaxf(mx, lp[k]);
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < n) acc += expf(lp[k] - mx);
    return mx + logf(acc);
  }
  float lp[16];
  float mx = -INFINITY;
#pragma unroll 1
  for (int k = 0; k < n; ++k) {
    lp[k] = gmm_component(g, n, k, x0, x1);
    mx = fmaxf(mx, lp[k]);
  }
  float acc = 0.f;
#pragma unroll 1
  for (int k = 0; k < n; ++k) acc += expf(lp[k] - mx);
  return mx + logf(acc);
}
__device__ __noinline__ float gmm_energy_cold(const float* __restrict__ g, int n, float x0, float x1) {
  return gmm_energy(g, n, x0, x1);
}

template <int TPC, int H0P, int H2P, int MINB>
__global__ void __launch_bounds__(CT, MINB) mc_nb_kernel(const NbParams p) {
  extern __shared__ __align__(16) float sm[];
  const vms_mc_nb_model& m = p.m;
  const Layout L = make_layout(m, H0P, H2P);
  const int tid = threadIdx.x;
  const int K = m.n_bins, nb = m.n_blocks, TS = table_stride(K);
  const int h0 = m.made_hidden[0], h1 = m.made_hidden[1], h2 = m.made_hidden[2];
  // ---- stage the model
  {
    const float* src = reinterpret_cast<const float*>(m.tables);
    for (int e = tid; e < 2 * nb * TS; e += CT) sm[L.tables + e] = __ldg(src + e);
    for (int e = tid; e < L.Hp * MROW; e += CT) {
      const int j = e / MROW, c = e - j * MROW;
      float v = 0.f;
      if (j < m.enc_hidden) {
        if (c < 2) v = __ldg(m.enc_W0 + c * m.enc_hidden + j);
        else if (c == 2) v = __ldg(m.enc_b0 + j);
        else if (c < 5) v = __ldg(m.enc_W1 + j * 2 + (c - 3));
      }
      if (j < m.dec_hidden) {
        if (c == 5) v = __ldg(m.dec_W0 + j);
        else if (c == 6) v = __ldg(m.dec_b0 + j);
        else if (c >= 7 && c < 11) v = __ldg(m.dec_W1 + j * 4 + (c - 7));
      }
      sm[L.mlp + e] = v;
    }
    if (tid < 2) sm[L.eb1 + tid] = __ldg(m.enc_b1 + tid);
    if (tid < 4) sm[L.db1 + tid] = __ldg(m.dec_b1 + tid);
    for (int e = tid; e < 4 * H0P; e += CT) {
      const int i = e >> 2, c = e & 3;
      float v = 0.f;
      if (i < h0) v = c < 2 ? __ldg(m.made_W[0] + c * h0 + i) : (c == 2 ? __ldg(m.made_Wc[0] + i) : __ldg(m.made_b[0] + i));
      sm[L.l0 + e] = v;
    }
    for (int e = tid; e < L.H1p * L.RS; e += CT) {
      const int j = e / L.RS, c = e - j * L.RS;
      float v = 0.f;
      if (j < h1) {
        if (c < H0P) { if (c < h0) v = __ldg(m.made_W[1] + c * h1 + j); }
        else if (c < H0P + H2P) { if (c - H0P < h2) v = __ldg(m.made_W[2] + j * h2 + (c - H0P)); }
        else if (c == H0P + H2P) v = __ldg(m.made_b[1] + j);
        else if (c == H0P + H2P + 1) v = __ldg(m.made_Wc[1] + j);
      }
      sm[L.mid + e] = v;
    }
    for (int e = tid; e < 6 * H2P + 8; e += CT) {
      float v = 0.f;
      if (e < H2P) { if (e < h2) v = __ldg(m.made_b[2] + e); }
      else if (e < 2 * H2P) { if (e - H2P < h2) v = __ldg(m.made_Wc[2] + (e - H2P)); }
      else if (e < 6 * H2P) { const int k = (e - 2 * H2P) >> 2, n = (e - 2 * H2P) & 3; if (k < h2) v = __ldg(m.made_W[3] + k * 4 + n); }
      else if (e < 6 * H2P + 4) v = __ldg(m.made_b[3] + (e - 6 * H2P));
      else v = __ldg(m.made_Wc[3] + (e - 6 * H2P - 4));
      sm[L.tail + e] = v;
    }
    const int n = m.n_comp;
    if (tid < n) sm[L.gmm + tid] = __ldg(m.gmm_log_w + tid);
    for (int e = tid; e < 2 * n; e += CT) {  // (component, dim) pairs
      const float sc = __ldg(m.gmm_scale + e), lc = __ldg(m.gmm_loc + e);
      float* c = sm + L.gmm + n + 3 * e;
      c[0] = sc;
      c[1] = lc / sc;
      c[2] = VMS_HALF_LOG_2PI + logf(sc);
    }
  }
  __syncthreads();
  const double* tables = reinterpret_cast<const double*>(sm + L.tables);
  const float* rows = sm + L.mlp;
  const float* eb1 = sm + L.eb1;
  const float* db1 = sm + L.db1;
  const float* gmm = sm + L.gmm;
  const bool is_tanh = m.made_act == VMS_ACT_TANH;
  const float act_lo = m.made_act == VMS_ACT_RELU ? 0.f : -INFINITY;
  const int first = m.made_first_dof;

  const int64_t chain = ((int64_t)blockIdx.x * CT + tid) / TPC;
  const int sub = tid & (TPC - 1);
  const bool live = chain < p.B;
  const int64_t cc = live ? chain : p.B - 1;  // idle lanes shadow the last chain (full-warp shuffles), writes predicated off
  float x1[DX];
  x1[0] = __ldg(p.x + cc * DX);
  x1[1] = __ldg(p.x + cc * DX + 1);
  float e_old = p.energies_valid ? p.E[cc] : gmm_energy_cold(gmm, m.n_comp, x1[0], x1[1]);
  unsigned n_accept = 0, n_unc = 0;
  U128 rs = {0ull, 0ull};
  const U128 jm = {p.jm_hi, p.jm_lo}, ja = {p.ja_hi, p.ja_lo};
  if (p.use_pcg)
    rs = pcg_advance(U128{p.s0_hi, p.s0_lo}, U128{p.inc_hi, p.inc_lo}, (unsigned long long)(p.chain0 + cc) + 1ull);

#pragma unroll 1
  for (int step = 0; step < p.n_steps; ++step) {
    // ---- noise of this step in the order the reference draws it (mcmc.py:100-102): eps(z1) | eps(z2) | eps(x2) [2]
    float nz[NN];
    if (p.noise) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.noise) + ((int64_t)step * p.B + cc));
      nz[0] = t.x; nz[1] = t.y; nz[2] = t.z; nz[3] = t.w;
    } else {
      const unsigned long long st = p.step0 + (unsigned long long)step;
      const uint2 key = make_uint2((unsigned)p.seed, (unsigned)(p.seed >> 32) ^ (unsigned)(st >> 32));
      const unsigned long long gc = (unsigned long long)(p.chain0 + cc);  // GLOBAL chain index: sharding-invariant noise
      const uint4 rnd = philox4x32(make_uint4((unsigned)gc, (unsigned)(gc >> 32), (unsigned)st, 0u), key);
      box_muller(rnd.x, rnd.y, nz[0], nz[1]);
      box_muller(rnd.z, rnd.w, nz[2], nz[3]);
    }
    // Forward move (dir 0): z2 ~ prior, z1 ~ encoder(x1), x2 ~ decoder(z2) with their log-probabilities (mcmc.py:100-103);
    // reverse move (dir 1): encoder(x2).log_prob(z2), prior.log_prob(z1), decoder(z1).log_prob(x1) (mcmc.py:106-109).
    // Both halves are: the prior's spline chain, encoder || decoder-mapping, one MADE pass, a blockwise Normal log_prob.
    float z1 = 0.f, z2 = 0.f, x2[DX] = {0.f, 0.f}, e_new = 0.f;
    float lq0 = 0.f, lq1 = 0.f, lz0 = 0.f, lz1 = 0.f, lx0 = 0.f, lx1 = 0.f;
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
      // ---- prior: base noise through the blocks in sampling direction (log p = log N(eps) - sum fldj), or z1 back through
      // them (log p = log N(base) + sum ildj)
      float v = dir ? z1 : nz[1], ld = 0.f;
#pragma unroll 1
      for (int b = 0; b < nb; ++b) {
        const float2 r = spline_apply(tables + (dir ? nb - 1 - b : b) * TS, K, m.range_min, v, dir != 0);
        v = r.x;
        ld += r.y;
      }
      if (!dir) { z2 = v; lz0 = std_normal_lp(nz[1]) - ld; } else { lz1 = std_normal_lp(v) + ld; }
      // ---- encoder(x) || decoder-mapping(z)
      const float zd = dir ? z1 : z2;
      float pe[2], pd[4];
      mlp_pair<TPC>(rows, L.Hp, sub, dir ? x2[0] : x1[0], dir ? x2[1] : x1[1], zd, pe, pd);
      const float loc = pe[0] + eb1[0], sc = softplus_tf(pe[1] + eb1[1]);
      if (!dir) z1 = __fadd_rn(__fmul_rn(nz[0], sc), loc);
      { const float t = normal_lp(dir ? z2 : z1, loc, sc); if (dir) lq1 = t; else lq0 = t; }
      const float4 in = make_float4(pd[0] + db1[0], pd[1] + db1[1], pd[2] + db1[2], pd[3] + db1[3]);
      // ---- decoder: tfp Autoregressive over a Blockwise Normal whose raw parameters are in + MADE(samples | z)
      float s0 = x1[0], s1 = x1[1];
      if (!dir) {
        s0 = s1 = 1.f;  // sample0 = ones (dists.py:338)
        if (first >= 0) {
          // The masks make the first dof's parameters a function of the conditional input alone (its output columns of the
          // last kernel are zero: fmaf(v, 0, o) == o) and hide the other dof from every hidden unit (its input row is zero),
          // so tfp's passes 2, 3 and the log_prob pass see bit-identical hidden activations: ONE pass below gives them all.
          const bool f = first != 0;
          const float4 bb = reinterpret_cast<const float4*>(sm + L.tail + 6 * H2P)[0];
          const float4 wc = reinterpret_cast<const float4*>(sm + L.tail + 6 * H2P)[1];
          const float lf = (f ? in.z : in.x) + fmaf(zd, f ? wc.z : wc.x, f ? bb.z : bb.x);
          const float cf = softplus_tf((f ? in.w : in.y) + fmaf(zd, f ? wc.w : wc.y, f ? bb.w : bb.y)) + VMS_EPS32;
          const float sf = __fadd_rn(__fmul_rn(f ? nz[3] : nz[2], cf), lf);
          if (f) s1 = sf; else s0 = sf;
        } else {
          const float2 s = ar_sample_generic<TPC, H0P, H2P>(sm, L, act_lo, is_tanh, sub, in, nz[2], nz[3], zd);
          s0 = s.x;
          s1 = s.y;
        }
      }
      const float4 mo = made_pass<TPC, H0P, H2P>(sm, L, act_lo, is_tanh, sub, s0, s1, zd);
      const float l0 = in.x + mo.x, c0 = softplus_tf(in.y + mo.y) + VMS_EPS32;
      const float l1 = in.z + mo.z, c1 = softplus_tf(in.w + mo.w) + VMS_EPS32;
      if (!dir && first >= 0) {  // the second dof's sample from the same pass
        if (first != 0) s0 = __fadd_rn(__fmul_rn(nz[2], c0), l0);
        else s1 = __fadd_rn(__fmul_rn(nz[3], c1), l1);
      }
      float lp = 0.f;
      lp += normal_lp(s0, l0, c0);
      lp += normal_lp(s1, l1, c1);
      if (dir) lx1 = lp; else lx0 = lp;
      if (!dir) {
        x2[0] = s0;
        x2[1] = s1;
        e_new = gmm_energy(gmm, m.n_comp, s0, s1);
      }
    }
    // ---- accept / reject (mcmc.py:103, :109, :116-120), NumPy's float32 evaluation for a float32 energy
    const float fwd = __fadd_rn(__fadd_rn(lq0, lz0), lx0);
    const float rev = __fadd_rn(__fadd_rn(lq1, lz1), lx1);
    const int64_t g = (int64_t)step * p.B + cc;
    const double la = (double)__fsub_rn(__fsub_rn(__fadd_rn(e_new, rev), e_old), fwd);
    double lu;
    if (p.use_pcg) {
      lu = log(pcg_uniform(rs));
      rs = add128(mul128(jm, rs), ja);
      if (fabs(la - lu) <= 1e-13 * fmax(1.0, fabs(lu))) n_unc += (live && sub == 0) ? 1u : 0u;
    } else {
      lu = __ldg(p.log_u + g);
    }
    const bool a = la >= lu;
    if (live && sub == 0) {
      if (p.log_u_trace) p.log_u_trace[g] = lu;
      if (p.acc_trace) p.acc_trace[g] = a ? 1 : 0;
      if (p.fwd_trace) p.fwd_trace[g] = fwd;
      if (p.rev_trace) p.rev_trace[g] = rev;
      if (p.e_new_trace) p.e_new_trace[g] = e_new;
      n_accept += a ? 1u : 0u;
    }
    if (a) {
      e_old = e_new;
      x1[0] = x2[0];
      x1[1] = x2[1];
    }
  }
  if (live && sub == 0) {
    p.x[chain * DX] = x1[0];
    p.x[chain * DX + 1] = x1[1];
    p.E[chain] = e_old;
  }
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

// Knot table of one spline from its raw parameters: the arithmetic of rqs_device.cuh `find_bin` / `load_slopes`
// (softmax over float32 exponentials accumulated in float64; knot k = bin_min + (scale E_k / total + 1e-2 k)).
__global__ void rqs_knot_table_kernel(const float* __restrict__ raw_w, const float* __restrict__ raw_h,
                                      const float* __restrict__ raw_s, int n, int K, float bin_min, float scale,
                                      double* __restrict__ tables) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n) return;
  const int sp = i >> 1, which = i & 1;  // one thread per (spline, widths | heights)
  const float* raw = (which ? raw_h : raw_w) + (int64_t)sp * K;
  double* tb = tables + (int64_t)sp * table_stride(K);
  double* kn = tb + which * (K + 1);
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, raw[k]);
  double tot = 0.0;
  for (int k = 0; k < K; ++k) tot += (double)expf(raw[k] - mx);
  const double c = (double)scale * rqsdev::recip_pos(tot), bm = (double)bin_min;
  double E = 0.0;
  for (int k = 0; k <= K; ++k) {
    kn[k] = bm + fma(c, E, 1e-2 * (double)k);
    if (k < K) E += (double)expf(raw[k] - mx);
  }
  if (which == 0) {
    float* dk = reinterpret_cast<float*>(tb + 2 * (K + 1));
    const float* rs = raw_s + (int64_t)sp * (K - 1);
    dk[0] = 1.f;
    dk[K] = 1.f;
    for (int k = 1; k < K; ++k) dk[k] = softplus_tf(rs[k - 1]) + 1e-2f;
    if (((K + 1) & 1) != 0) dk[K + 1] = 0.f;
  }
}

}  // namespace
}  // namespace vms

using namespace vms;

extern "C" {

int64_t vms_rqs_knot_table_doubles(int n_bins) { return n_bins >= 1 ? (int64_t)table_stride(n_bins) : 0; }

vms_status vms_rqs_knot_table(const float* raw_w, const float* raw_h, const float* raw_s, int n_splines, int n_bins,
                              float range_min, float range_max, double* tables, vms_stream stream) {
  VMS_REQUIRE(n_splines >= 0 && n_bins >= 2, VMS_ERR_SHAPE, "rqs_knot_table: need n_bins >= 2");
  VMS_REQUIRE(n_splines == 0 || (raw_w && raw_h && raw_s && tables), VMS_ERR_INVALID_ARG, "rqs_knot_table: NULL pointer");
  const float scale = (float)((double)range_max - (double)range_min - (double)n_bins * 1e-2);
  VMS_REQUIRE(scale > 0.f, VMS_ERR_INVALID_ARG, "bin_range too narrow for %d bins", n_bins);
  if (n_splines == 0) return VMS_OK;
  rqs_knot_table_kernel<<<(unsigned)((2 * n_splines + 63) / 64), 64, 0, as_stream(stream)>>>(raw_w, raw_h, raw_s, n_splines,
                                                                                             n_bins, range_min, scale, tables);
  VMS_LAUNCH_CHECK("rqs_knot_table_kernel");
  return VMS_OK;
}

int vms_mc_nb_supported(const vms_mc_nb_model* m) {
  if (!m) return 0;
  if (m->dx != DX || m->dz != DZ) return 0;
  if (m->enc_hidden < 1 || m->dec_hidden < 1 || m->enc_hidden > 1024 || m->dec_hidden > 1024) return 0;
  if (m->made_hidden[0] < 1 || m->made_hidden[0] > 16 || m->made_hidden[2] < 1 || m->made_hidden[2] > 16) return 0;
  if (m->made_hidden[1] < 1 || m->made_hidden[1] > 512) return 0;
  if (m->made_act != VMS_ACT_NONE && m->made_act != VMS_ACT_RELU && m->made_act != VMS_ACT_TANH) return 0;
  if (m->n_blocks < 1 || m->n_blocks > 16 || m->n_bins < 2 || m->n_bins > 256) return 0;
  if (m->n_comp < 1 || m->n_comp > 16) return 0;
  if (m->made_first_dof < -1 || m->made_first_dof > 1) return 0;
  return 1;
}

vms_status vms_mc_nb_run(const vms_mc_nb_model* model, float* x, float* E, int energies_valid, const float* noise,
                         unsigned long long seed, unsigned long long step0, const double* log_u, const vms_pcg64_stream* rng,
                         int64_t chain0, int64_t B, int n_steps, unsigned long long* n_acc, unsigned long long* n_uncertain,
                         uint8_t* acc_trace, float* fwd_trace, float* rev_trace, float* e_new_trace, double* log_u_trace,
                         vms_stream stream) {
  VMS_RANGE("vms_mc_nb_run");
  VMS_REQUIRE(model != nullptr, VMS_ERR_INVALID_ARG, "mc_nb_run: NULL model");
  VMS_REQUIRE(vms_mc_nb_supported(model), VMS_ERR_UNSUPPORTED,
              "mc_nb_run: built for the MC notebook's family (dx = 2, dz = 1, MADE hidden [<=16, <=512, <=16])");
  VMS_REQUIRE(B >= 0 && n_steps >= 0, VMS_ERR_SHAPE, "mc_nb_run: bad shape");
  VMS_REQUIRE((log_u != nullptr) != (rng != nullptr), VMS_ERR_INVALID_ARG, "mc_nb_run: exactly one of log_u / rng");
  VMS_REQUIRE(B == 0 || (x && E && n_acc), VMS_ERR_INVALID_ARG, "mc_nb_run: NULL pointer");
  const vms_mc_nb_model& m = *model;
  VMS_REQUIRE(m.enc_W0 && m.enc_b0 && m.enc_W1 && m.enc_b1 && m.dec_W0 && m.dec_b0 && m.dec_W1 && m.dec_b1 && m.tables &&
                  m.gmm_log_w && m.gmm_loc && m.gmm_scale, VMS_ERR_INVALID_ARG, "mc_nb_run: NULL model pointer");
  for (int k = 0; k < 4; ++k)
    VMS_REQUIRE(m.made_W[k] && m.made_b[k] && m.made_Wc[k], VMS_ERR_INVALID_ARG, "mc_nb_run: NULL MADE pointer");
  VMS_REQUIRE(((uintptr_t)noise & 15) == 0, VMS_ERR_INVALID_ARG, "mc_nb_run: noise must be 16-byte aligned");
  if (B == 0 || n_steps == 0) return VMS_OK;
  NbParams p = {};
  p.m = m;
  if (rng) {
    p.use_pcg = 1;
    p.s0_hi = rng->state_hi; p.s0_lo = rng->state_lo; p.inc_hi = rng->inc_hi; p.inc_lo = rng->inc_lo;
    p.jm_hi = rng->stride_mul_hi; p.jm_lo = rng->stride_mul_lo; p.ja_hi = rng->stride_add_hi; p.ja_lo = rng->stride_add_lo;
  }
  p.chain0 = chain0;
  p.B = B; p.n_steps = n_steps; p.x = x; p.E = E; p.energies_valid = energies_valid;
  p.noise = noise; p.seed = seed; p.step0 = step0; p.log_u = log_u;
  p.n_acc = n_acc; p.n_uncertain = n_uncertain;
  p.acc_trace = acc_trace; p.fwd_trace = fwd_trace; p.rev_trace = rev_trace; p.e_new_trace = e_new_trace;
  p.log_u_trace = log_u_trace;
  // outer MADE widths as staged: the notebook's [10, ., 10] exactly, anything else zero-padded to 12 or 16
  const int wide = m.made_hidden[0] > m.made_hidden[2] ? m.made_hidden[0] : m.made_hidden[2];
  const int HP = (m.made_hidden[0] == 10 && m.made_hidden[2] == 10) ? 10 : (wide <= 12 ? 12 : 16);
  const Layout L = make_layout(m, HP, HP);
  const size_t smem = (size_t)L.total * sizeof(float);
  VMS_REQUIRE(smem <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "mc_nb_run: model too large for shared memory");
  cudaStream_t st = as_stream(stream);
  // lanes per chain: one when the chains alone fill the GPU, two / four when they are scarce (identical results, see `combine`)
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // (B200, 148 SMs, measured: 1 lane wins from 16,384 chains up, 2 lanes around 8,192, 4 lanes below ~6,000)
  int tpc = B >= (int64_t)sms * 110 ? 1 : (B >= (int64_t)sms * 40 ? 2 : 4);
  if (const char* e = getenv("VMS_NB_TPC")) {  // cross-checks: force a lane count
    const int t = atoi(e);
    if (t == 1 || t == 2 || t == 4) tpc = t;
  }
  const unsigned grid = (unsigned)((B * tpc + CT - 1) / CT);
  int minb = 2;
  if (const char* e = getenv("VMS_NB_OCC")) minb = atoi(e);  // development aid: register budget of the one-lane kernel
#define VMS_NB_LAUNCH(T, H, M)                                                                                        \
  do {                                                                                                                \
    VMS_CUDA(cudaFuncSetAttribute(mc_nb_kernel<T, H, H, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mc_nb_kernel<T, H, H, M><<<grid, CT, smem, st>>>(p);                                                              \
  } while (0)
#define VMS_NB_SHAPE(T, M)                     \
  do {                                         \
    if (HP == 10) VMS_NB_LAUNCH(T, 10, M);     \
    else if (HP == 12) VMS_NB_LAUNCH(T, 12, M); \
    else VMS_NB_LAUNCH(T, 16, M);              \
  } while (0)
  if (tpc == 1) {
    if (minb == 3) VMS_NB_SHAPE(1, 3);
    else if (minb == 4) VMS_NB_SHAPE(1, 4);
    else VMS_NB_SHAPE(1, 2);
  } else if (tpc == 2) {
    VMS_NB_SHAPE(2, 1);
  } else {
    VMS_NB_SHAPE(4, 1);
  }
#undef VMS_NB_SHAPE
#undef VMS_NB_LAUNCH
  VMS_LAUNCH_CHECK("mc_nb_kernel");
  return VMS_OK;
}

}  // extern "C"
