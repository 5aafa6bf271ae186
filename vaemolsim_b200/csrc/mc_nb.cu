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

constexpr int CT = 128;   // threads per CTA = 32 chains
constexpr int TPC = 4;    // lanes per chain
constexpr int MROW = 12;  // floats per FCDeepNN hidden unit: enc W0[0..1][j] b0[j] W1[j][0..1] | dec W0[0][j] b0[j] W1[j][0..3] | pad
constexpr int DX = 2, DZ = 1, NN = 2 * DZ + DX;

struct NbParams {
  vms_mc_nb_model m;
  int P;  // padded width of the outer MADE hidden layers (multiple of 4)
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

// shared-memory image (offsets in floats; the knot tables come first for their 8-byte alignment)
struct Layout {
  int tables, mlp, eb1, db1, l0, mid, tail, gmm, total;
  int H, RS;
};
__host__ __device__ inline Layout make_layout(const vms_mc_nb_model& m, int P) {
  Layout L;
  L.H = m.enc_hidden > m.dec_hidden ? m.enc_hidden : m.dec_hidden;
  L.RS = 2 * P + 4;
  int o = 0;
  L.tables = o; o = (o + 2 * m.n_blocks * table_stride(m.n_bins) + 3) & ~3;  // every region starts on a 16-byte boundary
  L.mlp = o;    o += L.H * MROW;
  L.eb1 = o;    o += 4;
  L.db1 = o;    o += 4;
  L.l0 = o;     o += 4 * P;                 // per unit i: W0[0][i], W0[1][i], Wc0[i], b0[i]
  L.mid = o;    o += m.made_hidden[1] * L.RS;
  L.tail = o;   o += 2 * P + 4 * P + 8;     // b2 [P] | Wc2 [P] | W3 [P][4] | b3 [4] | Wc3 [4]
  L.gmm = o;    o += 5 * m.n_comp;          // log_w [n] | loc [n][2] | scale [n][2]
  L.total = (o + 3) & ~3;
  return L;
}

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == VMS_ACT_RELU) return fmaxf(v, 0.f);
  if (act == VMS_ACT_TANH) return tanhf(v);
  return v;
}

// One block of the MAF prior from its knot table (the formulas of rqs_device.cuh `octet_apply`).
__device__ __forceinline__ void spline_apply(const double* __restrict__ tb, int K, float bin_min, float v, bool inv, float& out,
                                             float& ldj) {
  const double* kx = tb;
  const double* ky = tb + (K + 1);
  const float* dks = reinterpret_cast<const float*>(tb + 2 * (K + 1));
  const double* ks = inv ? ky : kx;
  const double vd = (double)v;
  out = v;
  ldj = 0.f;
  // bin k covers [knot k, knot k+1); the range edges themselves are outside (TFP: identity)
  if (!(vd > (double)bin_min && vd < ks[K])) return;
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
  if (!inv) {
    const float num = b.hk * (sk * rr * rr + dk * u);
    out = (float)(b.lo_y + (double)(num / den));
  } else {
    out = (float)(b.lo_x + (double)(rr * b.wk));
  }
  const float Pq = dk1 * rr * rr + 2.f * sk * u + dk * omr * omr;
  ldj = logf((sk * sk) * Pq / (den * den));
  if (inv) ldj = -ldj;
}

// encoder(xe) and decoder-mapping(zd) side by side on this lane's hidden units j = sub, sub + 4, ...
__device__ __forceinline__ void mlp_pair(const float* __restrict__ rows, int H, int sub, float x0, float x1, float zd,
                                         float (&pe)[2], float (&pd)[4]) {
  pe[0] = pe[1] = 0.f;
  pd[0] = pd[1] = pd[2] = pd[3] = 0.f;
#pragma unroll 2
  for (int j = sub; j < H; j += TPC) {
    const float4* row = reinterpret_cast<const float4*>(rows + j * MROW);
    const float4 a = row[0], b = row[1], c = row[2];
    float he = fmaf(x1, a.y, x0 * a.x);
    he = fmaxf(he + a.z, 0.f);
    pe[0] = fmaf(he, a.w, pe[0]);
    pe[1] = fmaf(he, b.x, pe[1]);
    const float hd = fmaxf(fmaf(zd, b.y, b.z), 0.f);
    pd[0] = fmaf(hd, b.w, pd[0]);
    pd[1] = fmaf(hd, c.x, pd[1]);
    pd[2] = fmaf(hd, c.y, pd[2]);
    pd[3] = fmaf(hd, c.z, pd[3]);
  }
  pe[0] = quad_sum(pe[0]); pe[1] = quad_sum(pe[1]);
#pragma unroll
  for (int n = 0; n < 4; ++n) pd[n] = quad_sum(pd[n]);
}

// AutoregressiveNetwork(params = 2, event 2, conditional 1, hidden [h0, h1, h2]) on (s0, s1 | cond) -> out [dof][param]
template <int P>
__device__ __forceinline__ void made_pass(const float* __restrict__ sm, const Layout& L, int h1, int act, int sub, float s0,
                                          float s1, float cond, float (&out)[4]) {
  float a0[P], a2[P];
  const float4* l0 = reinterpret_cast<const float4*>(sm + L.l0);
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const float4 w = l0[i];
    a0[i] = act_fn(fmaf(cond, w.z, fmaf(s1, w.y, s0 * w.x)) + w.w, act);  // (padding units: act(0) = 0 for every supported act)
    a2[i] = 0.f;
  }
#pragma unroll 1
  for (int j = sub; j < h1; j += TPC) {
    const float4* row = reinterpret_cast<const float4*>(sm + L.mid + j * L.RS);
    const float4 t = row[2 * (P / 4)];  // b1[j], Wc1[j]
    float h = cond * t.y;
#pragma unroll
    for (int q = 0; q < P / 4; ++q) {
      const float4 w = row[q];
      h = fmaf(a0[4 * q], w.x, h); h = fmaf(a0[4 * q + 1], w.y, h); h = fmaf(a0[4 * q + 2], w.z, h); h = fmaf(a0[4 * q + 3], w.w, h);
    }
    h = act_fn(h + t.x, act);
#pragma unroll
    for (int q = 0; q < P / 4; ++q) {
      const float4 w = row[P / 4 + q];
      a2[4 * q] = fmaf(h, w.x, a2[4 * q]); a2[4 * q + 1] = fmaf(h, w.y, a2[4 * q + 1]);
      a2[4 * q + 2] = fmaf(h, w.z, a2[4 * q + 2]); a2[4 * q + 3] = fmaf(h, w.w, a2[4 * q + 3]);
    }
  }
  const float* tl = sm + L.tail;
  const float4* b3 = reinterpret_cast<const float4*>(tl + 6 * P);
  const float4 bb = b3[0], wc = b3[1];
  float o0 = fmaf(cond, wc.x, bb.x), o1 = fmaf(cond, wc.y, bb.y), o2 = fmaf(cond, wc.z, bb.z), o3 = fmaf(cond, wc.w, bb.w);
  const float4* W3 = reinterpret_cast<const float4*>(tl + 2 * P);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const float v = act_fn(fmaf(cond, tl[P + k], quad_sum(a2[k])) + tl[k], act);
    const float4 w = W3[k];
    o0 = fmaf(v, w.x, o0); o1 = fmaf(v, w.y, o1); o2 = fmaf(v, w.z, o2); o3 = fmaf(v, w.w, o3);
  }
  out[0] = o0; out[1] = o1; out[2] = o2; out[3] = o3;
}

// Mixture log-density as tfp evaluates it in float32 (mcmc.cu energy_gmm_kernel)
__device__ __forceinline__ float gmm_component(const float* __restrict__ g, int n, int k, float x0, float x1) {
  const float* loc = g + n;
  const float* sc = g + 3 * n;
  float s = 0.f;
  s += normal_lp(x0, loc[2 * k], sc[2 * k]);
  s += normal_lp(x1, loc[2 * k + 1], sc[2 * k + 1]);
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

template <int P, int MINB>
__global__ void __launch_bounds__(CT, MINB) mc_nb_kernel(const NbParams p) {
  extern __shared__ __align__(16) float sm[];
  const vms_mc_nb_model& m = p.m;
  const Layout L = make_layout(m, P);
  const int tid = threadIdx.x;
  const int K = m.n_bins, nb = m.n_blocks, TS = table_stride(K);
  const int h0 = m.made_hidden[0], h1 = m.made_hidden[1], h2 = m.made_hidden[2];
  // ---- stage the model
  {
    const float* src = reinterpret_cast<const float*>(m.tables);
    for (int e = tid; e < 2 * nb * TS; e += CT) sm[L.tables + e] = __ldg(src + e);
    for (int e = tid; e < L.H * MROW; e += CT) {
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
    for (int e = tid; e < 4 * P; e += CT) {
      const int i = e >> 2, c = e & 3;
      float v = 0.f;
      if (i < h0) v = c < 2 ? __ldg(m.made_W[0] + c * h0 + i) : (c == 2 ? __ldg(m.made_Wc[0] + i) : __ldg(m.made_b[0] + i));
      sm[L.l0 + e] = v;
    }
    for (int e = tid; e < h1 * L.RS; e += CT) {
      const int j = e / L.RS, c = e - j * L.RS;
      float v = 0.f;
      if (c < P) { if (c < h0) v = __ldg(m.made_W[1] + c * h1 + j); }
      else if (c < 2 * P) { if (c - P < h2) v = __ldg(m.made_W[2] + j * h2 + (c - P)); }
      else if (c == 2 * P) v = __ldg(m.made_b[1] + j);
      else if (c == 2 * P + 1) v = __ldg(m.made_Wc[1] + j);
      sm[L.mid + e] = v;
    }
    for (int e = tid; e < 6 * P + 8; e += CT) {
      float v = 0.f;
      if (e < P) { if (e < h2) v = __ldg(m.made_b[2] + e); }
      else if (e < 2 * P) { if (e - P < h2) v = __ldg(m.made_Wc[2] + (e - P)); }
      else if (e < 6 * P) { const int k = (e - 2 * P) >> 2, n = (e - 2 * P) & 3; if (k < h2) v = __ldg(m.made_W[3] + k * 4 + n); }
      else if (e < 6 * P + 4) v = __ldg(m.made_b[3] + (e - 6 * P));
      else v = __ldg(m.made_Wc[3] + (e - 6 * P - 4));
      sm[L.tail + e] = v;
    }
    const int n = m.n_comp;
    for (int e = tid; e < 5 * n; e += CT)
      sm[L.gmm + e] = e < n ? __ldg(m.gmm_log_w + e) : (e < 3 * n ? __ldg(m.gmm_loc + (e - n)) : __ldg(m.gmm_scale + (e - 3 * n)));
  }
  __syncthreads();
  const double* tables = reinterpret_cast<const double*>(sm + L.tables);
  const float* rows = sm + L.mlp;
  const float* eb1 = sm + L.eb1;
  const float* db1 = sm + L.db1;
  const float* gmm = sm + L.gmm;
  const int act = m.made_act;

  const int64_t chain = (int64_t)blockIdx.x * (CT / TPC) + (tid / TPC);
  const int sub = tid & (TPC - 1);
  const bool live = chain < p.B;
  const int64_t cc = live ? chain : p.B - 1;  // idle lanes shadow the last chain (full-warp shuffles), writes predicated off
  float x1[DX];
  x1[0] = __ldg(p.x + cc * DX);
  x1[1] = __ldg(p.x + cc * DX + 1);
  float e_old = p.energies_valid ? p.E[cc] : gmm_energy(gmm, m.n_comp, x1[0], x1[1]);
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
    // ---- z2 ~ prior: base noise through the chain's blocks in sampling direction; log p(z2) = log N(eps) - sum fldj
    float z2 = nz[1];
    float lz2 = normal_lp(nz[1], 0.f, 1.f);
    {
      float fl = 0.f;
#pragma unroll 1
      for (int b = 0; b < nb; ++b) {
        float y, l;
        spline_apply(tables + b * TS, K, m.range_min, z2, false, y, l);
        z2 = y;
        fl += l;
      }
      lz2 -= fl;
    }
    // ---- encoder(x1) || decoder-mapping(z2)
    float pe[2], pd[4];
    mlp_pair(rows, L.H, sub, x1[0], x1[1], z2, pe, pd);
    const float loc1 = pe[0] + eb1[0], sc1 = softplus_tf(pe[1] + eb1[1]);
    const float z1 = __fadd_rn(__fmul_rn(nz[0], sc1), loc1);
    const float lq1 = normal_lp(z1, loc1, sc1);
    // ---- x2 ~ decoder(z2): tfp Autoregressive sampling, sample0 = ones, D + 1 passes with the same noise, then log_prob
    float in2[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) in2[n] = pd[n] + db1[n];
    float s0 = 1.f, s1 = 1.f, mo[4];
    if (m.made_first_dof >= 0) {
      // The masks make the first dof's parameters a function of the conditional input alone (its output columns of the last
      // kernel are zero: fmaf(v, 0, o) == o) and hide the second dof from every hidden unit (its input row is zero), so
      // passes 2, 3 and the log_prob pass of tfp's procedure see bit-identical hidden activations: one pass gives them all.
      const bool f = m.made_first_dof != 0;
      const float4 bb = reinterpret_cast<const float4*>(sm + L.tail + 6 * P)[0], wc = reinterpret_cast<const float4*>(sm + L.tail + 6 * P)[1];
      const float lf = (f ? in2[2] : in2[0]) + fmaf(z2, f ? wc.z : wc.x, f ? bb.z : bb.x);
      const float cf = softplus_tf((f ? in2[3] : in2[1]) + fmaf(z2, f ? wc.w : wc.y, f ? bb.w : bb.y)) + VMS_EPS32;
      const float sf = __fadd_rn(__fmul_rn(f ? nz[3] : nz[2], cf), lf);
      if (f) s1 = sf; else s0 = sf;
      made_pass<P>(sm, L, h1, act, sub, s0, s1, z2, mo);
      const float lo_ = f ? in2[0] + mo[0] : in2[2] + mo[2];
      const float co = softplus_tf(f ? in2[1] + mo[1] : in2[3] + mo[3]) + VMS_EPS32;
      const float so = __fadd_rn(__fmul_rn(f ? nz[2] : nz[3], co), lo_);
      if (f) s0 = so; else s1 = so;
    } else {
#pragma unroll 1
      for (int pass = 0; pass < DX + 1; ++pass) {
        made_pass<P>(sm, L, h1, act, sub, s0, s1, z2, mo);
        const float l0 = in2[0] + mo[0], c0 = softplus_tf(in2[1] + mo[1]) + VMS_EPS32;
        const float l1 = in2[2] + mo[2], c1 = softplus_tf(in2[3] + mo[3]) + VMS_EPS32;
        s0 = __fadd_rn(__fmul_rn(nz[2], c0), l0);
        s1 = __fadd_rn(__fmul_rn(nz[3], c1), l1);
      }
      made_pass<P>(sm, L, h1, act, sub, s0, s1, z2, mo);
    }
    const float x2[DX] = {s0, s1};
    float lx2 = 0.f;
    lx2 += normal_lp(x2[0], in2[0] + mo[0], softplus_tf(in2[1] + mo[1]) + VMS_EPS32);
    lx2 += normal_lp(x2[1], in2[2] + mo[2], softplus_tf(in2[3] + mo[3]) + VMS_EPS32);
    const float e_new = gmm_energy(gmm, m.n_comp, x2[0], x2[1]);
    // ---- reverse move: encoder(x2) || decoder-mapping(z1); prior.log_prob(z1); decoder(z1).log_prob(x1)
    float pe2[2], pd1[4];
    mlp_pair(rows, L.H, sub, x2[0], x2[1], z1, pe2, pd1);
    const float lq2 = normal_lp(z2, pe2[0] + eb1[0], softplus_tf(pe2[1] + eb1[1]));
    float lz1;
    {
      float v = z1, il = 0.f;
#pragma unroll 1
      for (int b = nb - 1; b >= 0; --b) {
        float y, l;
        spline_apply(tables + b * TS, K, m.range_min, v, true, y, l);
        v = y;
        il += l;
      }
      lz1 = normal_lp(v, 0.f, 1.f) + il;
    }
    float in1[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) in1[n] = pd1[n] + db1[n];
    made_pass<P>(sm, L, h1, act, sub, x1[0], x1[1], z1, mo);
    float lx1 = 0.f;
    lx1 += normal_lp(x1[0], in1[0] + mo[0], softplus_tf(in1[1] + mo[1]) + VMS_EPS32);
    lx1 += normal_lp(x1[1], in1[2] + mo[2], softplus_tf(in1[3] + mo[3]) + VMS_EPS32);
    // ---- accept / reject (mcmc.py:103, :109, :116-120), NumPy's float32 evaluation for a float32 energy
    const float fwd = __fadd_rn(__fadd_rn(lq1, lz2), lx2);
    const float rev = __fadd_rn(__fadd_rn(lq2, lz1), lx1);
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
  const double c = (double)scale / tot, bm = (double)bin_min;
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
  const int wide = m.made_hidden[0] > m.made_hidden[2] ? m.made_hidden[0] : m.made_hidden[2];
  p.P = wide <= 12 ? 12 : 16;
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
  const Layout L = make_layout(m, p.P);
  const size_t smem = (size_t)L.total * sizeof(float);
  VMS_REQUIRE(smem <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "mc_nb_run: model too large for shared memory");
  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)((B + CT / TPC - 1) / (CT / TPC));
  int occ = 2;
  if (const char* e = getenv("VMS_NB_OCC")) occ = atoi(e);  // development aid: resident CTAs per SM the register budget targets
#define VMS_NB_LAUNCH(PP, OO)                                                                                        \
  do {                                                                                                               \
    VMS_CUDA(cudaFuncSetAttribute(mc_nb_kernel<PP, OO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    mc_nb_kernel<PP, OO><<<grid, CT, smem, st>>>(p);                                                                 \
  } while (0)
  if (p.P == 12) {
    if (occ <= 1) VMS_NB_LAUNCH(12, 1);
    else if (occ == 2) VMS_NB_LAUNCH(12, 2);
    else if (occ == 3) VMS_NB_LAUNCH(12, 3);
    else VMS_NB_LAUNCH(12, 4);
  } else {
    VMS_NB_LAUNCH(16, 2);
  }
#undef VMS_NB_LAUNCH
  VMS_LAUNCH_CHECK("mc_nb_kernel");
  return VMS_OK;
}

}  // extern "C"
