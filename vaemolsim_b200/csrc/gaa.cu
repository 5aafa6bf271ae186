// gaa.cu -- rank-2 geometric-algebra vector attention over a selected point cloud (mappings.py:480-688: AttentionBlock /
// ParticleEmbedding; arithmetic of geometric_algebra_attention.VectorAttention(rank=2, merge_fun='concat',
// join_fun='concat') and Keras LayerNormalization, restated in oracle/gaa.py -- parity unpinned, the package is not vendored).
//
// Pair tensors are laid out [B, n(i), n(j), .]: entry (i, j) belongs to the geometric product r_j * r_i, output row i sums
// over j (reduce = 0) or the cloud sums over all (i, j) (reduce = 1).
//
// Two implementations:
//   * op-by-op kernels the host tape composes with vms_dense_forward / _backward (training, any width):
//       vms_gaa_pair_invariants, vms_layernorm_forward / _backward, vms_gaa_pair_merge / _backward,
//       vms_gaa_attend / _backward -- pair tensors live in HBM, every reduction in a fixed order;
//   * vms_gaa_attention_forward: ONE kernel per attention layer (inference / sampling): a thread owns a pair from its two
//     invariants to its score, weights are shared-memory broadcasts, the softmax is accumulated online (running max / sum /
//     weighted value per thread, merged per output row in a fixed order) -- no pair tensor ever reaches HBM.
#include "common.cuh"
#include "dense.cuh"
#include <stdlib.h>

namespace vms {
namespace {

constexpr float kMaskedScore = -1e9f;  // geometric_algebra_attention: masked logits

__device__ __forceinline__ float act_apply(float x, int act) {
  return act == VMS_ACT_RELU ? fmaxf(x, 0.f) : (act == VMS_ACT_TANH ? tanhf(x) : x);
}
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  return act == VMS_ACT_RELU ? (y > 0.f ? 1.f : 0.f) : (act == VMS_ACT_TANH ? 1.f - y * y : 1.f);
}

// (r_j . r_i, |r_j ^ r_i|) with TF's op order: nine separate products, then sums (no contraction)
__device__ __forceinline__ void pair_invariants(const float* a, const float* b, float& dot, float& nrm) {
  dot = __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
  const float xy = __fsub_rn(__fmul_rn(a[0], b[1]), __fmul_rn(a[1], b[0]));
  const float xz = __fsub_rn(__fmul_rn(a[0], b[2]), __fmul_rn(a[2], b[0]));
  const float yz = __fsub_rn(__fmul_rn(a[1], b[2]), __fmul_rn(a[2], b[1]));
  nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(xy, xy), __fmul_rn(xz, xz)), __fmul_rn(yz, yz)));
}

__global__ void gaa_pair_inv_kernel(const float* __restrict__ coords, int64_t total, int n, float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int j = (int)(p % n);
  const int64_t bi = p / n;
  const int i = (int)(bi % n);
  const int64_t b = bi / n;
  float dot, nrm;
  pair_invariants(coords + (b * n + j) * 3, coords + (b * n + i) * 3, dot, nrm);
  reinterpret_cast<float2*>(out)[p] = make_float2(dot, nrm);
}

// ------------------------------------------------------------------------------------------------- layer normalisation
// eight lanes per row: lane l owns columns l, l + 8, ...; row sums = ascending per lane, then xor 4, 2, 1
constexpr int LNG = 8;

__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t R, int H,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps, int act, float* __restrict__ y, int64_t ldy,
                                                            float* __restrict__ stats) {
  const int64_t row_raw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LNG;
  const bool ok = row_raw < R;
  const int64_t row = ok ? row_raw : R - 1;
  const int l = threadIdx.x & (LNG - 1);
  const float* xr = x + row * ldx;
  float s = 0.f;
  for (int k = l; k < H; k += LNG) s += xr[k];
  const float mean = group_sum(s) / (float)H;
  float q = 0.f;
  for (int k = l; k < H; k += LNG) {
    const float d = xr[k] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(group_sum(q) / (float)H + eps);
  if (!ok) return;
  float* yr = y + row * ldy;
  for (int k = l; k < H; k += LNG) yr[k] = act_apply((xr[k] - mean) * rstd * gamma[k] + beta[k], act);
  if (l == 0 && stats) reinterpret_cast<float2*>(stats)[row] = make_float2(mean, rstd);
}

// reverse mode: g_x += rstd (dxh - mean(dxh) - xhat mean(dxh xhat)), dxh = g_y act'(y) gamma; per-CTA column partials of
// d gamma = sum g_pre xhat and d beta = sum g_pre (E = columns per lane, registers)
template <int E>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t R, int H,
                                                            const float* __restrict__ gamma, const float* __restrict__ stats,
                                                            int act, const float* __restrict__ y, int64_t ldy,
                                                            const float* __restrict__ gy, int64_t ldgy,
                                                            float* __restrict__ gx, int64_t ldgx, float* __restrict__ part) {
  extern __shared__ float sm[];  // [32 groups][2 H]
  const int l = threadIdx.x & (LNG - 1);
  const int grp = threadIdx.x / LNG;
  const int groups = blockDim.x / LNG;
  float ag[E], ab[E];
#pragma unroll
  for (int e = 0; e < E; ++e) ag[e] = ab[e] = 0.f;
  const int64_t n_iter = (R + (int64_t)gridDim.x * groups - 1) / ((int64_t)gridDim.x * groups);
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row_raw = (it * gridDim.x + blockIdx.x) * groups + grp;
    const bool ok = row_raw < R;
    const int64_t row = ok ? row_raw : R - 1;
    const float2 st = reinterpret_cast<const float2*>(stats)[row];
    const float* xr = x + row * ldx;
    const float* yr = y + row * ldy;
    const float* gr = gy + row * ldgy;
    float dxh[E], xh[E];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int k = l + e * LNG;
      dxh[e] = xh[e] = 0.f;
      if (k < H) {
        xh[e] = (xr[k] - st.x) * st.y;
        const float gp = ok ? gr[k] * act_grad_from_out(yr[k], act) : 0.f;
        ab[e] += gp;
        ag[e] += gp * xh[e];
        dxh[e] = gp * gamma[k];
        s1 += dxh[e];
        s2 += dxh[e] * xh[e];
      }
    }
    s1 = group_sum(s1) / (float)H;
    s2 = group_sum(s2) / (float)H;
    if (ok && gx) {
      float* gxr = gx + row * ldgx;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int k = l + e * LNG;
        if (k < H) gxr[k] += st.y * (dxh[e] - s1 - xh[e] * s2);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int k = l + e * LNG;
    if (k < H) {
      sm[grp * 2 * H + k] = ag[e];
      sm[grp * 2 * H + H + k] = ab[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < groups; ++g) s += sm[g * 2 * H + c];
    part[(int64_t)blockIdx.x * 2 * H + c] = s;
  }
}

int ln_bwd_grid(int64_t R) {
  const int64_t want = (R + 31) / 32;
  const int64_t cap = 2 * (int64_t)sm_count();
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------- merged values of a pair
// out[b, i, j, :] = u[b, j, :] + w[b, i, :]
__global__ void gaa_pair_merge_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ w, int64_t ldw,
                                      int64_t total, int n, int D, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int d = (int)(e % D);
  const int64_t p = e / D;
  const int j = (int)(p % n);
  const int64_t bi = p / n;  // b * n + i
  const int64_t b = bi / n;
  out[e] = u[(b * n + j) * ldu + d] + w[bi * ldw + d];
}

// g_u[b, p, :] += sum_i g[b, i, p, :],  g_w[b, p, :] += sum_j g[b, p, j, :]   (ascending order)
__global__ void gaa_pair_merge_bwd_kernel(const float* __restrict__ g, int64_t total, int n, int D, float* __restrict__ gu,
                                          int64_t ldgu, float* __restrict__ gw, int64_t ldgw) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over [B, n, D]
  if (e >= total) return;
  const int d = (int)(e % D);
  const int64_t bp = e / D;
  const int p = (int)(bp % n);
  const int64_t b = bp / n;
  const float* gb = g + b * n * n * D;
  float su = 0.f, sw = 0.f;
  for (int q = 0; q < n; ++q) {
    su += gb[((int64_t)q * n + p) * D + d];
    sw += gb[((int64_t)p * n + q) * D + d];
  }
  gu[bp * ldgu + d] += su;
  gw[bp * ldgw + d] += sw;
}

// ------------------------------------------------------------------------------------------------------- attention
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float masked_score(const float* scores, const uint8_t* mask, int64_t b, int n, int i, int j,
                                              int64_t idx) {
  if (mask && !(mask[b * n + i] && mask[b * n + j])) return kMaskedScore;
  return scores[idx];
}

// reduce = 0: one warp per output row (b, i): softmax over j, out[b, i, :] = sum_j att_ij val_ij
__global__ void __launch_bounds__(256) gaa_attend_rows_kernel(const float* __restrict__ scores, const float* __restrict__ val,
                                                              const uint8_t* __restrict__ mask, int64_t rows, int n, int D,
                                                              float* __restrict__ out, float* __restrict__ att) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = w / n;
  const int i = (int)(w % n);
  float m = -INFINITY;
  for (int j = lane; j < n; j += 32) m = fmaxf(m, masked_score(scores, mask, b, n, i, j, w * n + j));
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float e = expf(masked_score(scores, mask, b, n, i, j, w * n + j) - m);
    att[w * n + j] = e;
    s += e;
  }
  s = warp_sum(s);
  for (int j = lane; j < n; j += 32) att[w * n + j] = att[w * n + j] / s;
  __syncwarp();
  for (int d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += att[w * n + j] * val[(w * n + j) * D + d];
    out[w * D + d] = acc;
  }
}

__global__ void __launch_bounds__(256) gaa_attend_rows_bwd_kernel(const float* __restrict__ att, const float* __restrict__ val,
                                                                  const uint8_t* __restrict__ mask, int64_t rows, int n, int D,
                                                                  const float* __restrict__ g_out, float* __restrict__ g_scores,
                                                                  float* __restrict__ g_val) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= rows) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = w / n;
  const int i = (int)(w % n);
  const float* go = g_out + w * D;
  float S = 0.f;
  for (int j = lane; j < n; j += 32) {
    float dot = 0.f;
    for (int d = 0; d < D; ++d) dot += go[d] * val[(w * n + j) * D + d];
    S += att[w * n + j] * dot;
  }
  S = warp_sum(S);
  for (int j = lane; j < n; j += 32) {
    const float a = att[w * n + j];
    float dot = 0.f;
    for (int d = 0; d < D; ++d) dot += go[d] * val[(w * n + j) * D + d];
    const bool masked = mask && !(mask[b * n + i] && mask[b * n + j]);
    if (g_scores && !masked) g_scores[w * n + j] += a * (dot - S);
    if (g_val)
      for (int d = 0; d < D; ++d) g_val[(w * n + j) * D + d] += a * go[d];
  }
}

// reduce = 1: one CTA per cloud: ONE softmax over all n^2 pairs, out[b, :] = sum_ij att_ij val_ij
constexpr int AT = 256;
__device__ float block_reduce(float v, float* red, bool is_max) {
  const int t = threadIdx.x;
  red[t] = v;
  __syncthreads();
  for (int o = AT / 2; o; o >>= 1) {
    if (t < o) red[t] = is_max ? fmaxf(red[t], red[t + o]) : red[t] + red[t + o];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(AT) gaa_attend_all_kernel(const float* __restrict__ scores, const float* __restrict__ val,
                                                            const uint8_t* __restrict__ mask, int n, int D,
                                                            float* __restrict__ out, float* __restrict__ att) {
  __shared__ float red[AT];
  __shared__ float part[AT / 32][33];
  const int64_t b = blockIdx.x;
  const int t = threadIdx.x;
  const int N2 = n * n;
  const int64_t base = b * N2;
  float m = -INFINITY;
  for (int p = t; p < N2; p += AT) m = fmaxf(m, masked_score(scores, mask, b, n, p / n, p % n, base + p));
  m = block_reduce(m, red, true);
  float s = 0.f;
  for (int p = t; p < N2; p += AT) {
    const float e = expf(masked_score(scores, mask, b, n, p / n, p % n, base + p) - m);
    att[base + p] = e;
    s += e;
  }
  s = block_reduce(s, red, false);
  for (int p = t; p < N2; p += AT) att[base + p] = att[base + p] / s;
  __syncthreads();
  const int c = t >> 5, lane = t & 31;
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int d = d0 + lane;
    float acc = 0.f;
    if (d < D)
      for (int p = c; p < N2; p += AT / 32) acc += att[base + p] * val[(base + p) * D + d];
    part[c][lane] = acc;
    __syncthreads();
    if (c == 0 && d < D) {
      float o = 0.f;
      for (int k = 0; k < AT / 32; ++k) o += part[k][lane];
      out[b * D + d] = o;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(AT) gaa_attend_all_bwd_kernel(const float* __restrict__ att, const float* __restrict__ val,
                                                                const uint8_t* __restrict__ mask, int n, int D,
                                                                const float* __restrict__ g_out, float* __restrict__ g_scores,
                                                                float* __restrict__ g_val) {
  __shared__ float red[AT];
  const int64_t b = blockIdx.x;
  const int t = threadIdx.x;
  const int N2 = n * n;
  const int64_t base = b * N2;
  const float* go = g_out + b * D;
  float S = 0.f;
  for (int p = t; p < N2; p += AT) {
    float dot = 0.f;
    for (int d = 0; d < D; ++d) dot += go[d] * val[(base + p) * D + d];
    S += att[base + p] * dot;
  }
  S = block_reduce(S, red, false);
  for (int p = t; p < N2; p += AT) {
    const float a = att[base + p];
    float dot = 0.f;
    for (int d = 0; d < D; ++d) dot += go[d] * val[(base + p) * D + d];
    const bool masked = mask && !(mask[b * n + p / n] && mask[b * n + p % n]);
    if (g_scores && !masked) g_scores[base + p] += a * (dot - S);
    if (g_val)
      for (int d = 0; d < D; ++d) g_val[(base + p) * D + d] += a * go[d];
  }
}

__global__ void gaa_zero_mask_kernel(const float* __restrict__ coords, int64_t total, uint8_t* __restrict__ mask) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const float* r = coords + p * 3;
  mask[p] = (r[0] != 0.f || r[1] != 0.f || r[2] != 0.f) ? 1 : 0;
}

// ===================================================================================== fused forward (one attention layer)
// One CTA per cloud (grid-stride).  A pack kernel lays the layer's weights out as ONE zero-padded image (HP / DP columns);
// the attention kernel reads it either from the CONSTANT bank (CW: every FFMA takes its weight as a c[bank][offset]
// operand -- no load instruction, no shared-memory wavefront; the image travels by cudaMemcpyToSymbolAsync on the stream)
// or from shared memory (16-byte broadcasts).  Shared memory also holds the coordinates, the mask, the per-particle halves
// of the merged values u = v M0, w = v M1, and the per-thread online-softmax states when they are merged.
// Thread t owns output row i = t / TPR and walks j = t % TPR, + TPR, ...:
//   inv(2) -> h = inv Wv1 + bv1 -> LayerNorm -> act -> iv = . Wv2 + bv2 -> joined = iv J1 + (u_j + w_i) J2
//   -> score = act(joined Ws1 + bs1) Ws2 + bs2  -> running (max, sum, sum of exp x joined).
template <int HP, int DP>
struct GaaImg {
  static constexpr int oWv1 = 0;                  // [2][HP]
  static constexpr int obv1 = oWv1 + 2 * HP;      // [HP]
  static constexpr int oLg = obv1 + HP;           // [HP]
  static constexpr int oLb = oLg + HP;            // [HP]
  static constexpr int oWv2 = oLb + HP;           // [HP][DP]
  static constexpr int obv2 = oWv2 + HP * DP;     // [DP]
  static constexpr int oJ1 = obv2 + DP;           // [DP][DP]
  static constexpr int oJ2 = oJ1 + DP * DP;       // [DP][DP]
  static constexpr int oWs1 = oJ2 + DP * DP;      // [DP][HP]
  static constexpr int obs1 = oWs1 + DP * HP;     // [HP]
  static constexpr int oWs2 = obs1 + HP;          // [HP]
  static constexpr int obs2 = oWs2 + HP;          // [4] (one value)
  static constexpr int nPair = obs2 + 4;          // what the pair loop reads
  static constexpr int oM0 = nPair;               // [DP][DP]
  static constexpr int oM1 = oM0 + DP * DP;       // [DP][DP]
  static constexpr int nW = oM1 + DP * DP;
};
constexpr int kGaaConstFloats = GaaImg<64, 32>::nPair;
__constant__ float c_gaa_w[kGaaConstFloats];

struct GaaFwdParams {
  const float* coords; const float* v; int64_t ldv; const uint8_t* mask;
  int64_t B; int n, D, H, reduce, act, tpr;
  const float *M0, *M1, *J1, *J2, *Ws1, *bs1, *Ws2, *bs2, *Wv1, *bv1, *lg, *lb, *Wv2, *bv2;
  float ln_eps;
  float* out;
  float* img;  // the packed weight image in global memory
};

template <int HP, int DP>
__global__ void gaa_pack_kernel(GaaFwdParams p) {
  using L = GaaImg<HP, DP>;
  float* W = p.img;
  const int D = p.D, H = p.H;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L::nW; k += gridDim.x * blockDim.x) {
    float val = 0.f;
    int o = k;
    if (o < L::obv1) { const int r = o / HP, c = o % HP; if (c < H) val = p.Wv1[r * H + c]; }
    else if (o < L::oLg) { o -= L::obv1; if (o < H) val = p.bv1[o]; }
    else if (o < L::oLb) { o -= L::oLg; if (o < H) val = p.lg[o]; }
    else if (o < L::oWv2) { o -= L::oLb; if (o < H) val = p.lb[o]; }
    else if (o < L::obv2) { o -= L::oWv2; const int r = o / DP, c = o % DP; if (r < H && c < D) val = p.Wv2[r * D + c]; }
    else if (o < L::oJ1) { o -= L::obv2; if (o < D) val = p.bv2[o]; }
    else if (o < L::oJ2) { o -= L::oJ1; const int r = o / DP, c = o % DP; if (r < D && c < D) val = p.J1[r * D + c]; }
    else if (o < L::oWs1) { o -= L::oJ2; const int r = o / DP, c = o % DP; if (r < D && c < D) val = p.J2[r * D + c]; }
    else if (o < L::obs1) { o -= L::oWs1; const int r = o / HP, c = o % HP; if (r < D && c < H) val = p.Ws1[r * H + c]; }
    else if (o < L::oWs2) { o -= L::obs1; if (o < H) val = p.bs1[o]; }
    else if (o < L::obs2) { o -= L::oWs2; if (o < H) val = p.Ws2[o]; }
    else if (o < L::oM0) { o -= L::obs2; if (o == 0) val = p.bs2[0]; }
    else if (o < L::oM1) { o -= L::oM0; const int r = o / DP, c = o % DP; if (r < D && c < D) val = p.M0[r * D + c]; }
    else { o -= L::oM1; const int r = o / DP, c = o % DP; if (r < D && c < D) val = p.M1[r * D + c]; }
    W[k] = val;
  }
}

template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  return ACT == VMS_ACT_RELU ? fmaxf(x, 0.f) : (ACT == VMS_ACT_TANH ? tanhf(x) : x);
}

template <int HP, int DP, int ACT, bool CW, int NP>
__global__ void __launch_bounds__(256) gaa_attention_fwd_kernel(GaaFwdParams p) {
  using L = GaaImg<HP, DP>;
  extern __shared__ __align__(16) float smem[];
  const int n = p.n, D = p.D, H = p.H;
  float* W = smem;                                  // !CW: the pair-loop part of the image
  float* s_r = smem + (CW ? 0 : L::nPair);          // [n][3] (+ pad to a multiple of 4)
  float* s_u = s_r + ((3 * n + 3) & ~3);            // [n][DP]
  float* s_w = s_u + n * DP;                        // [n][DP]
  float* s_v = s_w + n * DP;                        // [n][DP]
  float* s_st = s_v + n * DP;                       // [blockDim][DP + 2] online-softmax states
  uint8_t* s_m = reinterpret_cast<uint8_t*>(s_st + blockDim.x * (DP + 2));  // [n]
  const int t = threadIdx.x;
  if (!CW)
    for (int k = t; k < L::nPair; k += blockDim.x) W[k] = p.img[k];
#define WT(off) (CW ? c_gaa_w[(off)] : W[(off)])
  const float* gM0 = p.img + L::oM0;
  const float* gM1 = p.img + L::oM1;
  const float inv_H = 1.f / (float)H;
  const int tpr = p.tpr;
  const int rows_per_pass = blockDim.x / tpr;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    for (int k = t; k < 3 * n; k += blockDim.x) s_r[k] = p.coords[b * n * 3 + k];
    for (int k = t; k < n; k += blockDim.x) s_m[k] = p.mask ? p.mask[b * n + k] : 1;
    for (int k = t; k < n * DP; k += blockDim.x) {
      const int q = k / DP, d = k % DP;
      s_v[k] = d < D ? p.v[(b * n + q) * p.ldv + d] : 0.f;
    }
    __syncthreads();
    for (int k = t; k < n * DP; k += blockDim.x) {  // u = v M0, w = v M1 (contraction ascending, as the Dense kernels)
      const int q = k / DP, d = k % DP;
      float su = 0.f, sw = 0.f;
      for (int c = 0; c < D; ++c) {
        const float vv = s_v[q * DP + c];
        su = fmaf(vv, __ldg(gM0 + c * DP + d), su);
        sw = fmaf(vv, __ldg(gM1 + c * DP + d), sw);
      }
      s_u[k] = su;
      s_w[k] = sw;
    }
    __syncthreads();

    float red_m = -INFINITY, red_l = 0.f;  // reduce = 1: threads 0 .. DP-1 accumulate the cloud's state over the passes
    float red_o = 0.f;                     // (each keeps m / l redundantly and column t of the weighted sum)

    for (int i0 = 0; i0 < n; i0 += rows_per_pass) {
      const int i = i0 + t / tpr;
      const int sub = t % tpr;
      const bool active = (t / tpr) < rows_per_pass && i < n;
      float m = -INFINITY, l = 0.f;
      float acc[DP];
#pragma unroll
      for (int d = 0; d < DP; ++d) acc[d] = 0.f;
      if (active) {
        const float* ri = s_r + 3 * i;
        const bool mi = s_m[i];
        // NP pairs per trip: every weight load feeds NP FMAs (the shared-memory pipe, not the FMA pipe, bounds NP = 1)
        for (int j0 = sub; j0 < n; j0 += NP * tpr) {
          // the weights are loop-invariant loads: without this fence the compiler hoists all ~3,000 of them out of the pair
          // loop and spills them to local memory (12 KB of stack per thread)
          asm volatile("" ::: "memory");
          int jj[NP];
          bool ok[NP];
          float dot[NP], nrm[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            ok[q] = j0 + q * tpr < n;
            jj[q] = ok[q] ? j0 + q * tpr : j0;
            pair_invariants(s_r + 3 * jj[q], ri, dot[q], nrm[q]);
          }
          // value net, first layer + layer normalisation (two-pass moments over the H true columns)
          float h[NP][HP];
#pragma unroll
          for (int k = 0; k < HP; ++k) {
            const float w0 = WT(L::oWv1 + k), w1 = WT(L::oWv1 + HP + k), bb = WT(L::obv1 + k);
#pragma unroll
            for (int q = 0; q < NP; ++q) h[q][k] = fmaf(nrm[q], w1, fmaf(dot[q], w0, bb));
          }
          float mean[NP], rstd[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float mu = 0.f;
#pragma unroll
            for (int k = 0; k < HP; ++k) mu += (k < H) ? h[q][k] : 0.f;
            mu *= inv_H;
            float var = 0.f;
#pragma unroll
            for (int k = 0; k < HP; ++k) {
              const float dlt = h[q][k] - mu;
              var += (k < H) ? dlt * dlt : 0.f;
            }
            mean[q] = mu;
            rstd[q] = rsqrtf(var * inv_H + p.ln_eps);
          }
          float iv[NP][DP];
#pragma unroll
          for (int d = 0; d < DP; ++d) {
            const float bb = WT(L::obv2 + d);
#pragma unroll
            for (int q = 0; q < NP; ++q) iv[q][d] = bb;
          }
#pragma unroll
          for (int k = 0; k < HP; ++k) {
            const float lg = WT(L::oLg + k), lb = WT(L::oLb + k);  // padded columns: gamma = beta = 0
            float a[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) a[q] = act_t<ACT>((h[q][k] - mean[q]) * rstd[q] * lg + lb);
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              const float w = WT(L::oWv2 + k * DP + d);
#pragma unroll
              for (int q = 0; q < NP; ++q) iv[q][d] = fmaf(a[q], w, iv[q][d]);
            }
          }
          // joined = iv J1 + (u_j + w_i) J2
          float jn[NP][DP];
#pragma unroll
          for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int d = 0; d < DP; ++d) jn[q][d] = 0.f;
#pragma unroll
          for (int c = 0; c < DP; ++c) {
#pragma unroll
            for (int d = 0; d < DP; ++d) {
              const float w = WT(L::oJ1 + c * DP + d);
#pragma unroll
              for (int q = 0; q < NP; ++q) jn[q][d] = fmaf(iv[q][c], w, jn[q][d]);
            }
          }
          {
            float j2[NP][DP];
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
              for (int d = 0; d < DP; ++d) j2[q][d] = 0.f;
#pragma unroll
            for (int c = 0; c < DP; ++c) {
              float a[NP];
#pragma unroll
              for (int q = 0; q < NP; ++q) a[q] = s_u[jj[q] * DP + c] + s_w[i * DP + c];
#pragma unroll
              for (int d = 0; d < DP; ++d) {
                const float w = WT(L::oJ2 + c * DP + d);
#pragma unroll
                for (int q = 0; q < NP; ++q) j2[q][d] = fmaf(a[q], w, j2[q][d]);
              }
            }
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
              for (int d = 0; d < DP; ++d) jn[q][d] += j2[q][d];
          }
          // score net
          float sc[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) sc[q] = WT(L::obs2);
#pragma unroll
          for (int k0 = 0; k0 < HP; k0 += 4) {
            float hh[NP][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float bb = WT(L::obs1 + k0 + e);
#pragma unroll
              for (int q = 0; q < NP; ++q) hh[q][e] = bb;
            }
#pragma unroll
            for (int c = 0; c < DP; ++c) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float w = WT(L::oWs1 + c * HP + k0 + e);
#pragma unroll
                for (int q = 0; q < NP; ++q) hh[q][e] = fmaf(jn[q][c], w, hh[q][e]);
              }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float w = WT(L::oWs2 + k0 + e);
#pragma unroll
              for (int q = 0; q < NP; ++q) sc[q] = fmaf(act_t<ACT>(hh[q][e]), w, sc[q]);
            }
          }
          // online softmax, pair by pair in ascending j (the same order as NP = 1)
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            if (!ok[q]) continue;
            float s1 = sc[q];
            if (!(mi && s_m[jj[q]])) s1 = kMaskedScore;
            const float mn = fmaxf(m, s1);
            const float scale = expf(m - mn);  // first pair: exp(-inf) = 0
            const float e = expf(s1 - mn);
            l = l * scale + e;
#pragma unroll
            for (int d = 0; d < DP; ++d) acc[d] = acc[d] * scale + e * jn[q][d];
            m = mn;
          }
        }
      }
      // merge the states of the threads of a row (reduce = 0) or of the pass (reduce = 1) in ascending thread order
      float* st = s_st + t * (DP + 2);
      st[0] = m; st[1] = l;
#pragma unroll
      for (int d = 0; d < DP; ++d) st[2 + d] = acc[d];
      __syncthreads();
      if (!p.reduce) {
        if (active) {
          const int first = (t / tpr) * tpr;
          float M = -INFINITY;
          for (int q = 0; q < tpr; ++q) M = fmaxf(M, s_st[(first + q) * (DP + 2)]);
          float Ls = 0.f;
          for (int q = 0; q < tpr; ++q) {
            const float* sq = s_st + (first + q) * (DP + 2);
            if (sq[1] > 0.f) Ls += sq[1] * expf(sq[0] - M);
          }
          for (int d = sub; d < D; d += tpr) {
            float o = 0.f;
            for (int q = 0; q < tpr; ++q) {
              const float* sq = s_st + (first + q) * (DP + 2);
              if (sq[1] > 0.f) o += sq[2 + d] * expf(sq[0] - M);
            }
            p.out[(b * n + i) * D + d] = o / Ls;
          }
        }
      } else if (t < DP) {
        // threads 0 .. DP-1 fold this pass into the cloud state (each keeps m / l redundantly, column t of acc)
        const int nact = min(rows_per_pass, n - i0) * tpr;
        float M = red_m;
        for (int q = 0; q < nact; ++q) M = fmaxf(M, s_st[q * (DP + 2)]);
        const float sc0 = red_l > 0.f ? expf(red_m - M) : 0.f;
        float Ls = red_l * sc0, o = red_o * sc0;
        for (int q = 0; q < nact; ++q) {
          const float* sq = s_st + q * (DP + 2);
          if (sq[1] > 0.f) {
            const float f = expf(sq[0] - M);
            Ls += sq[1] * f;
            o += sq[2 + t] * f;
          }
        }
        red_m = M; red_l = Ls; red_o = o;
      }
      __syncthreads();
    }
    if (p.reduce && t < D) p.out[b * D + t] = red_o / red_l;
  }
#undef WT
}

// ---- the same kernel with PACKED float32 arithmetic (fma.rn.f32x2 = SASS FFMA2: two independent round-to-nearest FMAs per
// instruction on a 64-bit register pair; measured by vms_probe_ffma2: the same FLOP rate as FFMA with half the issue slots).
// Two pairs per trip; every per-pair vector lives as (column d, column d + 1) register pairs, the weight pair of two adjacent
// columns is one 8-byte half of a 16-byte shared-memory broadcast, the per-pair scalar multiplier is duplicated once per row
// of the weight matrix.  Same operations in the same order as the scalar kernels above (results agree to the last bits; the
// softmax update's multiply-adds may contract differently).
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

template <int HP, int DP, int ACT>
__global__ void __launch_bounds__(256) gaa_attention_fwd_x2_kernel(GaaFwdParams p) {
  using L = GaaImg<HP, DP>;
  constexpr int NP = 2, H2 = HP / 2, D2 = DP / 2;
  extern __shared__ __align__(16) float smem[];
  const int n = p.n, D = p.D, H = p.H;
  float* W = smem;
  float* s_r = smem + L::nPair;
  float* s_u = s_r + ((3 * n + 3) & ~3);
  float* s_w = s_u + n * DP;
  float* s_v = s_w + n * DP;
  float* s_st = s_v + n * DP;
  uint8_t* s_m = reinterpret_cast<uint8_t*>(s_st + blockDim.x * (DP + 2));
  const int t = threadIdx.x;
  for (int k = t; k < L::nPair; k += blockDim.x) W[k] = p.img[k];
  const f2_t* W2 = reinterpret_cast<const f2_t*>(W);  // (every image offset is a multiple of four floats)
  const float* gM0 = p.img + L::oM0;
  const float* gM1 = p.img + L::oM1;
  const float inv_H = 1.f / (float)H;
  const int tpr = p.tpr;
  const int rows_per_pass = blockDim.x / tpr;

  for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
    __syncthreads();
    for (int k = t; k < 3 * n; k += blockDim.x) s_r[k] = p.coords[b * n * 3 + k];
    for (int k = t; k < n; k += blockDim.x) s_m[k] = p.mask ? p.mask[b * n + k] : 1;
    for (int k = t; k < n * DP; k += blockDim.x) {
      const int q = k / DP, d = k % DP;
      s_v[k] = d < D ? p.v[(b * n + q) * p.ldv + d] : 0.f;
    }
    __syncthreads();
    for (int k = t; k < n * DP; k += blockDim.x) {
      const int q = k / DP, d = k % DP;
      float su = 0.f, sw = 0.f;
      for (int c = 0; c < D; ++c) {
        const float vv = s_v[q * DP + c];
        su = fmaf(vv, __ldg(gM0 + c * DP + d), su);
        sw = fmaf(vv, __ldg(gM1 + c * DP + d), sw);
      }
      s_u[k] = su;
      s_w[k] = sw;
    }
    __syncthreads();

    float red_m = -INFINITY, red_l = 0.f, red_o = 0.f;
    for (int i0 = 0; i0 < n; i0 += rows_per_pass) {
      const int i = i0 + t / tpr;
      const int sub = t % tpr;
      const bool active = (t / tpr) < rows_per_pass && i < n;
      float m = -INFINITY, l = 0.f;
      f2_t acc[D2];
#pragma unroll
      for (int d = 0; d < D2; ++d) acc[d] = 0ull;
      if (active) {
        const float* ri = s_r + 3 * i;
        const bool mi = s_m[i];
        const f2_t* wi2 = reinterpret_cast<const f2_t*>(s_w + i * DP);
        for (int j0 = sub; j0 < n; j0 += NP * tpr) {
          asm volatile("" ::: "memory");  // (keeps the loop-invariant weight loads inside the loop, see above)
          int jj[NP];
          bool ok[NP];
          f2_t dot2[NP], nrm2[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            ok[q] = j0 + q * tpr < n;
            jj[q] = ok[q] ? j0 + q * tpr : j0;
            float dt, nr;
            pair_invariants(s_r + 3 * jj[q], ri, dt, nr);
            dot2[q] = f2_pack(dt, dt);
            nrm2[q] = f2_pack(nr, nr);
          }
          // value net, first layer: h = nrm w1 + (dot w0 + b)
          f2_t h[NP][H2];
#pragma unroll
          for (int k = 0; k < H2; ++k) {
            const f2_t w0 = W2[(L::oWv1 >> 1) + k], w1 = W2[((L::oWv1 + HP) >> 1) + k], bb = W2[(L::obv1 >> 1) + k];
#pragma unroll
            for (int q = 0; q < NP; ++q) h[q][k] = f2_fma(nrm2[q], w1, f2_fma(dot2[q], w0, bb));
          }
          // layer normalisation over the H true columns (scalar sums in column order, as the scalar kernel)
          f2_t mean2[NP], rstd2[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            float mu = 0.f;
#pragma unroll
            for (int k = 0; k < H2; ++k) {
              float a0, a1;
              f2_unpack(h[q][k], a0, a1);
              mu += (2 * k < H) ? a0 : 0.f;
              mu += (2 * k + 1 < H) ? a1 : 0.f;
            }
            mu *= inv_H;
            float var = 0.f;
#pragma unroll
            for (int k = 0; k < H2; ++k) {
              float a0, a1;
              f2_unpack(h[q][k], a0, a1);
              const float d0 = a0 - mu, d1 = a1 - mu;
              var += (2 * k < H) ? d0 * d0 : 0.f;
              var += (2 * k + 1 < H) ? d1 * d1 : 0.f;
            }
            const float rs = rsqrtf(var * inv_H + p.ln_eps);
            mean2[q] = f2_pack(-mu, -mu);
            rstd2[q] = f2_pack(rs, rs);
          }
          f2_t iv[NP][D2];
#pragma unroll
          for (int d = 0; d < D2; ++d) {
            const f2_t bb = W2[(L::obv2 >> 1) + d];
#pragma unroll
            for (int q = 0; q < NP; ++q) iv[q][d] = bb;
          }
#pragma unroll
          for (int k = 0; k < H2; ++k) {
            const f2_t lg = W2[(L::oLg >> 1) + k], lb = W2[(L::oLb >> 1) + k];
            f2_t a_lo[NP], a_hi[NP];  // activations of hidden units 2k, 2k + 1, each duplicated into a pair
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              // ((h - mean) rstd) gamma + beta, the scalar kernel's operation order
              const f2_t y = f2_fma(f2_mul(f2_add(h[q][k], mean2[q]), rstd2[q]), lg, lb);
              float y0, y1;
              f2_unpack(y, y0, y1);
              y0 = act_t<ACT>(y0);
              y1 = act_t<ACT>(y1);
              a_lo[q] = f2_pack(y0, y0);
              a_hi[q] = f2_pack(y1, y1);
            }
#pragma unroll
            for (int d = 0; d < D2; ++d) {
              const f2_t w0 = W2[((L::oWv2 + (2 * k) * DP) >> 1) + d], w1 = W2[((L::oWv2 + (2 * k + 1) * DP) >> 1) + d];
#pragma unroll
              for (int q = 0; q < NP; ++q) iv[q][d] = f2_fma(a_hi[q], w1, f2_fma(a_lo[q], w0, iv[q][d]));
            }
          }
          // joined = iv J1 + (u_j + w_i) J2
          f2_t jn[NP][D2];
#pragma unroll
          for (int q = 0; q < NP; ++q)
#pragma unroll
            for (int d = 0; d < D2; ++d) jn[q][d] = 0ull;
#pragma unroll
          for (int c = 0; c < D2; ++c) {
            f2_t a_lo[NP], a_hi[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              float x0, x1;
              f2_unpack(iv[q][c], x0, x1);
              a_lo[q] = f2_pack(x0, x0);
              a_hi[q] = f2_pack(x1, x1);
            }
#pragma unroll
            for (int d = 0; d < D2; ++d) {
              const f2_t w0 = W2[((L::oJ1 + (2 * c) * DP) >> 1) + d], w1 = W2[((L::oJ1 + (2 * c + 1) * DP) >> 1) + d];
#pragma unroll
              for (int q = 0; q < NP; ++q) jn[q][d] = f2_fma(a_hi[q], w1, f2_fma(a_lo[q], w0, jn[q][d]));
            }
          }
          {
            f2_t j2[NP][D2];
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
              for (int d = 0; d < D2; ++d) j2[q][d] = 0ull;
#pragma unroll
            for (int c = 0; c < D2; ++c) {
              f2_t a_lo[NP], a_hi[NP];
              const f2_t wi = wi2[c];
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                const f2_t mg = f2_add(reinterpret_cast<const f2_t*>(s_u + jj[q] * DP)[c], wi);
                float x0, x1;
                f2_unpack(mg, x0, x1);
                a_lo[q] = f2_pack(x0, x0);
                a_hi[q] = f2_pack(x1, x1);
              }
#pragma unroll
              for (int d = 0; d < D2; ++d) {
                const f2_t w0 = W2[((L::oJ2 + (2 * c) * DP) >> 1) + d], w1 = W2[((L::oJ2 + (2 * c + 1) * DP) >> 1) + d];
#pragma unroll
                for (int q = 0; q < NP; ++q) j2[q][d] = f2_fma(a_hi[q], w1, f2_fma(a_lo[q], w0, j2[q][d]));
              }
            }
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
              for (int d = 0; d < D2; ++d) jn[q][d] = f2_add(jn[q][d], j2[q][d]);
          }
          // score net: eight hidden units (four pairs) at a time
          float sc[NP];
#pragma unroll
          for (int q = 0; q < NP; ++q) sc[q] = W[L::obs2];
#pragma unroll
          for (int k0 = 0; k0 < H2; k0 += 2) {
            f2_t hh[NP][2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const f2_t bb = W2[(L::obs1 >> 1) + k0 + e];
#pragma unroll
              for (int q = 0; q < NP; ++q) hh[q][e] = bb;
            }
#pragma unroll
            for (int c = 0; c < D2; ++c) {
              f2_t a_lo[NP], a_hi[NP];
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                float x0, x1;
                f2_unpack(jn[q][c], x0, x1);
                a_lo[q] = f2_pack(x0, x0);
                a_hi[q] = f2_pack(x1, x1);
              }
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const f2_t w0 = W2[((L::oWs1 + (2 * c) * HP) >> 1) + k0 + e], w1 = W2[((L::oWs1 + (2 * c + 1) * HP) >> 1) + k0 + e];
#pragma unroll
                for (int q = 0; q < NP; ++q) hh[q][e] = f2_fma(a_hi[q], w1, f2_fma(a_lo[q], w0, hh[q][e]));
              }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float w0, w1;
              f2_unpack(W2[(L::oWs2 >> 1) + k0 + e], w0, w1);
#pragma unroll
              for (int q = 0; q < NP; ++q) {
                float x0, x1;
                f2_unpack(hh[q][e], x0, x1);
                sc[q] = fmaf(act_t<ACT>(x0), w0, sc[q]);
                sc[q] = fmaf(act_t<ACT>(x1), w1, sc[q]);
              }
            }
          }
          // online softmax, pair by pair in ascending j
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            if (!ok[q]) continue;
            float s1 = sc[q];
            if (!(mi && s_m[jj[q]])) s1 = kMaskedScore;
            const float mn = fmaxf(m, s1);
            const float scale = expf(m - mn);
            const float e = expf(s1 - mn);
            l = l * scale + e;
            const f2_t sc2 = f2_pack(scale, scale), e2 = f2_pack(e, e);
#pragma unroll
            for (int d = 0; d < D2; ++d) acc[d] = f2_fma(acc[d], sc2, f2_mul(e2, jn[q][d]));
            m = mn;
          }
        }
      }
      float* st = s_st + t * (DP + 2);
      st[0] = m; st[1] = l;
#pragma unroll
      for (int d = 0; d < D2; ++d) f2_unpack(acc[d], st[2 + 2 * d], st[3 + 2 * d]);
      __syncthreads();
      if (!p.reduce) {
        if (active) {
          const int first = (t / tpr) * tpr;
          float M = -INFINITY;
          for (int q = 0; q < tpr; ++q) M = fmaxf(M, s_st[(first + q) * (DP + 2)]);
          float Ls = 0.f;
          for (int q = 0; q < tpr; ++q) {
            const float* sq = s_st + (first + q) * (DP + 2);
            if (sq[1] > 0.f) Ls += sq[1] * expf(sq[0] - M);
          }
          for (int d = sub; d < D; d += tpr) {
            float o = 0.f;
            for (int q = 0; q < tpr; ++q) {
              const float* sq = s_st + (first + q) * (DP + 2);
              if (sq[1] > 0.f) o += sq[2 + d] * expf(sq[0] - M);
            }
            p.out[(b * n + i) * D + d] = o / Ls;
          }
        }
      } else if (t < DP) {
        const int nact = min(rows_per_pass, n - i0) * tpr;
        float M = red_m;
        for (int q = 0; q < nact; ++q) M = fmaxf(M, s_st[q * (DP + 2)]);
        const float sc0 = red_l > 0.f ? expf(red_m - M) : 0.f;
        float Ls = red_l * sc0, o = red_o * sc0;
        for (int q = 0; q < nact; ++q) {
          const float* sq = s_st + q * (DP + 2);
          if (sq[1] > 0.f) {
            const float f = expf(sq[0] - M);
            Ls += sq[1] * f;
            o += sq[2 + t] * f;
          }
        }
        red_m = M; red_l = Ls; red_o = o;
      }
      __syncthreads();
    }
    if (p.reduce && t < D) p.out[b * D + t] = red_o / red_l;
  }
}

// the packed image (global memory), one buffer per device, reused by successive launches on the library's stream
float* gaa_image_buffer() {
  static float* buf[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!buf[dev] && cudaMalloc(&buf[dev], sizeof(float) * GaaImg<64, 32>::nW) != cudaSuccess) buf[dev] = nullptr;
  return buf[dev];
}

bool gaa_const_weights() {  // VMS_GAA_CONST=1 (and a build with -DVMS_GAA_WITH_CONST): weights as constant-bank operands
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VMS_GAA_CONST");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

// pairs per trip of the pair loop: 2 when a thread walks at least four pairs (k = 50, scalar arithmetic: 5.6 vs 6.4 ms per
// 4,096-site embedding -- every weight broadcast then feeds two FMAs; packed arithmetic on top: 4.7 ms), 1 for small clouds
// (k = 10: 0.56 vs 0.93 ms); VMS_GAA_NP=1|2 forces it
int gaa_pairs_per_trip(int n, int tpr) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("VMS_GAA_NP");
    forced = (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
  }
  if (forced) return forced;
  return (n + tpr - 1) / tpr >= 4 ? 2 : 1;
}

template <int HP, int DP, int ACT, bool CW, int NP>
vms_status gaa_fused_launch2(GaaFwdParams& p, cudaStream_t st) {
  using L = GaaImg<HP, DP>;
  int threads = ((p.n * p.tpr + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < DP) threads = ((DP + 31) / 32) * 32;
  const size_t fl = (size_t)(CW ? 0 : L::nPair) + ((3 * p.n + 3) & ~3) + (size_t)3 * p.n * DP + (size_t)threads * (DP + 2);
  const size_t smem = fl * sizeof(float) + (size_t)((p.n + 15) & ~15);
  VMS_REQUIRE(smem <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "gaa_attention_forward: cloud of %d particles does not fit",
              p.n);
  p.img = gaa_image_buffer();
  VMS_REQUIRE(p.img, VMS_ERR_CUDA, "gaa_attention_forward: cannot allocate the weight image");
  gaa_pack_kernel<HP, DP><<<(L::nW + 255) / 256, 256, 0, st>>>(p);
  VMS_LAUNCH_CHECK("gaa_pack_kernel");
  if (CW) VMS_CUDA(cudaMemcpyToSymbolAsync(c_gaa_w, p.img, sizeof(float) * L::nPair, 0, cudaMemcpyDeviceToDevice, st));
  VMS_CUDA(cudaFuncSetAttribute(gaa_attention_fwd_kernel<HP, DP, ACT, CW, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t cap = 8 * (int64_t)sm_count();
  const int grid = (int)(p.B < cap ? p.B : cap);
  gaa_attention_fwd_kernel<HP, DP, ACT, CW, NP><<<grid, threads, smem, st>>>(p);
  VMS_LAUNCH_CHECK("gaa_attention_fwd_kernel");
  return VMS_OK;
}

template <int HP, int DP, int ACT>
vms_status gaa_fused_launch_x2(GaaFwdParams& p, cudaStream_t st) {
  using L = GaaImg<HP, DP>;
  int threads = ((p.n * p.tpr + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < DP) threads = ((DP + 31) / 32) * 32;
  const size_t fl = (size_t)L::nPair + ((3 * p.n + 3) & ~3) + (size_t)3 * p.n * DP + (size_t)threads * (DP + 2);
  const size_t smem = fl * sizeof(float) + (size_t)((p.n + 15) & ~15);
  VMS_REQUIRE(smem <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "gaa_attention_forward: cloud of %d particles does not fit",
              p.n);
  p.img = gaa_image_buffer();
  VMS_REQUIRE(p.img, VMS_ERR_CUDA, "gaa_attention_forward: cannot allocate the weight image");
  gaa_pack_kernel<HP, DP><<<(L::nW + 255) / 256, 256, 0, st>>>(p);
  VMS_LAUNCH_CHECK("gaa_pack_kernel");
  VMS_CUDA(cudaFuncSetAttribute(gaa_attention_fwd_x2_kernel<HP, DP, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t cap = 8 * (int64_t)sm_count();
  const int grid = (int)(p.B < cap ? p.B : cap);
  gaa_attention_fwd_x2_kernel<HP, DP, ACT><<<grid, threads, smem, st>>>(p);
  VMS_LAUNCH_CHECK("gaa_attention_fwd_x2_kernel");
  return VMS_OK;
}

template <int HP, int DP>
vms_status gaa_fused_launch(GaaFwdParams& p, cudaStream_t st) {
  static int x2 = -1;  // VMS_GAA_X2=0: the scalar two-pair kernel instead of the packed one
  if (x2 < 0) {
    const char* e = getenv("VMS_GAA_X2");
    x2 = (e && e[0] == '0') ? 0 : 1;
  }
  if (x2 && gaa_pairs_per_trip(p.n, p.tpr) == 2) {
    if (p.act == VMS_ACT_NONE) return gaa_fused_launch_x2<HP, DP, VMS_ACT_NONE>(p, st);
    if (p.act == VMS_ACT_RELU) return gaa_fused_launch_x2<HP, DP, VMS_ACT_RELU>(p, st);
    if (p.act == VMS_ACT_TANH) return gaa_fused_launch_x2<HP, DP, VMS_ACT_TANH>(p, st);
  }
#ifdef VMS_GAA_WITH_CONST  // constant-bank weights: measured slower than shared-memory broadcasts (7.4 vs 6.4 ms), not built by default
  if (gaa_const_weights()) {
#define VMS_GAA_ACT(A) if (p.act == A) return gaa_fused_launch2<HP, DP, A, true, 1>(p, st)
    VMS_GAA_ACT(VMS_ACT_NONE);
    VMS_GAA_ACT(VMS_ACT_RELU);
    VMS_GAA_ACT(VMS_ACT_TANH);
#undef VMS_GAA_ACT
  }
#endif
  // small clouds (a thread walks fewer than four pairs) and VMS_GAA_X2=0: the scalar one-pair-per-trip kernel
#define VMS_GAA_ACT(A) if (p.act == A) return gaa_fused_launch2<HP, DP, A, false, 1>(p, st)
  VMS_GAA_ACT(VMS_ACT_NONE);
  VMS_GAA_ACT(VMS_ACT_RELU);
  VMS_GAA_ACT(VMS_ACT_TANH);
#undef VMS_GAA_ACT
  return VMS_ERR_INVALID_ARG;
}

}  // namespace
}  // namespace vms

using namespace vms;

extern "C" {

vms_status vms_gaa_pair_invariants(const float* coords, int64_t B, int n, float* out, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && n >= 1 && (B == 0 || (coords && out)), VMS_ERR_INVALID_ARG, "gaa_pair_invariants: bad arguments");
  const int64_t total = B * n * n;
  if (total == 0) return VMS_OK;
  gaa_pair_inv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(coords, total, n, out);
  VMS_LAUNCH_CHECK("gaa_pair_inv_kernel");
  return VMS_OK;
}

vms_status vms_gaa_zero_mask(const float* coords, int64_t n_particles, uint8_t* mask, vms_stream stream) {
  VMS_REQUIRE(n_particles >= 0 && (n_particles == 0 || (coords && mask)), VMS_ERR_INVALID_ARG, "gaa_zero_mask: bad arguments");
  if (n_particles == 0) return VMS_OK;
  gaa_zero_mask_kernel<<<(unsigned)((n_particles + 255) / 256), 256, 0, as_stream(stream)>>>(coords, n_particles, mask);
  VMS_LAUNCH_CHECK("gaa_zero_mask_kernel");
  return VMS_OK;
}

vms_status vms_layernorm_forward(const float* x, int64_t ld_x, int64_t R, int H, const float* gamma, const float* beta,
                                 float eps, int act, float* y, int64_t ld_y, float* stats, vms_stream stream) {
  VMS_REQUIRE(R >= 0 && H >= 1 && (R == 0 || (x && y && gamma && beta)), VMS_ERR_INVALID_ARG, "layernorm_forward: bad arguments");
  VMS_REQUIRE(act >= 0 && act <= 2 && eps >= 0.f, VMS_ERR_INVALID_ARG, "layernorm_forward: bad activation / epsilon");
  if (R == 0) return VMS_OK;
  layernorm_fwd_kernel<<<(unsigned)((R * LNG + 255) / 256), 256, 0, as_stream(stream)>>>(x, ld_x, R, H, gamma, beta, eps, act, y,
                                                                                         ld_y, stats);
  VMS_LAUNCH_CHECK("layernorm_fwd_kernel");
  return VMS_OK;
}

size_t vms_layernorm_backward_workspace(int64_t R, int H) {
  return (size_t)ln_bwd_grid(R) * 2 * (size_t)(H > 0 ? H : 1) * sizeof(float);
}

vms_status vms_layernorm_backward(const float* x, int64_t ld_x, int64_t R, int H, const float* gamma, const float* stats,
                                  int act, const float* y, int64_t ld_y, const float* g_y, int64_t ld_gy, float* g_x,
                                  int64_t ld_gx, float* g_gamma, float* g_beta, void* workspace, vms_stream stream) {
  VMS_REQUIRE(R >= 0 && H >= 1 && (R == 0 || (x && y && gamma && stats && g_y && workspace)), VMS_ERR_INVALID_ARG,
              "layernorm_backward: bad arguments");
  VMS_REQUIRE(H <= 128, VMS_ERR_UNSUPPORTED, "layernorm_backward: H = %d > 128", H);
  if (R == 0) return VMS_OK;
  cudaStream_t st = as_stream(stream);
  const int grid = ln_bwd_grid(R);
  const size_t smem = (size_t)32 * 2 * H * sizeof(float);
  float* part = (float*)workspace;
  const int E = (H + LNG - 1) / LNG;
#define VMS_LN_BWD(EE)                                                                                                    \
  layernorm_bwd_kernel<EE><<<grid, 256, smem, st>>>(x, ld_x, R, H, gamma, stats, act, y, ld_y, g_y, ld_gy, g_x, ld_gx, part)
  if (E <= 2) VMS_LN_BWD(2);
  else if (E <= 4) VMS_LN_BWD(4);
  else if (E <= 5) VMS_LN_BWD(5);
  else if (E <= 8) VMS_LN_BWD(8);
  else VMS_LN_BWD(16);
#undef VMS_LN_BWD
  VMS_LAUNCH_CHECK("layernorm_bwd_kernel");
  return sum_partials_launch(part, grid, 2 * (int64_t)H, H, g_gamma, H, g_beta, 1.f, 1, st);
}

vms_status vms_gaa_pair_merge(const float* u, int64_t ld_u, const float* w, int64_t ld_w, int64_t B, int n, int D, float* out,
                              vms_stream stream) {
  VMS_REQUIRE(B >= 0 && n >= 1 && D >= 1 && (B == 0 || (u && w && out)), VMS_ERR_INVALID_ARG, "gaa_pair_merge: bad arguments");
  const int64_t total = B * n * n * D;
  if (total == 0) return VMS_OK;
  gaa_pair_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(u, ld_u, w, ld_w, total, n, D, out);
  VMS_LAUNCH_CHECK("gaa_pair_merge_kernel");
  return VMS_OK;
}

vms_status vms_gaa_pair_merge_backward(const float* g, int64_t B, int n, int D, float* g_u, int64_t ld_gu, float* g_w,
                                       int64_t ld_gw, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && n >= 1 && D >= 1 && (B == 0 || (g && g_u && g_w)), VMS_ERR_INVALID_ARG,
              "gaa_pair_merge_backward: bad arguments");
  const int64_t total = B * n * D;
  if (total == 0) return VMS_OK;
  gaa_pair_merge_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(g, total, n, D, g_u, ld_gu, g_w,
                                                                                            ld_gw);
  VMS_LAUNCH_CHECK("gaa_pair_merge_bwd_kernel");
  return VMS_OK;
}

vms_status vms_gaa_attend(const float* scores, const float* values, const uint8_t* mask, int64_t B, int n, int D, int reduce,
                          float* out, float* attention, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && n >= 1 && D >= 1 && (B == 0 || (scores && values && out && attention)), VMS_ERR_INVALID_ARG,
              "gaa_attend: bad arguments");
  if (B == 0) return VMS_OK;
  cudaStream_t st = as_stream(stream);
  if (reduce) {
    VMS_REQUIRE(B < (1LL << 31), VMS_ERR_SHAPE, "gaa_attend: too many clouds");
    gaa_attend_all_kernel<<<(unsigned)B, AT, 0, st>>>(scores, values, mask, n, D, out, attention);
    VMS_LAUNCH_CHECK("gaa_attend_all_kernel");
  } else {
    const int64_t rows = B * n;
    gaa_attend_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(scores, values, mask, rows, n, D, out, attention);
    VMS_LAUNCH_CHECK("gaa_attend_rows_kernel");
  }
  return VMS_OK;
}

vms_status vms_gaa_attend_backward(const float* attention, const float* values, const uint8_t* mask, int64_t B, int n, int D,
                                   int reduce, const float* g_out, float* g_scores, float* g_values, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && n >= 1 && D >= 1 && (B == 0 || (attention && values && g_out)), VMS_ERR_INVALID_ARG,
              "gaa_attend_backward: bad arguments");
  if (B == 0) return VMS_OK;
  cudaStream_t st = as_stream(stream);
  if (reduce) {
    gaa_attend_all_bwd_kernel<<<(unsigned)B, AT, 0, st>>>(attention, values, mask, n, D, g_out, g_scores, g_values);
    VMS_LAUNCH_CHECK("gaa_attend_all_bwd_kernel");
  } else {
    const int64_t rows = B * n;
    gaa_attend_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(attention, values, mask, rows, n, D, g_out, g_scores,
                                                                           g_values);
    VMS_LAUNCH_CHECK("gaa_attend_rows_bwd_kernel");
  }
  return VMS_OK;
}

int vms_gaa_attention_forward_supported(int n, int D, int H) {
  return (n >= 1 && D >= 1 && D <= 32 && H >= 1 && H <= 64) ? 1 : 0;
}

vms_status vms_gaa_attention_forward(const float* coords, const float* values, int64_t ld_v, const uint8_t* mask, int64_t B,
                                     int n, int D, int H, const vms_gaa_weights* w, int reduce, int act, float ln_eps,
                                     float* out, vms_stream stream) {
  VMS_RANGE("vms_gaa_attention_forward");
  VMS_REQUIRE(coords && values && w && out && B >= 0, VMS_ERR_INVALID_ARG, "gaa_attention_forward: NULL pointer");
  VMS_REQUIRE(vms_gaa_attention_forward_supported(n, D, H), VMS_ERR_UNSUPPORTED,
              "gaa_attention_forward: needs D <= 32 and H <= 64 (got D = %d, H = %d)", D, H);
  VMS_REQUIRE(w->merge0 && w->merge1 && w->join1 && w->join2 && w->score_w1 && w->score_b1 && w->score_w2 && w->score_b2 &&
                  w->value_w1 && w->value_b1 && w->value_gamma && w->value_beta && w->value_w2 && w->value_b2,
              VMS_ERR_INVALID_ARG, "gaa_attention_forward: NULL weight");
  VMS_REQUIRE(act >= 0 && act <= 2, VMS_ERR_INVALID_ARG, "gaa_attention_forward: unknown activation");
  if (B == 0) return VMS_OK;
  GaaFwdParams p = {};
  p.coords = coords; p.v = values; p.ldv = ld_v; p.mask = mask; p.B = B; p.n = n; p.D = D; p.H = H; p.reduce = reduce; p.act = act;
  p.M0 = w->merge0; p.M1 = w->merge1; p.J1 = w->join1; p.J2 = w->join2;
  p.Ws1 = w->score_w1; p.bs1 = w->score_b1; p.Ws2 = w->score_w2; p.bs2 = w->score_b2;
  p.Wv1 = w->value_w1; p.bv1 = w->value_b1; p.lg = w->value_gamma; p.lb = w->value_beta; p.Wv2 = w->value_w2; p.bv2 = w->value_b2;
  p.ln_eps = ln_eps; p.out = out;
  int tpr = 256 / n;
  if (tpr < 1) tpr = 1;
  if (tpr > n) tpr = n;
  p.tpr = tpr;
  cudaStream_t st = as_stream(stream);
  // register arrays are compile-time: the widths of the reference's defaults (hidden 40, embedding 20) or the maximum
  if (H <= 40 && D <= 20) return gaa_fused_launch<40, 20>(p, st);
  return gaa_fused_launch<64, 32>(p, st);
}

}  // extern "C"
