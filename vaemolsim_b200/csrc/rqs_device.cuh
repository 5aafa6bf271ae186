// rqs_device.cuh -- octet-cooperative rational-quadratic-spline device routines shared by the standalone RQS kernels
// (rqs.cu) and the fused ELBO kernel (elbo_fused.cu).  See rqs.cu for the design notes and the reference citations
// (flows.py:86-101, :394-409, :204-207, :512-515; tfp.bijectors.RationalQuadraticSpline).
//
// 8 lanes ("octet") cooperate on one element; each lane owns BPL consecutive bins.  NC selects the read-only global
// path (__ldg) for the raw logits; the fused kernel reads them from shared memory (NC = false).
#pragma once
#include "common.cuh"
#include <math.h>

namespace vms {
namespace rqsdev {

constexpr int kOct = 8;

template <bool NC>
__device__ __forceinline__ float ld1(const float* p) {
  if (NC) return __ldg(p);
  return *p;
}
template <bool NC>
__device__ __forceinline__ float4 ld4(const float* p) {
  if (NC) return __ldg(reinterpret_cast<const float4*>(p));
  return *reinterpret_cast<const float4*>(p);
}

__device__ __forceinline__ float oct_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}
// inclusive prefix sum over the 8 lanes of an octet
__device__ __forceinline__ double oct_scan(double v, int j) {
#pragma unroll
  for (int d = 1; d < kOct; d <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, d, kOct);
    if (j >= d) v += t;
  }
  return v;
}

// logits of this lane's BPL consecutive bins (bins >= K read as -inf)
template <int BPL, bool VEC, bool NC>
__device__ __forceinline__ void load_bins(const float* __restrict__ row, int j, int K, float (&out)[BPL]) {
  if (VEC) {
#pragma unroll
    for (int q = 0; q < BPL / 4; ++q) {
      const int k0 = j * BPL + 4 * q;
      if (k0 < K) {  // K % 4 == 0 on the vector path: a 4-bin group is entirely valid or entirely padding
        const float4 t = ld4<NC>(row + k0);
        out[4 * q] = t.x; out[4 * q + 1] = t.y; out[4 * q + 2] = t.z; out[4 * q + 3] = t.w;
      } else {
        out[4 * q] = out[4 * q + 1] = out[4 * q + 2] = out[4 * q + 3] = -INFINITY;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < BPL; ++i) {
      const int k = j * BPL + i;
      out[i] = k < K ? ld1<NC>(row + k) : -INFINITY;
    }
  }
}

template <int BPL, bool VEC>
__device__ __forceinline__ void store_bins(float* __restrict__ row, int j, int K, const float (&v)[BPL]) {
  if (VEC) {
#pragma unroll
    for (int q = 0; q < BPL / 4; ++q) {
      const int k0 = j * BPL + 4 * q;
      if (k0 < K) *reinterpret_cast<float4*>(row + k0) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < BPL; ++i) {
      const int k = j * BPL + i;
      if (k < K) row[k] = v[i];
    }
  }
}

// What the owning lane knows about the element's bin.
struct Bin {
  double lo_x, lo_y;      // lower knots
  float wk, hk;           // bin width / height
  float e_w, e_h;         // exp(logit - max) of the bin
  float elt_w, elt_h;     // sum of exps of all lower bins
  int idx;
  bool found;
};

// 1 / x for a positive, normal double (a softmax denominator, 1 <= x <= bins): float32 reciprocal refined by two Newton
// steps in float64 (relative error ~1e-16 after the second; the IEEE division it replaces costs ~30 dependent FP64
// instructions per lane and was the top stall of the tensor-core coupling kernel, where nothing hides it).
__device__ __forceinline__ double recip_pos(double x) {
  double r = (double)__frcp_rn((float)x);
  r = fma(r, fma(-x, r, 1.0), r);
  r = fma(r, fma(-x, r, 1.0), r);
  return r;
}

// Softmax statistics, octet scan, knot walk over this lane's bins.  On return ew / eh hold exp(logit - max),
// inv_tot_* = 1 / sum(exp), c* = scale / sum(exp).
template <int BPL>
__device__ __forceinline__ Bin find_bin(float (&ew)[BPL], float (&eh)[BPL], int j, int K, double vd, bool inverse_dir,
                                        float bin_min, float scale, double& cwd, double& chd, double& totw,
                                        double& toth) {
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
  for (int i = 0; i < BPL; ++i) { mw = fmaxf(mw, ew[i]); mh = fmaxf(mh, eh[i]); }
  mw = oct_max(mw);
  mh = oct_max(mh);
  double sw = 0.0, sh = 0.0;
#pragma unroll
  for (int i = 0; i < BPL; ++i) {
    ew[i] = expf(ew[i] - mw);  // exp(-inf) = 0 for padding bins
    eh[i] = expf(eh[i] - mh);
    sw += (double)ew[i];
    sh += (double)eh[i];
  }
  const double iw = oct_scan(sw, j), ih = oct_scan(sh, j);
  totw = __shfl_sync(0xffffffffu, iw, kOct - 1, kOct);
  toth = __shfl_sync(0xffffffffu, ih, kOct - 1, kOct);
  cwd = (double)scale * recip_pos(totw);
  chd = (double)scale * recip_pos(toth);
  // knots: k-th lower knot = bin_min + (scale * E_k / total + 1e-2 k), E_k = sum of exps of bins < k
  Bin b;
  b.found = false;
  b.idx = 0;
  b.lo_x = b.lo_y = 0.0; b.wk = b.hk = 1.f; b.e_w = b.e_h = 0.f; b.elt_w = b.elt_h = 0.f;
  // a lane's first lower knot and its left neighbour's last upper knot are computed from the SAME scan value, so the
  // bins tile the range without gaps or overlaps
  double Ex = __shfl_up_sync(0xffffffffu, iw, 1, kOct), Ey = __shfl_up_sync(0xffffffffu, ih, 1, kOct);
  if (j == 0) { Ex = 0.0; Ey = 0.0; }
  const double bm = (double)bin_min;
  double lox = bm + fma(cwd, Ex, 1e-2 * (double)(j * BPL));
  double loy = bm + fma(chd, Ey, 1e-2 * (double)(j * BPL));
#pragma unroll
  for (int i = 0; i < BPL; ++i) {
    const int k = j * BPL + i;
    const double Ex1 = i == BPL - 1 ? iw : Ex + (double)ew[i];
    const double Ey1 = i == BPL - 1 ? ih : Ey + (double)eh[i];
    const double hix = bm + fma(cwd, Ex1, 1e-2 * (double)(k + 1));
    const double hiy = bm + fma(chd, Ey1, 1e-2 * (double)(k + 1));
    const double lo = inverse_dir ? loy : lox, hi = inverse_dir ? hiy : hix;
    // bin k covers [lo, hi); the range edge itself is outside (TFP: x <= kx[0] or x >= kx[K] => identity)
    if (k < K && vd >= lo && vd < hi && vd > bm) {
      b.found = true;
      b.idx = k;
      b.lo_x = lox; b.lo_y = loy;
      b.wk = (float)(hix - lox); b.hk = (float)(hiy - loy);
      b.e_w = ew[i]; b.e_h = eh[i];
      b.elt_w = (float)Ex; b.elt_h = (float)Ey;
    }
    Ex = Ex1; Ey = Ey1; lox = hix; loy = hiy;
  }
  return b;
}

// relative position r in the bin for either direction (TFP _forward / _inverse)
__device__ __forceinline__ float rel_pos(const Bin& b, double vd, float sk, float dk, float dk1, bool inverse_dir) {
  if (!inverse_dir) return (float)(vd - b.lo_x) / b.wk;
  const float ry = (float)(vd - b.lo_y);
  const float t2 = ry * (dk1 + dk - 2.f * sk);
  const float a = b.hk * (sk - dk) + t2;
  const float bb = b.hk * dk - t2;
  const float c = -sk * ry;
  const float disc = bb * bb - 4.f * a * c;
  const float r = (2.f * c) / (-bb - sqrtf(disc));
  return ry == 0.f ? 0.f : r;
}

template <bool NC>
__device__ __forceinline__ void load_slopes(const float* __restrict__ ps, int idx, int K, float& s_lo, float& s_hi,
                                            float& dk, float& dk1) {
  s_lo = idx > 0 ? ld1<NC>(ps + idx - 1) : 0.f;
  s_hi = idx < K - 1 ? ld1<NC>(ps + idx) : 0.f;
  dk = idx == 0 ? 1.0f : softplus_tf(s_lo) + 1e-2f;
  dk1 = idx == K - 1 ? 1.0f : softplus_tf(s_hi) + 1e-2f;
}


// ---------------------------------------------------------------------------------------------- octet-level ops
// One element per octet.  pw / ph point at the element's K raw width / height logits, ps at its K-1 raw slopes.
// Every lane of the warp must call these (full-mask shuffles).  Results:
//   out, ldj   valid on the WRITER lane (the lane owning the bin; lane 0 of the octet when out of range: identity, 0)
//   ldj_all    the element's log-det broadcast to all 8 lanes
template <int BPL, bool VEC, bool NC>
__device__ __forceinline__ void octet_apply(const float* __restrict__ pw, const float* __restrict__ ph,
                                            const float* __restrict__ ps, float v, int j, int K, bool inv, float bin_min,
                                            float scale, float& out, float& ldj, float& ldj_all, bool& writer) {
  float ew[BPL], eh[BPL];
  load_bins<BPL, VEC, NC>(pw, j, K, ew);
  load_bins<BPL, VEC, NC>(ph, j, K, eh);
  const double vd = (double)v;
  double cwd, chd, totw, toth;
  const Bin b = find_bin<BPL>(ew, eh, j, K, vd, inv, bin_min, scale, cwd, chd, totw, toth);
  out = v;
  ldj = 0.f;
  if (b.found) {
    float s_lo, s_hi, dk, dk1;
    load_slopes<NC>(ps, b.idx, K, s_lo, s_hi, dk, dk1);
    const float sk = b.hk / b.wk;
    const float rr = rel_pos(b, vd, sk, dk, dk1, inv);
    const float omr = 1.f - rr, u = rr * omr;
    const float den = sk + (dk1 + dk - 2.f * sk) * u;
    if (!inv) {
      const float num = b.hk * (sk * rr * rr + dk * u);
      out = (float)(b.lo_y + (double)(num / den));
    } else {
      out = (float)(b.lo_x + (double)(rr * b.wk));
    }
    const float P = dk1 * rr * rr + 2.f * sk * u + dk * omr * omr;
    ldj = logf((sk * sk) * P / (den * den));
    if (inv) ldj = -ldj;
  }
  const unsigned found_mask = __ballot_sync(0xffffffffu, b.found);
  const unsigned oct_mask = (found_mask >> ((threadIdx.x & 31) & ~(kOct - 1))) & 0xffu;
  writer = oct_mask ? b.found : (j == 0);  // out-of-range: lane 0 writes the identity
  const int src = oct_mask ? (__ffs(oct_mask) - 1) : 0;
  ldj_all = __shfl_sync(0xffffffffu, ldj, src, kOct);
}

// Reverse mode of octet_apply.  g_out / g_ldj: upstream gradients of out / ldj.
//   g_in valid on the WRITER lane;  gw / gh / gs: this lane's BPL entries of the raw-logit gradients (bins j*BPL + i;
//   gs[i] belongs to raw slope index j*BPL + i, valid for < K-1).  tf.where semantics out of range: g_in = g_out, 0.
template <int BPL, bool VEC, bool NC>
__device__ __forceinline__ void octet_backward(const float* __restrict__ pw, const float* __restrict__ ph,
                                               const float* __restrict__ ps, float v, float g_out, float g_ldj, int j,
                                               int K, bool inv, float bin_min, float scale, float& g_in, bool& writer,
                                               float (&gw)[BPL], float (&gh)[BPL], float (&gs)[BPL]) {
  float ew[BPL], eh[BPL];
  load_bins<BPL, VEC, NC>(pw, j, K, ew);
  load_bins<BPL, VEC, NC>(ph, j, K, eh);
  const double vd = (double)v;
  double cwd, chd, totw, toth;
  const Bin b = find_bin<BPL>(ew, eh, j, K, vd, inv, bin_min, scale, cwd, chd, totw, toth);
  // owner-lane results, broadcast to the octet below
  float g_xk = 0.f, g_w = 0.f, g_yk = 0.f, g_h = 0.f, gs_lo = 0.f, gs_hi = 0.f, dotw = 0.f, doth = 0.f;
  g_in = g_out;
  if (b.found) {
    float s_lo, s_hi, dk, dk1;
    load_slopes<NC>(ps, b.idx, K, s_lo, s_hi, dk, dk1);
    // local derivatives: y = yk + h N/Q, L = log(s^2 P / Q^2)
    const float h = b.hk, w = b.wk, s = h / w;
    const float rr = rel_pos(b, vd, s, dk, dk1, inv);
    const float omr = 1.f - rr, u = rr * omr, tm = 1.f - 2.f * rr;
    const float dd = dk1 + dk - 2.f * s;
    const float N = s * rr * rr + dk * u;
    const float Q = s + dd * u;
    const float P = dk1 * rr * rr + 2.f * s * u + dk * omr * omr;
    const float N_r = 2.f * s * rr + dk * tm;
    const float Q_r = dd * tm;
    const float Q_s = 1.f - 2.f * u;
    const float P_r = 2.f * dk1 * rr + 2.f * s * tm - 2.f * dk * omr;
    const float iQ = 1.f / Q, iQ2 = iQ * iQ, iP = 1.f / P, iw = 1.f / w;
    const float y_r = h * (N_r * Q - N * Q_r) * iQ2;
    const float y_s = h * (rr * rr * Q - N * Q_s) * iQ2;
    const float y_dk = h * u * (Q - N) * iQ2;
    const float y_dk1 = -h * N * u * iQ2;
    const float y_h = N * iQ;
    const float L_r = P_r * iP - 2.f * Q_r * iQ;
    const float L_s = 2.f / s + 2.f * u * iP - 2.f * Q_s * iQ;
    const float L_dk = omr * omr * iP - 2.f * u * iQ;
    const float L_dk1 = rr * rr * iP - 2.f * u * iQ;
    const float F_x = y_r * iw, L_x = L_r * iw;
    float gy, gL;
    if (inv) {
      const float G = g_out - g_ldj * L_x;
      g_in = G / F_x;
      gy = -g_in;
      gL = -g_ldj;
    } else {
      gy = g_out;
      gL = g_ldj;
      g_in = gy * F_x + gL * L_x;
    }
    const float g_r = gy * y_r + gL * L_r;
    const float g_s = gy * y_s + gL * L_s;
    const float g_dk = gy * y_dk + gL * L_dk;
    const float g_dk1 = gy * y_dk1 + gL * L_dk1;
    g_h = gy * y_h + g_s * iw;
    g_yk = gy;
    g_w = -(g_s * s + g_r * rr) * iw;
    g_xk = -g_r * iw;
    if (b.idx > 0) gs_lo = g_dk * sigmoidf_(s_lo);
    if (b.idx < K - 1) gs_hi = g_dk1 * sigmoidf_(s_hi);
    // softmax Jacobian dot products: sum_i p_i g_b[i] with g_b[i] = [i<idx] g_k + [i==idx] g_bin
    dotw = (g_xk * b.elt_w + g_w * b.e_w) * (float)(1.0 / totw);
    doth = (g_yk * b.elt_h + g_h * b.e_h) * (float)(1.0 / toth);
  }
  const unsigned found_mask = __ballot_sync(0xffffffffu, b.found);
  const unsigned oct_mask = (found_mask >> ((threadIdx.x & 31) & ~(kOct - 1))) & 0xffu;
  const int src = oct_mask ? (__ffs(oct_mask) - 1) : 0;
  writer = oct_mask ? b.found : (j == 0);
  // the shuffles must be executed by every lane of the warp (octets diverge on oct_mask): select afterwards
  const int idx_owner = __shfl_sync(0xffffffffu, b.idx, src, kOct);
  const int idx = oct_mask ? idx_owner : -2;  // -2: no bin matches any k
  g_xk = __shfl_sync(0xffffffffu, g_xk, src, kOct);
  g_w = __shfl_sync(0xffffffffu, g_w, src, kOct);
  g_yk = __shfl_sync(0xffffffffu, g_yk, src, kOct);
  g_h = __shfl_sync(0xffffffffu, g_h, src, kOct);
  dotw = __shfl_sync(0xffffffffu, dotw, src, kOct);
  doth = __shfl_sync(0xffffffffu, doth, src, kOct);
  gs_lo = __shfl_sync(0xffffffffu, gs_lo, src, kOct);
  gs_hi = __shfl_sync(0xffffffffu, gs_hi, src, kOct);
  const float cw = (float)cwd, ch = (float)chd;
#pragma unroll
  for (int i = 0; i < BPL; ++i) {
    const int k = j * BPL + i;
    const float gbw = k < idx ? g_xk : (k == idx ? g_w : 0.f);
    const float gbh = k < idx ? g_yk : (k == idx ? g_h : 0.f);
    gw[i] = cw * ew[i] * (gbw - dotw);  // out of range: every factor in brackets is 0
    gh[i] = ch * eh[i] * (gbh - doth);
    gs[i] = k == idx - 1 ? gs_lo : (k == idx ? gs_hi : 0.f);
  }
}

}  // namespace rqsdev
}  // namespace vms
