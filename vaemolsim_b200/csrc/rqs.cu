// rqs.cu -- K1: fused spline-bin search + rational-quadratic-spline evaluation with per-element log-det (sm_100a).
//
// Replaces flows.py:86-101 / :394-409 (softmax*scale+1e-2, softplus+1e-2 "activations") and the
// tfp.bijectors.RationalQuadraticSpline object built at flows.py:204-207 / :512-515 (forward, inverse, fldj,
// ildj = -fldj(inverse)), plus TF autodiff through both (backward kernel).
//
// Design (B200).  The op is HBM-bound: (3K-1) raw logits in, 2 scalars out per element (392 B @ K=32), so the kernel
// is built around coalesced 16-byte loads and a small instruction count per element:
//   * 8 lanes ("octet") cooperate on one element; each lane owns K/8 consecutive bins (4 bins @ K<=32: ONE float4
//     load per logit array per lane, 128 contiguous bytes per element, 4 elements per warp instruction);
//   * softmax max / sum and the cumulative sum are 3-step shuffles inside the octet (not 5-step warp scans), the
//     per-lane part is a 4-iteration sequential loop;
//   * no shared memory and ~64 registers => full occupancy, memory-level parallelism comes from resident warps;
//   * exactly one lane finds the bin; it evaluates the spline and writes y / log-det; it alone reads the TWO knot
//     slopes the element needs (the other K-3 raw slopes never leave HBM);
//   * prefix sums and knot positions are accumulated in float64 (B200 runs FP64 at half the FP32 rate): float32 knots
//     (what TFP computes) carry ~ulp(range) error that narrow bins amplify to 1e-4 in the log-det (parity).
#include "common.cuh"
#include <math.h>

namespace {

constexpr int kThreads = 256;
constexpr int kOct = 8;

struct RqsParams {
  int64_t n_rows;
  int n_dims;  // Dt: transformed dims per row, processed sequentially by the row's octet
  int K;
  float bin_min, bin_max, scale;  // scale = bin_max - bin_min - K*1e-2 (flows.py:92)
  const float* v_in;  int64_t ld_in;
  const float* raw_w; int64_t ld_w;
  const float* raw_h; int64_t ld_h;
  const float* raw_s; int64_t ld_s;
  float* v_out; int64_t ld_out;
  float* ldj;      // [n_rows, Dt] or null
  float* ldj_sum;  // [n_rows] or null
  int accumulate;
  int inverse_dir;
  // backward only
  const float* g_out; int64_t ld_g_out;
  const float* g_ldj_sum;
  float* g_in; int64_t ld_g_in;
  float* g_w; int64_t ld_gw;
  float* g_h; int64_t ld_gh;
  float* g_s; int64_t ld_gs;
};

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
template <int BPL, bool VEC>
__device__ __forceinline__ void load_bins(const float* __restrict__ row, int j, int K, float (&out)[BPL]) {
  if (VEC) {
#pragma unroll
    for (int q = 0; q < BPL / 4; ++q) {
      const int k0 = j * BPL + 4 * q;
      if (k0 < K) {  // K % 4 == 0 on the vector path: a 4-bin group is entirely valid or entirely padding
        const float4 t = __ldg(reinterpret_cast<const float4*>(row + k0));
        out[4 * q] = t.x; out[4 * q + 1] = t.y; out[4 * q + 2] = t.z; out[4 * q + 3] = t.w;
      } else {
        out[4 * q] = out[4 * q + 1] = out[4 * q + 2] = out[4 * q + 3] = -INFINITY;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < BPL; ++i) {
      const int k = j * BPL + i;
      out[i] = k < K ? __ldg(row + k) : -INFINITY;
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
  cwd = (double)scale / totw;
  chd = (double)scale / toth;
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

__device__ __forceinline__ void load_slopes(const float* __restrict__ ps, int idx, int K, float& s_lo, float& s_hi,
                                            float& dk, float& dk1) {
  s_lo = idx > 0 ? __ldg(ps + idx - 1) : 0.f;
  s_hi = idx < K - 1 ? __ldg(ps + idx) : 0.f;
  dk = idx == 0 ? 1.0f : vms::softplus_tf(s_lo) + 1e-2f;
  dk1 = idx == K - 1 ? 1.0f : vms::softplus_tf(s_hi) + 1e-2f;
}

template <int BPL, bool VEC>
__global__ void __launch_bounds__(kThreads) rqs_apply_kernel(const RqsParams p) {
  const int j = threadIdx.x & (kOct - 1);
  const int64_t oct0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kOct;
  const int64_t n_oct = (int64_t)gridDim.x * (kThreads / kOct);
  const int K = p.K;
  const bool inv = p.inverse_dir != 0;
  const int64_t n_iter = (p.n_rows + n_oct - 1) / n_oct;  // uniform trip count: shuffles need all lanes present
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row = oct0 + it * n_oct;
    const bool active = row < p.n_rows;
    const int64_t r = active ? row : p.n_rows - 1;  // inactive octets redo the last row, writes predicated off
    float ldj_acc = 0.f;
    for (int d = 0; d < p.n_dims; ++d) {
      float ew[BPL], eh[BPL];
      load_bins<BPL, VEC>(p.raw_w + r * p.ld_w + (int64_t)d * K, j, K, ew);
      load_bins<BPL, VEC>(p.raw_h + r * p.ld_h + (int64_t)d * K, j, K, eh);
      const float v = __ldg(p.v_in + r * p.ld_in + d);
      const double vd = (double)v;
      double cwd, chd, totw, toth;
      const Bin b = find_bin<BPL>(ew, eh, j, K, vd, inv, p.bin_min, p.scale, cwd, chd, totw, toth);
      float out = v, ldj = 0.f;
      if (b.found) {
        float s_lo, s_hi, dk, dk1;
        load_slopes(p.raw_s + r * p.ld_s + (int64_t)d * (K - 1), b.idx, K, s_lo, s_hi, dk, dk1);
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
      const bool writer = oct_mask ? b.found : (j == 0);  // out-of-range: lane 0 writes the identity
      if (writer && active) {
        p.v_out[r * p.ld_out + d] = out;
        if (p.ldj) p.ldj[r * p.n_dims + d] = ldj;
      }
      if (p.ldj_sum) {
        const int src = oct_mask ? (__ffs(oct_mask) - 1) : 0;
        ldj_acc += __shfl_sync(0xffffffffu, ldj, src, kOct);
      }
    }
    if (p.ldj_sum && j == 0 && active) {
      float* dst = p.ldj_sum + r;
      *dst = p.accumulate ? *dst + ldj_acc : ldj_acc;
    }
  }
}

template <int BPL, bool VEC>
__global__ void __launch_bounds__(kThreads) rqs_backward_kernel(const RqsParams p) {
  const int j = threadIdx.x & (kOct - 1);
  const int64_t oct0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kOct;
  const int64_t n_oct = (int64_t)gridDim.x * (kThreads / kOct);
  const int K = p.K;
  const bool inv = p.inverse_dir != 0;
  const int64_t n_iter = (p.n_rows + n_oct - 1) / n_oct;
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row = oct0 + it * n_oct;
    const bool active = row < p.n_rows;
    const int64_t r = active ? row : p.n_rows - 1;
    const float g_ldj = p.g_ldj_sum ? __ldg(p.g_ldj_sum + r) : 0.f;
    for (int d = 0; d < p.n_dims; ++d) {
      float ew[BPL], eh[BPL];
      load_bins<BPL, VEC>(p.raw_w + r * p.ld_w + (int64_t)d * K, j, K, ew);
      load_bins<BPL, VEC>(p.raw_h + r * p.ld_h + (int64_t)d * K, j, K, eh);
      const float v = __ldg(p.v_in + r * p.ld_in + d);
      const float g_out = __ldg(p.g_out + r * p.ld_g_out + d);
      const double vd = (double)v;
      double cwd, chd, totw, toth;
      const Bin b = find_bin<BPL>(ew, eh, j, K, vd, inv, p.bin_min, p.scale, cwd, chd, totw, toth);
      // owner-lane results, broadcast to the octet below
      float g_in = g_out, g_xk = 0.f, g_w = 0.f, g_yk = 0.f, g_h = 0.f, gs_lo = 0.f, gs_hi = 0.f, dotw = 0.f, doth = 0.f;
      if (b.found) {
        float s_lo, s_hi, dk, dk1;
        load_slopes(p.raw_s + r * p.ld_s + (int64_t)d * (K - 1), b.idx, K, s_lo, s_hi, dk, dk1);
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
        if (b.idx > 0) gs_lo = g_dk * vms::sigmoidf_(s_lo);
        if (b.idx < K - 1) gs_hi = g_dk1 * vms::sigmoidf_(s_hi);
        // softmax Jacobian dot products: sum_i p_i g_b[i] with g_b[i] = [i<idx] g_k + [i==idx] g_bin
        dotw = (g_xk * b.elt_w + g_w * b.e_w) * (float)(1.0 / totw);
        doth = (g_yk * b.elt_h + g_h * b.e_h) * (float)(1.0 / toth);
      }
      const unsigned found_mask = __ballot_sync(0xffffffffu, b.found);
      const unsigned oct_mask = (found_mask >> ((threadIdx.x & 31) & ~(kOct - 1))) & 0xffu;
      const int src = oct_mask ? (__ffs(oct_mask) - 1) : 0;
      if ((oct_mask ? b.found : (j == 0)) && active) p.g_in[r * p.ld_g_in + d] = g_in;
      // the shuffle must be executed by every lane of the warp (octets diverge on oct_mask): select afterwards
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
      float gw[BPL], gh[BPL], gs[BPL];
#pragma unroll
      for (int i = 0; i < BPL; ++i) {
        const int k = j * BPL + i;
        const float gbw = k < idx ? g_xk : (k == idx ? g_w : 0.f);
        const float gbh = k < idx ? g_yk : (k == idx ? g_h : 0.f);
        gw[i] = cw * ew[i] * (gbw - dotw);  // out of range: every factor in brackets is 0
        gh[i] = ch * eh[i] * (gbh - doth);
        gs[i] = k == idx - 1 ? gs_lo : (k == idx ? gs_hi : 0.f);
      }
      if (active) {
        store_bins<BPL, VEC>(p.g_w + r * p.ld_gw + (int64_t)d * K, j, K, gw);
        store_bins<BPL, VEC>(p.g_h + r * p.ld_gh + (int64_t)d * K, j, K, gh);
        store_bins<BPL, false>(p.g_s + r * p.ld_gs + (int64_t)d * (K - 1), j, K - 1, gs);
      }
    }
  }
}

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

vms_status configure(RqsParams& p, int& grid) {
  VMS_REQUIRE(p.K >= 2 && p.K <= 64, VMS_ERR_INVALID_ARG, "num_bins must be in [2, 64], got %d", p.K);
  VMS_REQUIRE(p.n_dims >= 1 && p.n_dims <= 4096, VMS_ERR_INVALID_ARG, "n_dims must be in [1, 4096], got %d", p.n_dims);
  VMS_REQUIRE(p.n_rows >= 0, VMS_ERR_INVALID_ARG, "n_rows < 0");
  VMS_REQUIRE(p.bin_max > p.bin_min, VMS_ERR_INVALID_ARG, "bin_range must be increasing");
  // flows.py:92: Python float arithmetic, then one cast to float32
  p.scale = (float)((double)p.bin_max - (double)p.bin_min - (double)p.K * 1e-2);
  VMS_REQUIRE(p.scale > 0.f, VMS_ERR_INVALID_ARG, "bin_range too narrow for %d bins", p.K);
  const int64_t n_blocks = (p.n_rows + (kThreads / kOct) - 1) / (kThreads / kOct);
  const int64_t cap = 16LL * vms::sm_count();
  grid = (int)(n_blocks < cap ? n_blocks : cap);
  return VMS_OK;
}

bool vector_ok(const RqsParams& p, bool backward) {
  if (p.K % 4) return false;
  bool ok = aligned16(p.raw_w) && aligned16(p.raw_h) && p.ld_w % 4 == 0 && p.ld_h % 4 == 0;
  if (backward) ok = ok && aligned16(p.g_w) && aligned16(p.g_h) && p.ld_gw % 4 == 0 && p.ld_gh % 4 == 0;
  return ok;
}

#define VMS_RQS_DISPATCH(KERNEL, name)                                                    \
  do {                                                                                    \
    if (p.n_rows == 0) return VMS_OK;                                                     \
    if (p.K <= 32) {                                                                      \
      if (vec) KERNEL<4, true><<<grid, kThreads, 0, st>>>(p);                             \
      else KERNEL<4, false><<<grid, kThreads, 0, st>>>(p);                                \
    } else {                                                                              \
      if (vec) KERNEL<8, true><<<grid, kThreads, 0, st>>>(p);                             \
      else KERNEL<8, false><<<grid, kThreads, 0, st>>>(p);                                \
    }                                                                                     \
    VMS_LAUNCH_CHECK(name);                                                               \
    return VMS_OK;                                                                        \
  } while (0)

vms_status run_apply(RqsParams p, cudaStream_t st) {
  int grid;
  vms_status s = configure(p, grid);
  if (s) return s;
  const bool vec = vector_ok(p, false);
  VMS_RQS_DISPATCH(rqs_apply_kernel, "rqs_apply_kernel");
}
vms_status run_backward(RqsParams p, cudaStream_t st) {
  int grid;
  vms_status s = configure(p, grid);
  if (s) return s;
  const bool vec = vector_ok(p, true);
  VMS_RQS_DISPATCH(rqs_backward_kernel, "rqs_backward_kernel");
}

RqsParams from_args(const vms_rqs_args& a) {
  RqsParams p = {};
  p.n_rows = a.n_rows; p.n_dims = a.n_dims; p.K = a.num_bins;
  p.bin_min = a.bin_min; p.bin_max = a.bin_max;
  p.v_in = a.v_in; p.ld_in = a.ld_in;
  p.raw_w = a.raw_w; p.ld_w = a.ld_w;
  p.raw_h = a.raw_h; p.ld_h = a.ld_h;
  p.raw_s = a.raw_s; p.ld_s = a.ld_s;
  p.v_out = a.v_out; p.ld_out = a.ld_out;
  p.ldj = a.ldj; p.ldj_sum = a.ldj_sum;
  p.accumulate = a.accumulate; p.inverse_dir = a.inverse_dir;
  return p;
}

}  // namespace

namespace vms {
vms_status rqs_prepare() { return VMS_OK; }  // no opt-in shared memory any more; kept for elbo.cu
}  // namespace vms

extern "C" {

vms_status vms_rqs_apply(const vms_rqs_args* a, vms_stream stream) {
  VMS_REQUIRE(a, VMS_ERR_INVALID_ARG, "args is NULL");
  VMS_REQUIRE(a->n_rows == 0 || (a->v_in && a->raw_w && a->raw_h && a->raw_s && a->v_out), VMS_ERR_INVALID_ARG,
              "vms_rqs_apply: NULL tensor pointer");
  return run_apply(from_args(*a), vms::as_stream(stream));
}

vms_status vms_rqs_apply_backward(const vms_rqs_bwd_args* a, vms_stream stream) {
  VMS_REQUIRE(a, VMS_ERR_INVALID_ARG, "args is NULL");
  RqsParams p = from_args(a->fwd);
  VMS_REQUIRE(p.n_rows == 0 || (p.v_in && p.raw_w && p.raw_h && p.raw_s && a->g_out && a->g_in && a->g_raw_w &&
                                a->g_raw_h && a->g_raw_s),
              VMS_ERR_INVALID_ARG, "vms_rqs_apply_backward: NULL tensor pointer");
  p.g_out = a->g_out; p.ld_g_out = a->ld_g_out;
  p.g_ldj_sum = a->g_ldj_sum;
  p.g_in = a->g_in; p.ld_g_in = a->ld_g_in;
  p.g_w = a->g_raw_w; p.ld_gw = a->ld_gw;
  p.g_h = a->g_raw_h; p.ld_gh = a->ld_gh;
  p.g_s = a->g_raw_s; p.ld_gs = a->ld_gs;
  return run_backward(p, vms::as_stream(stream));
}

static vms_status contiguous(const float* v, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                             float lo, float hi, float* out, float* ldj, int inverse_dir, vms_stream stream) {
  VMS_REQUIRE(n >= 0, VMS_ERR_INVALID_ARG, "n_elem < 0");
  VMS_REQUIRE(n == 0 || (v && rw && rh && rs && out), VMS_ERR_INVALID_ARG, "vms_rqs: NULL tensor pointer");
  vms_rqs_args a = {};
  a.n_rows = n; a.n_dims = 1; a.num_bins = K; a.bin_min = lo; a.bin_max = hi;
  a.v_in = v; a.ld_in = 1;
  a.raw_w = rw; a.ld_w = K; a.raw_h = rh; a.ld_h = K; a.raw_s = rs; a.ld_s = K - 1;
  a.v_out = out; a.ld_out = 1;
  a.ldj = ldj; a.ldj_sum = nullptr; a.accumulate = 0; a.inverse_dir = inverse_dir;
  return run_apply(from_args(a), vms::as_stream(stream));
}

vms_status vms_rqs_forward(const float* x, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                           float lo, float hi, float* y, float* ldj, vms_stream stream) {
  return contiguous(x, rw, rh, rs, n, K, lo, hi, y, ldj, 0, stream);
}
vms_status vms_rqs_inverse(const float* y, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                           float lo, float hi, float* x, float* ldj, vms_stream stream) {
  return contiguous(y, rw, rh, rs, n, K, lo, hi, x, ldj, 1, stream);
}

vms_status vms_rqs_backward(const float* v_in, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                            float lo, float hi, int inverse_dir, const float* g_out, const float* g_ldj, float* g_in,
                            float* g_rw, float* g_rh, float* g_rs, vms_stream stream) {
  VMS_REQUIRE(n >= 0, VMS_ERR_INVALID_ARG, "n_elem < 0");
  VMS_REQUIRE(n == 0 || (v_in && rw && rh && rs && g_out && g_in && g_rw && g_rh && g_rs), VMS_ERR_INVALID_ARG,
              "vms_rqs_backward: NULL tensor pointer");
  RqsParams p = {};
  p.n_rows = n; p.n_dims = 1; p.K = K; p.bin_min = lo; p.bin_max = hi;
  p.v_in = v_in; p.ld_in = 1;
  p.raw_w = rw; p.ld_w = K; p.raw_h = rh; p.ld_h = K; p.raw_s = rs; p.ld_s = K - 1;
  p.inverse_dir = inverse_dir;
  p.g_out = g_out; p.ld_g_out = 1;
  p.g_ldj_sum = g_ldj;
  p.g_in = g_in; p.ld_g_in = 1;
  p.g_w = g_rw; p.ld_gw = K; p.g_h = g_rh; p.ld_gh = K; p.g_s = g_rs; p.ld_gs = K - 1;
  return run_backward(p, vms::as_stream(stream));
}

}  // extern "C"
