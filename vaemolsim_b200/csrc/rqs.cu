// rqs.cu -- K1: fused spline-bin search + rational-quadratic-spline evaluation with per-element log-det (sm_100a).
//
// Replaces flows.py:86-101 / :394-409 (softmax*scale+1e-2, softplus+1e-2 "activations") and the
// tfp.bijectors.RationalQuadraticSpline object built at flows.py:204-207 / :512-515 (forward, inverse, fldj,
// ildj = -fldj(inverse)), plus TF autodiff through both (backward kernel).
//
// Design (B200): the op is HBM-bound -- (3K-1) raw logits in, 2 scalars out per element (392 B @ K=32).
//   * one THREAD per element, so the softmax / cumulative-sum / bin search are plain sequential loops with no
//     shuffles (a warp-per-element layout costs ~100 warp instructions per element and becomes issue-bound);
//   * logits are streamed HBM -> shared memory with cp.async (LDGSTS) in a 2-stage pipeline, warps copy whole
//     rows so global reads are 128-byte coalesced; the shared row stride is odd so thread-per-row reads are
//     bank-conflict free for any K;
//   * persistent grid: 2 CTAs per SM x SM count, grid-stride over tiles of ~128 elements;
//   * backward writes the (3K-1) logit gradients in place into the staged tile and stores them coalesced.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kStages = 2;

struct RqsParams {
  int64_t n_rows;
  int n_dims;   // Dt
  int K;
  int tile_rows;  // rows per tile
  int lds;        // shared row stride (floats), odd
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

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Stage one tile of raw logits: rows [row0, row0+rows) of the three arrays into smem rows of stride lds.
__device__ __forceinline__ void load_tile(const RqsParams& p, float* st, int64_t row0, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwh = p.n_dims * p.K, ns = p.n_dims * (p.K - 1);
  for (int r = warp; r < rows; r += kWarps) {
    float* dst = st + (size_t)r * p.lds;
    const float* gw = p.raw_w + (row0 + r) * p.ld_w;
    const float* gh = p.raw_h + (row0 + r) * p.ld_h;
    const float* gs = p.raw_s + (row0 + r) * p.ld_s;
    for (int c = lane; c < nwh; c += 32) cp_async4(dst + c, gw + c);
    for (int c = lane; c < nwh; c += 32) cp_async4(dst + nwh + c, gh + c);
    for (int c = lane; c < ns; c += 32) cp_async4(dst + 2 * nwh + c, gs + c);
  }
}

__device__ __forceinline__ void store_tile(const RqsParams& p, const float* st, int64_t row0, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwh = p.n_dims * p.K, ns = p.n_dims * (p.K - 1);
  for (int r = warp; r < rows; r += kWarps) {
    const float* src = st + (size_t)r * p.lds;
    float* gw = p.g_w + (row0 + r) * p.ld_gw;
    float* gh = p.g_h + (row0 + r) * p.ld_gh;
    float* gs = p.g_s + (row0 + r) * p.ld_gs;
    for (int c = lane; c < nwh; c += 32) gw[c] = src[c];
    for (int c = lane; c < nwh; c += 32) gh[c] = src[nwh + c];
    for (int c = lane; c < ns; c += 32) gs[c] = src[2 * nwh + c];
  }
}

// Per-element spline state after the bin search.
struct Bin {
  float xk, yk, wk, hk, sk, dk, dk1;
  float cw, ch;        // scale / sum(exp)
  float elt_w, elt_h;  // sum_{j<idx} exp_j (backward: softmax-Jacobian dot products)
  int idx;
  bool oob;
};

// Softmax statistics + sequential knot walk.  pw / ph are overwritten with exp(logit - max).
template <int KT>
__device__ __forceinline__ Bin find_bin(float* pw, float* ph, const float* ps, int Krt, float v, bool inverse_dir,
                                        float bin_min, float scale) {
  const int K = KT ? KT : Krt;
  Bin b;
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    mw = fmaxf(mw, pw[k]);
    mh = fmaxf(mh, ph[k]);
  }
  float sw = 0.f, sh = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    float ew = expf(pw[k] - mw), eh = expf(ph[k] - mh);
    sw += ew;
    sh += eh;
    pw[k] = ew;
    ph[k] = eh;
  }
  b.cw = scale / sw;
  b.ch = scale / sh;
  // knots: kx[0] = bin_min, kx[k+1] = cumsum(bw)[k] + bin_min (TFP _knot_positions); same for ky.
  float cx = 0.f, cy = 0.f, ex = 0.f, ey = 0.f;
  float cx_i = 0.f, cy_i = 0.f, ex_i = 0.f, ey_i = 0.f;
  int idx = 0;
#pragma unroll 8
  for (int k = 0; k < K; ++k) {
    float knot = (inverse_dir ? cy : cx) + bin_min;
    if (k == 0 || v >= knot) {  // largest k with knot_k <= v, floored at 0 (searchsorted 'right' - 1)
      idx = k;
      cx_i = cx; cy_i = cy; ex_i = ex; ey_i = ey;
    }
    float ew = pw[k], eh = ph[k];
    cx += fmaf(ew, b.cw, 1e-2f);
    cy += fmaf(eh, b.ch, 1e-2f);
    ex += ew;
    ey += eh;
  }
  float vmax = (inverse_dir ? cy : cx) + bin_min;
  b.oob = (v <= bin_min) || (v >= vmax);
  if (b.oob) { idx = 0; cx_i = cy_i = ex_i = ey_i = 0.f; }
  b.idx = idx;
  b.elt_w = ex_i;
  b.elt_h = ey_i;
  float bw = fmaf(pw[idx], b.cw, 1e-2f), bh = fmaf(ph[idx], b.ch, 1e-2f);
  b.xk = cx_i + bin_min;
  b.yk = cy_i + bin_min;
  b.wk = ((cx_i + bw) + bin_min) - b.xk;
  b.hk = ((cy_i + bh) + bin_min) - b.yk;
  b.sk = b.hk / b.wk;
  b.dk = idx == 0 ? 1.0f : vms::softplus_tf(ps[idx - 1]) + 1e-2f;
  b.dk1 = idx == K - 1 ? 1.0f : vms::softplus_tf(ps[idx]) + 1e-2f;
  return b;
}

// relative position r in the bin for either direction (TFP _forward / _inverse)
__device__ __forceinline__ float rel_pos(const Bin& b, float v, bool inverse_dir) {
  if (!inverse_dir) return (v - b.xk) / b.wk;
  float ry = v - b.yk;
  float t2 = ry * (b.dk1 + b.dk - 2.f * b.sk);
  float a = b.hk * (b.sk - b.dk) + t2;
  float bb = b.hk * b.dk - t2;
  float c = -b.sk * ry;
  float disc = bb * bb - 4.f * a * c;
  float r = (2.f * c) / (-bb - sqrtf(disc));
  return ry == 0.f ? 0.f : r;
}

template <int KT>
__global__ void __launch_bounds__(kThreads) rqs_apply_kernel(const RqsParams p) {
  extern __shared__ float smem[];
  const int stage_floats = p.tile_rows * p.lds;
  float* ldj_s = smem + kStages * stage_floats;  // [tile_rows * Dt] scratch for the event sum
  const int64_t n_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;
  const int K = KT ? KT : p.K;
  const int nwh = p.n_dims * K;
  const bool inv = p.inverse_dir != 0;

  int64_t tile = blockIdx.x;
  int stage = 0;
  if (tile < n_tiles) {
    int64_t row0 = tile * p.tile_rows;
    int rows = (int)min((int64_t)p.tile_rows, p.n_rows - row0);
    load_tile(p, smem, row0, rows);
  }
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * p.tile_rows;
    const int rows = (int)min((int64_t)p.tile_rows, p.n_rows - row0);
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      int64_t nrow0 = next * p.tile_rows;
      int nrows = (int)min((int64_t)p.tile_rows, p.n_rows - nrow0);
      load_tile(p, smem + (stage ^ 1) * stage_floats, nrow0, nrows);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float* st = smem + stage * stage_floats;
    const int n_el = rows * p.n_dims;
    for (int e = threadIdx.x; e < n_el; e += kThreads) {
      const int r = e / p.n_dims, d = e - r * p.n_dims;
      float* row = st + (size_t)r * p.lds;
      const float v = p.v_in[(row0 + r) * p.ld_in + d];
      Bin b = find_bin<KT>(row + d * K, row + nwh + d * K, row + 2 * nwh + d * (K - 1), K, v, inv, p.bin_min, p.scale);
      float out = v, ldj = 0.f;
      if (!b.oob) {
        float rr = rel_pos(b, v, inv);
        float omr = 1.f - rr, u = rr * omr;
        float den = b.sk + (b.dk1 + b.dk - 2.f * b.sk) * u;
        if (!inv) {
          float num = b.hk * (b.sk * rr * rr + b.dk * u);
          out = b.yk + num / den;
        } else {
          out = rr * b.wk + b.xk;
        }
        float P = b.dk1 * rr * rr + 2.f * b.sk * u + b.dk * omr * omr;
        ldj = logf((b.sk * b.sk) * P / (den * den));
        if (inv) ldj = -ldj;
      }
      p.v_out[(row0 + r) * p.ld_out + d] = out;
      if (p.ldj) p.ldj[(row0 + r) * p.n_dims + d] = ldj;
      if (p.ldj_sum) ldj_s[e] = ldj;
    }
    __syncthreads();  // tile consumed (and ldj_s complete) before the stage is refilled
    if (p.ldj_sum) {
      for (int r = threadIdx.x; r < rows; r += kThreads) {
        float s = 0.f;
        for (int d = 0; d < p.n_dims; ++d) s += ldj_s[r * p.n_dims + d];
        float* dst = p.ldj_sum + row0 + r;
        *dst = p.accumulate ? *dst + s : s;
      }
      // ldj_s is rewritten only after the next iteration's first __syncthreads
    }
    stage ^= 1;
  }
  cp_async_wait<0>();
}

template <int KT>
__global__ void __launch_bounds__(kThreads) rqs_backward_kernel(const RqsParams p) {
  extern __shared__ float smem[];
  const int stage_floats = p.tile_rows * p.lds;
  const int64_t n_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;
  const int K = KT ? KT : p.K;
  const int nwh = p.n_dims * K;
  const bool inv = p.inverse_dir != 0;

  int64_t tile = blockIdx.x;
  int stage = 0;
  if (tile < n_tiles) {
    int64_t row0 = tile * p.tile_rows;
    int rows = (int)min((int64_t)p.tile_rows, p.n_rows - row0);
    load_tile(p, smem, row0, rows);
  }
  cp_async_commit();
  for (; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * p.tile_rows;
    const int rows = (int)min((int64_t)p.tile_rows, p.n_rows - row0);
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) {
      int64_t nrow0 = next * p.tile_rows;
      int nrows = (int)min((int64_t)p.tile_rows, p.n_rows - nrow0);
      load_tile(p, smem + (stage ^ 1) * stage_floats, nrow0, nrows);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float* st = smem + stage * stage_floats;
    const int n_el = rows * p.n_dims;
    for (int e = threadIdx.x; e < n_el; e += kThreads) {
      const int r = e / p.n_dims, d = e - r * p.n_dims;
      float* row = st + (size_t)r * p.lds;
      float* pw = row + d * K;
      float* ph = row + nwh + d * K;
      float* ps = row + 2 * nwh + d * (K - 1);
      const float v = p.v_in[(row0 + r) * p.ld_in + d];
      const float g_out = p.g_out[(row0 + r) * p.ld_g_out + d];
      const float g_ldj = p.g_ldj_sum ? p.g_ldj_sum[row0 + r] : 0.f;
      Bin b = find_bin<KT>(pw, ph, ps, K, v, inv, p.bin_min, p.scale);
      float g_in = g_out;
      float g_xk = 0.f, g_w = 0.f, g_yk = 0.f, g_h = 0.f, g_dk = 0.f, g_dk1 = 0.f;
      if (!b.oob) {
        // local derivatives: y = yk + h N/Q, L = log(s^2 P / Q^2)
        const float rr = rel_pos(b, v, inv);
        const float s = b.sk, dk = b.dk, dk1 = b.dk1, h = b.hk, w = b.wk;
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
        g_dk = gy * y_dk + gL * L_dk;
        g_dk1 = gy * y_dk1 + gL * L_dk1;
        g_h = gy * y_h + g_s * iw;
        g_yk = gy;
        g_w = -(g_s * s + g_r * rr) * iw;
        g_xk = -g_r * iw;
      }
      p.g_in[(row0 + r) * p.ld_g_in + d] = g_in;
      // softmax Jacobian: g_raw[j] = c e_j (g_b[j] - sum_i p_i g_b[i]),  g_b[j] = [j<idx] g_k + [j==idx] g_bin
      const int idx = b.idx;
      const float ew_i = pw[idx], eh_i = ph[idx];
      const float dotw = (g_xk * b.elt_w + g_w * ew_i) * (b.cw / p.scale);
      const float doth = (g_yk * b.elt_h + g_h * eh_i) * (b.ch / p.scale);
      const float s_lo = idx > 0 ? ps[idx - 1] : 0.f, s_hi = idx < K - 1 ? ps[idx] : 0.f;
#pragma unroll 8
      for (int k = 0; k < K; ++k) {
        float gbw = k < idx ? g_xk : (k == idx ? g_w : 0.f);
        float gbh = k < idx ? g_yk : (k == idx ? g_h : 0.f);
        pw[k] = b.cw * pw[k] * (gbw - dotw);
        ph[k] = b.ch * ph[k] * (gbh - doth);
      }
#pragma unroll 8
      for (int k = 0; k < K - 1; ++k) ps[k] = 0.f;
      if (!b.oob) {
        if (idx > 0) ps[idx - 1] = g_dk * vms::sigmoidf_(s_lo);
        if (idx < K - 1) ps[idx] = g_dk1 * vms::sigmoidf_(s_hi);
      }
    }
    __syncthreads();
    store_tile(p, st, row0, rows);
    __syncthreads();  // stores read the stage; it is refilled by the next iteration's prefetch
    stage ^= 1;
  }
  cp_async_wait<0>();
}

vms_status configure(RqsParams& p, size_t& smem_bytes, int& grid, bool need_ldj_scratch) {
  VMS_REQUIRE(p.K >= 2 && p.K <= 64, VMS_ERR_INVALID_ARG, "num_bins must be in [2, 64], got %d", p.K);
  VMS_REQUIRE(p.n_dims >= 1 && p.n_dims <= 128, VMS_ERR_INVALID_ARG, "n_dims must be in [1, 128], got %d", p.n_dims);
  VMS_REQUIRE(p.n_rows >= 0, VMS_ERR_INVALID_ARG, "n_rows < 0");
  VMS_REQUIRE(p.bin_max > p.bin_min, VMS_ERR_INVALID_ARG, "bin_range must be increasing");
  // flows.py:92: Python float arithmetic, then one cast to float32
  p.scale = (float)((double)p.bin_max - (double)p.bin_min - (double)p.K * 1e-2);
  VMS_REQUIRE(p.scale > 0.f, VMS_ERR_INVALID_ARG, "bin_range too narrow for %d bins", p.K);
  const int width = p.n_dims * (3 * p.K - 1);
  p.lds = width | 1;
  int tr = kThreads / p.n_dims;
  if (tr < 1) tr = 1;
  const size_t budget = 96 * 1024;  // two stages <= 96 KB => 2 CTAs / SM
  while (tr > 1 && (size_t)kStages * tr * p.lds * 4 + (size_t)tr * p.n_dims * 4 > budget) tr >>= 1;
  p.tile_rows = tr;
  smem_bytes = (size_t)kStages * tr * p.lds * 4 + (need_ldj_scratch ? (size_t)tr * p.n_dims * 4 : 0);
  VMS_REQUIRE(smem_bytes <= (size_t)vms::max_smem_optin(), VMS_ERR_UNSUPPORTED,
              "RQS tile needs %zu bytes of shared memory", smem_bytes);
  const int64_t n_tiles = (p.n_rows + tr - 1) / tr;
  const int64_t cap = 2LL * vms::sm_count();
  grid = (int)(n_tiles < cap ? n_tiles : cap);
  return VMS_OK;
}

// Opt in to > 48 KB of dynamic shared memory once per kernel instantiation (never inside a graph capture after
// vms::rqs_prepare() has run).
template <typename KernelT>
vms_status prepare_kernel(KernelT kern) {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!done[dev]) {
    VMS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, vms::max_smem_optin()));
    done[dev] = true;
  }
  return VMS_OK;
}

template <typename KernelT>
vms_status launch(KernelT kern, const RqsParams& p, size_t smem_bytes, int grid, cudaStream_t st, const char* name) {
  if (p.n_rows == 0) return VMS_OK;
  vms_status s = prepare_kernel(kern);
  if (s) return s;
  kern<<<grid, kThreads, smem_bytes, st>>>(p);
  VMS_LAUNCH_CHECK(name);
  return VMS_OK;
}

vms_status run_apply(RqsParams p, cudaStream_t st) {
  size_t smem;
  int grid;
  vms_status s = configure(p, smem, grid, p.ldj_sum != nullptr);
  if (s) return s;
  if (p.K == 32) return launch(rqs_apply_kernel<32>, p, smem, grid, st, "rqs_apply<32>");
  if (p.K == 20) return launch(rqs_apply_kernel<20>, p, smem, grid, st, "rqs_apply<20>");
  return launch(rqs_apply_kernel<0>, p, smem, grid, st, "rqs_apply<K>");
}
vms_status run_backward(RqsParams p, cudaStream_t st) {
  size_t smem;
  int grid;
  vms_status s = configure(p, smem, grid, false);
  if (s) return s;
  if (p.K == 32) return launch(rqs_backward_kernel<32>, p, smem, grid, st, "rqs_backward<32>");
  if (p.K == 20) return launch(rqs_backward_kernel<20>, p, smem, grid, st, "rqs_backward<20>");
  return launch(rqs_backward_kernel<0>, p, smem, grid, st, "rqs_backward<K>");
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
vms_status rqs_prepare() {
  vms_status s;
  if ((s = prepare_kernel(rqs_apply_kernel<32>))) return s;
  if ((s = prepare_kernel(rqs_apply_kernel<20>))) return s;
  if ((s = prepare_kernel(rqs_apply_kernel<0>))) return s;
  if ((s = prepare_kernel(rqs_backward_kernel<32>))) return s;
  if ((s = prepare_kernel(rqs_backward_kernel<20>))) return s;
  return prepare_kernel(rqs_backward_kernel<0>);
}
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
