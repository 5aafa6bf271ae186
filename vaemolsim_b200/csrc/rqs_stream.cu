// rqs_stream.cu -- K1, streaming form: the contiguous RQS op (x [n], raw_w / raw_h [n, K], raw_s [n, K-1]) at HBM speed.
//
// Same arithmetic and the same reference lines as rqs.cu / rqs_device.cuh (flows.py:86-101, :204-207; TFP
// RationalQuadraticSpline forward / inverse / log-det and their reverse mode).  What changes is the mapping.
//
// Why a second kernel.  ncu on the octet kernel of rqs.cu (8 lanes per element) at n = 2 M, K = 32: issue-active 68 %,
// 170 warp-instructions per element, DRAM at 20-26 % of the measured copy peak -- it is ISSUE-bound: the bin search is
// cooperative, but everything after it (the quadratic, the log-det, divisions, softplus of two slopes) runs with one
// useful lane in eight.  Here ONE THREAD owns one element, so every lane does useful work (~30 warp-instructions per
// element), and coalescing is recovered by staging:
//   * a CTA of 128 threads takes tiles of 128 consecutive elements; the tile's width / height logits (32 KB @ K = 32)
//     are copied global -> shared with 16-byte cp.async (fully coalesced, no registers).  One stage per CTA and
//     4-5 CTAs per SM: while one CTA waits for its tile the others evaluate theirs (16-20 warps / SM);
//   * rows are stored with a pitch of K + 4 floats: a thread reads ITS row with 128-bit loads and a quarter-warp touches
//     all 32 banks (144-byte pitch) -- conflict-free;
//   * nothing per-bin lives in registers: the thread makes three passes over its row (max; sums of exps per group of
//     4 bins; knot walk over group boundaries, then inside one group), recomputing the few exps it needs with
//     ex2.approx instead of keeping 2K exps alive -- ~96 registers instead of 255, which is what buys the occupancy;
//     prefix sums / knots are accumulated in float64 as in rqs_device.cuh;
//   * only the TWO slopes of the bin are read, straight from global memory (1-2 sectors per element instead of the
//     124-byte row: the other K - 3 raw slopes never leave HBM);
//   * the backward kernel writes its gradients back into the thread's own shared-memory rows (slope gradients into a
//     third tile) and the CTA copies the tile out with coalesced 16-byte stores.
// exp: this file uses ex2.approx(x log2 e) (2 ulp) for every exp of the softmax, consistently in all passes; the
// octet kernel uses expf.  The two agree to ~3e-7 relative, far inside the parity tolerance (tests/test_gpu_kernels.py).
// Anything that is not the contiguous, 16-byte-aligned form with K in {20, 32} (strided coupling-layer views, other
// K) stays on the octet kernel.
#include "common.cuh"
#include <math.h>

namespace vms {

namespace {

constexpr int ST = 128;  // threads per CTA = elements per tile

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template <int K>
struct Stage {
  static constexpr int PITCH = K + 4;
  float w[ST * PITCH];
  float h[ST * PITCH];
  float x[ST];
  float g[2 * ST];  // backward: g_out | g_ldj
};
template <int K>
struct StageBwd : Stage<K> {
  float s[ST * (K - 1) + 4];  // slope gradients of the tile (contiguous block in global memory)
};

struct StreamParams {
  const float *x, *raw_w, *raw_h, *raw_s;
  int64_t n;
  float bin_min, scale;
  float *y, *ldj;
  // backward
  const float *g_out, *g_ldj;
  float *g_in, *g_w, *g_h, *g_s;
};

__device__ __forceinline__ float fast_exp(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
  return r;
}

// Issue the copies of one tile into the stage.  Full tiles use 16-byte cp.async; the ragged last tile uses plain loads
// (zero-filled rows past the end keep the arithmetic finite).
template <int K, bool BWD>
__device__ __forceinline__ void load_tile(Stage<K>& st, const StreamParams& p, int64_t e0) {
  const int tid = threadIdx.x;
  const int64_t left = p.n - e0;
  if (left >= ST) {
    constexpr int CPR = K / 4;  // 16-byte chunks per width / height row
#pragma unroll
    for (int i = 0; i < CPR; ++i) {
      const int c = tid + ST * i, row = c / CPR, q = c - row * CPR;
      cp_async16(st.w + row * Stage<K>::PITCH + 4 * q, p.raw_w + (e0 + row) * K + 4 * q);
      cp_async16(st.h + row * Stage<K>::PITCH + 4 * q, p.raw_h + (e0 + row) * K + 4 * q);
    }
    if (tid < ST / 4) cp_async16(st.x + 4 * tid, p.x + e0 + 4 * tid);
    if (BWD) {
      if (tid >= ST / 4 && tid < ST / 2) cp_async16(st.g + 4 * (tid - ST / 4), p.g_out + e0 + 4 * (tid - ST / 4));
      if (p.g_ldj && tid >= ST / 2 && tid < 3 * ST / 4)
        cp_async16(st.g + ST + 4 * (tid - ST / 2), p.g_ldj + e0 + 4 * (tid - ST / 2));
    }
  } else {
    const int nv = (int)left;
    for (int c = tid; c < ST * K; c += ST) {
      const int row = c / K, k = c - row * K;
      st.w[row * Stage<K>::PITCH + k] = row < nv ? __ldg(p.raw_w + (e0 + row) * K + k) : 0.f;
      st.h[row * Stage<K>::PITCH + k] = row < nv ? __ldg(p.raw_h + (e0 + row) * K + k) : 0.f;
    }
    st.x[tid] = tid < nv ? __ldg(p.x + e0 + tid) : 0.f;
    if (BWD) {
      st.g[tid] = tid < nv ? __ldg(p.g_out + e0 + tid) : 0.f;
      st.g[ST + tid] = (p.g_ldj && tid < nv) ? __ldg(p.g_ldj + e0 + tid) : 0.f;
    }
  }
}

// What one thread knows about its element after the softmax / bin search (no per-bin arrays).
struct Elem {
  float mw, mh;           // row maxima
  double totw, toth;      // sums of exps
  double cw, ch;          // scale / total
  int idx;                // bin, -1 when out of range
  double lo_x, lo_y;      // lower knots of the bin
  float wk, hk;           // bin width / height
  float e_w, e_h;         // exps of the bin
  float elt_w, elt_h;     // sums of exps of lower bins
};

// float -> double of a non-negative, normal-or-zero float (every softmax exp is one) with three integer instructions
// instead of F2F.F64.F32: the conversion shares the quarter-rate XU pipe with MUFU.EX2, and ncu showed that pipe 70 %
// busy -- the bottleneck -- when both were on it.  (+0 maps to 2^-127 instead of 0: irrelevant in sums that are >= 1.)
__device__ __forceinline__ double f2d_pos(float f) {
  return __longlong_as_double(((long long)__float_as_int(f) << 29) + (896LL << 52));
}

template <int K, bool INV>
__device__ __forceinline__ void search(const float* __restrict__ wrow, const float* __restrict__ hrow, float v, float bin_min,
                                       float scale, Elem& E) {
  constexpr int G = K / 4;  // groups of 4 bins = one 128-bit shared-memory load
  // pass 1: maxima
  float mw = -INFINITY, mh = -INFINITY;
#pragma unroll
  for (int q = 0; q < G; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(wrow + 4 * q);
    const float4 b = *reinterpret_cast<const float4*>(hrow + 4 * q);
    mw = fmaxf(fmaxf(fmaxf(mw, a.x), fmaxf(a.y, a.z)), a.w);
    mh = fmaxf(fmaxf(fmaxf(mh, b.x), fmaxf(b.y, b.z)), b.w);
  }
  // (compiler barriers between the passes: without them the row loads are CSE'd across passes and the 2K logits stay
  // live in registers -- the very thing the multi-pass form is there to avoid)
  asm volatile("" ::: "memory");
  // pass 2: sums of exps per group of 4 bins and in total (float64 accumulation, fixed order)
  double gw[G], gh[G];
  double tw = 0.0, th = 0.0;
#pragma unroll
  for (int q = 0; q < G; ++q) {
    const float4 a = *reinterpret_cast<const float4*>(wrow + 4 * q);
    const float4 b = *reinterpret_cast<const float4*>(hrow + 4 * q);
    gw[q] = ((f2d_pos(fast_exp(a.x - mw)) + f2d_pos(fast_exp(a.y - mw))) + f2d_pos(fast_exp(a.z - mw))) +
            f2d_pos(fast_exp(a.w - mw));
    gh[q] = ((f2d_pos(fast_exp(b.x - mh)) + f2d_pos(fast_exp(b.y - mh))) + f2d_pos(fast_exp(b.z - mh))) +
            f2d_pos(fast_exp(b.w - mh));
    tw += gw[q];
    th += gh[q];
  }
  E.mw = mw; E.mh = mh; E.totw = tw; E.toth = th;
  E.cw = (double)scale / tw;
  E.ch = (double)scale / th;
  // pass 3: knot walk on the searched axis, first over group boundaries, then inside the group.
  // k-th upper knot = bin_min + (scale * E_{k+1} / total + 1e-2 (k+1)), E_k = sum of exps of bins < k (rqs_device.cuh);
  // knots are increasing, so the bins whose upper knot is <= v are exactly the bins below the one containing v.
  const double vd = (double)v, bm = (double)bin_min;
  const float* srow = INV ? hrow : wrow;
  const float* orow = INV ? wrow : hrow;
  const float ms = INV ? mh : mw, mo = INV ? mw : mh;
  const double cs = INV ? E.ch : E.cw, ts = INV ? th : tw;
  double run = 0.0, Ps = 0.0, Po = 0.0;
  int gcnt = 0;
#pragma unroll
  for (int q = 0; q < G; ++q) {
    run += INV ? gh[q] : gw[q];
    const double hi = bm + fma(cs, q == G - 1 ? ts : run, 1e-2 * (double)(4 * q + 4));
    if (vd >= hi) {
      ++gcnt;
      Ps = run;
      Po += INV ? gw[q] : gh[q];
    }
  }
  int cnt = 4 * gcnt;
  if (gcnt < G) {
    const float4 a = *reinterpret_cast<const float4*>(srow + 4 * gcnt);
    const float4 b = *reinterpret_cast<const float4*>(orow + 4 * gcnt);
    const float as[4] = {a.x, a.y, a.z, a.w}, bo[4] = {b.x, b.y, b.z, b.w};
    run = Ps;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = 4 * gcnt + i;
      run += f2d_pos(fast_exp(as[i] - ms));
      const double hi = bm + fma(cs, k == K - 1 ? ts : run, 1e-2 * (double)(k + 1));
      if (vd >= hi) {
        ++cnt;
        Ps = run;
        Po += f2d_pos(fast_exp(bo[i] - mo));
      }
    }
  }
  // TFP: x <= kx[0] or x >= kx[K] => identity
  E.idx = (vd > bm && cnt < K) ? cnt : -1;
  const int i = E.idx < 0 ? 0 : E.idx;
  if (E.idx < 0) { Ps = 0.0; Po = 0.0; }
  const float es = fast_exp(srow[i] - ms), eo = fast_exp(orow[i] - mo);
  const double co = INV ? E.cw : E.ch, to = INV ? tw : th;
  const double lo_s = bm + fma(cs, Ps, 1e-2 * (double)i), lo_o = bm + fma(co, Po, 1e-2 * (double)i);
  const double hi_s = bm + fma(cs, i == K - 1 ? ts : Ps + f2d_pos(es), 1e-2 * (double)(i + 1));
  const double hi_o = bm + fma(co, i == K - 1 ? to : Po + f2d_pos(eo), 1e-2 * (double)(i + 1));
  E.lo_x = INV ? lo_o : lo_s;
  E.lo_y = INV ? lo_s : lo_o;
  E.wk = (float)(INV ? hi_o - lo_o : hi_s - lo_s);
  E.hk = (float)(INV ? hi_s - lo_s : hi_o - lo_o);
  E.e_w = INV ? eo : es;
  E.e_h = INV ? es : eo;
  E.elt_w = (float)(INV ? Po : Ps);
  E.elt_h = (float)(INV ? Ps : Po);
}

// relative position r in the bin for either direction (TFP _forward / _inverse); identical to rqs_device.cuh
__device__ __forceinline__ float rel_pos_s(double vd, double lo_x, double lo_y, float wk, float hk, float sk, float dk,
                                           float dk1, bool inv) {
  if (!inv) return (float)(vd - lo_x) / wk;
  const float ry = (float)(vd - lo_y);
  const float t2 = ry * (dk1 + dk - 2.f * sk);
  const float a = hk * (sk - dk) + t2;
  const float bb = hk * dk - t2;
  const float c = -sk * ry;
  const float disc = bb * bb - 4.f * a * c;
  const float r = (2.f * c) / (-bb - sqrtf(disc));
  return ry == 0.f ? 0.f : r;
}

template <int K, bool INV>
__global__ void __launch_bounds__(ST, 5) rqs_stream_apply_kernel(const StreamParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  Stage<K>& st = *reinterpret_cast<Stage<K>*>(smraw);
  const int tid = threadIdx.x;
  const int64_t n_tiles = (p.n + ST - 1) / ST;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    load_tile<K, false>(st, p, tile * ST);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int64_t e = tile * ST + tid;
    const float v = st.x[tid];
    Elem E;
    search<K, INV>(st.w + tid * Stage<K>::PITCH, st.h + tid * Stage<K>::PITCH, v, p.bin_min, p.scale, E);
    float out = v, ldj = 0.f;
    if (E.idx >= 0 && e < p.n) {
      const float* ps = p.raw_s + e * (K - 1);
      const float s_lo = E.idx > 0 ? __ldg(ps + E.idx - 1) : 0.f;
      const float s_hi = E.idx < K - 1 ? __ldg(ps + E.idx) : 0.f;
      const float dk = E.idx == 0 ? 1.0f : softplus_tf(s_lo) + 1e-2f;
      const float dk1 = E.idx == K - 1 ? 1.0f : softplus_tf(s_hi) + 1e-2f;
      const double vd = (double)v;
      const float sk = E.hk / E.wk;
      const float rr = rel_pos_s(vd, E.lo_x, E.lo_y, E.wk, E.hk, sk, dk, dk1, INV);
      const float omr = 1.f - rr, u = rr * omr;
      const float den = sk + (dk1 + dk - 2.f * sk) * u;
      if (!INV) {
        const float num = E.hk * (sk * rr * rr + dk * u);
        out = (float)(E.lo_y + (double)(num / den));
      } else {
        out = (float)(E.lo_x + (double)(rr * E.wk));
      }
      const float P = dk1 * rr * rr + 2.f * sk * u + dk * omr * omr;
      ldj = logf((sk * sk) * P / (den * den));
      if (INV) ldj = -ldj;
    }
    if (e < p.n) {
      p.y[e] = out;
      if (p.ldj) p.ldj[e] = ldj;
    }
    __syncthreads();  // the stage is refilled at the top of the loop
  }
}

template <int K, bool INV>
__global__ void __launch_bounds__(ST, 4) rqs_stream_backward_kernel(const StreamParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  StageBwd<K>& st = *reinterpret_cast<StageBwd<K>*>(smraw);
  const int tid = threadIdx.x;
  const int64_t n_tiles = (p.n + ST - 1) / ST;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    load_tile<K, true>(st, p, tile * ST);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int64_t e0 = tile * ST, e = e0 + tid;
    const float v = st.x[tid];
    const float g_out = st.g[tid], g_ldj = p.g_ldj ? st.g[ST + tid] : 0.f;
    float* wrow = st.w + tid * Stage<K>::PITCH;
    float* hrow = st.h + tid * Stage<K>::PITCH;
    float* srow = st.s + tid * (K - 1);
    Elem E;
    search<K, INV>(wrow, hrow, v, p.bin_min, p.scale, E);
    float g_in = g_out, g_xk = 0.f, g_w = 0.f, g_yk = 0.f, g_h = 0.f, gs_lo = 0.f, gs_hi = 0.f, dotw = 0.f, doth = 0.f;
    if (E.idx >= 0 && e < p.n) {
      const float* ps = p.raw_s + e * (K - 1);
      const float s_lo = E.idx > 0 ? __ldg(ps + E.idx - 1) : 0.f;
      const float s_hi = E.idx < K - 1 ? __ldg(ps + E.idx) : 0.f;
      const float dk = E.idx == 0 ? 1.0f : softplus_tf(s_lo) + 1e-2f;
      const float dk1 = E.idx == K - 1 ? 1.0f : softplus_tf(s_hi) + 1e-2f;
      // local derivatives: y = yk + h N/Q, L = log(s^2 P / Q^2); same expressions as rqs_device.cuh
      const float h = E.hk, w = E.wk, s = h / w;
      const float rr = rel_pos_s((double)v, E.lo_x, E.lo_y, w, h, s, dk, dk1, INV);
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
      if (INV) {
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
      if (E.idx > 0) gs_lo = g_dk * sigmoidf_(s_lo);
      if (E.idx < K - 1) gs_hi = g_dk1 * sigmoidf_(s_hi);
      dotw = (g_xk * E.elt_w + g_w * E.e_w) * (float)(1.0 / E.totw);
      doth = (g_yk * E.elt_h + g_h * E.e_h) * (float)(1.0 / E.toth);
    }
    // pass 4: raw-logit gradients overwrite the thread's own rows (softmax Jacobian; out of range => all zero)
    asm volatile("" ::: "memory");
    const int idx = E.idx < 0 ? -2 : E.idx;
    const float cw = (float)E.cw, ch = (float)E.ch;
#pragma unroll 2
    for (int q = 0; q < K / 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(wrow + 4 * q);
      const float4 b = *reinterpret_cast<const float4*>(hrow + 4 * q);
      const float aw[4] = {a.x, a.y, a.z, a.w}, bh[4] = {b.x, b.y, b.z, b.w};
      float gw[4], gh[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = 4 * q + i;
        const float gbw = k < idx ? g_xk : (k == idx ? g_w : 0.f);
        const float gbh = k < idx ? g_yk : (k == idx ? g_h : 0.f);
        gw[i] = cw * fast_exp(aw[i] - E.mw) * (gbw - dotw);
        gh[i] = ch * fast_exp(bh[i] - E.mh) * (gbh - doth);
      }
      *reinterpret_cast<float4*>(wrow + 4 * q) = make_float4(gw[0], gw[1], gw[2], gw[3]);
      *reinterpret_cast<float4*>(hrow + 4 * q) = make_float4(gh[0], gh[1], gh[2], gh[3]);
    }
#pragma unroll
    for (int k = 0; k < K - 1; ++k) srow[k] = k == idx - 1 ? gs_lo : (k == idx ? gs_hi : 0.f);
    if (e < p.n) p.g_in[e] = g_in;
    __syncthreads();
    // coalesced copy-out of the tile's gradients
    const int64_t left = p.n - e0;
    if (left >= ST) {
      constexpr int CPR = K / 4;
#pragma unroll
      for (int i = 0; i < CPR; ++i) {
        const int c = tid + ST * i, row = c / CPR, q = c - row * CPR;
        *reinterpret_cast<float4*>(p.g_w + (e0 + row) * K + 4 * q) =
            *reinterpret_cast<const float4*>(st.w + row * Stage<K>::PITCH + 4 * q);
        *reinterpret_cast<float4*>(p.g_h + (e0 + row) * K + 4 * q) =
            *reinterpret_cast<const float4*>(st.h + row * Stage<K>::PITCH + 4 * q);
      }
      constexpr int SCH = ST * (K - 1) / 4;
      for (int c = tid; c < SCH; c += ST)
        *reinterpret_cast<float4*>(p.g_s + e0 * (K - 1) + 4 * c) = *reinterpret_cast<const float4*>(st.s + 4 * c);
    } else {
      const int nv = (int)left;
      for (int c = tid; c < nv * K; c += ST) {
        const int row = c / K, k = c - row * K;
        p.g_w[(e0 + row) * K + k] = st.w[row * Stage<K>::PITCH + k];
        p.g_h[(e0 + row) * K + k] = st.h[row * Stage<K>::PITCH + k];
      }
      for (int c = tid; c < nv * (K - 1); c += ST) p.g_s[e0 * (K - 1) + c] = st.s[c];
    }
    __syncthreads();
  }
}

inline bool al16(const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

template <int K>
vms_status launch(const StreamParams& p, bool inv, bool bwd, cudaStream_t st) {
  const size_t smem = bwd ? sizeof(StageBwd<K>) : sizeof(Stage<K>);
  const int64_t n_tiles = (p.n + ST - 1) / ST;
  const int64_t cap = (bwd ? 4LL : 5LL) * sm_count();
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
#define VMS_RS_LAUNCH(KERN)                                                                              \
  do {                                                                                                   \
    static bool configured = false;                                                                      \
    if (!configured) {                                                                                   \
      VMS_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
      configured = true;                                                                                 \
    }                                                                                                    \
    KERN<<<grid, ST, smem, st>>>(p);                                                                     \
  } while (0)
  if (!bwd) {
    if (inv) VMS_RS_LAUNCH((rqs_stream_apply_kernel<K, true>));
    else VMS_RS_LAUNCH((rqs_stream_apply_kernel<K, false>));
    VMS_LAUNCH_CHECK("rqs_stream_apply_kernel");
  } else {
    if (inv) VMS_RS_LAUNCH((rqs_stream_backward_kernel<K, true>));
    else VMS_RS_LAUNCH((rqs_stream_backward_kernel<K, false>));
    VMS_LAUNCH_CHECK("rqs_stream_backward_kernel");
  }
#undef VMS_RS_LAUNCH
  return VMS_OK;
}

}  // namespace

// Returns true (and the launch status in *status) when the streaming kernels take the call.
bool rqs_stream_try(const float* v, const float* rw, const float* rh, const float* rs, int64_t n, int K, float bin_min,
                    float bin_max, int inverse_dir, float* out, float* ldj, const float* g_out, const float* g_ldj,
                    float* g_in, float* g_w, float* g_h, float* g_s, bool backward, cudaStream_t st, vms_status* status) {
  if (K != 32 && K != 20) return false;
  if (n < 4 * ST) return false;  // tiny calls: the octet kernel has the shorter critical path
  if (!(al16(v) && al16(rw) && al16(rh) && al16(rs) && al16(g_out) && al16(g_ldj) && al16(g_w) && al16(g_h) && al16(g_s)))
    return false;
  StreamParams p = {};
  p.x = v; p.raw_w = rw; p.raw_h = rh; p.raw_s = rs; p.n = n;
  p.bin_min = bin_min;
  p.scale = (float)((double)bin_max - (double)bin_min - (double)K * 1e-2);  // flows.py:92
  p.y = out; p.ldj = ldj;
  p.g_out = g_out; p.g_ldj = g_ldj; p.g_in = g_in; p.g_w = g_w; p.g_h = g_h; p.g_s = g_s;
  *status = K == 32 ? launch<32>(p, inverse_dir != 0, backward, st) : launch<20>(p, inverse_dir != 0, backward, st);
  return true;
}

}  // namespace vms
