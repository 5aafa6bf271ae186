// distsel.cu -- K6: minimum-image distance-based local-environment selection (exact top-k, bit-exact indices).
//
// Replaces `DistanceSelection.call` mappings.py:362-455:
//   :404      local = coords - ref
//   :408-412  local -= box * tf.round(local / box)         (round half to even; separate div / mul / sub ops)
//   :417-426  ragged -> dense with float32.max padding (d^2 = +inf), padded up to max_included
//   :429      d^2 = reduce_sum(local * local)               ((x^2 + y^2) + z^2, products rounded separately)
//   :433      top_k(-d^2, k)                                (ascending d^2, ties -> lower index first)
//   :436-441  gather, zero where d^2 > cutoff^2;  :443-453 same gather + mask on particle_info
//
// Design (B200): one CTA per reference row streams that row's N coordinates exactly ONCE.  The streaming pass keeps
// candidates -- d^2 <= cutoff^2 (all that can ever be non-zero in the output) and, when the caller wants tf.math.top_k's
// exact INDICES for beyond-cutoff fill slots too, d^2 <= a per-row sample threshold that bounds the k-th smallest d^2
// (sample_threshold) -- as 64-bit keys (d^2 bits << 32 | index) in shared memory; the key order IS the top_k order, so a
// histogram refinement + shared-memory bitonic sort of ~1.3 k keys finishes the row.  Only when more than kCap candidates
// survive does the row fall back to an exact 4-pass radix select over d^2 (re-reading the row from L2).
// Round 2: at the C3 shape (N = 10^4, k = 50, P = 2, indices on) only ~11 particles lie inside the cutoff, so round 1's
// kernel took the five-pass fallback on EVERY row (0.77 ms, 0.10 of the HBM peak); with the sample threshold, the lean
// quad routine and the refinement the same call takes 0.167 ms (0.46), the values-only call 0.124 ms (0.61).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace vms {

constexpr int DT = 256;     // threads per CTA
constexpr int kCap = 2048;  // candidate capacity (keys) per row

struct DistSelParams {
  const float* coords; const int64_t* row_splits; int64_t B, N;
  const float* ref; const float* box; int box_per_row;
  float sq_cut; int k;
  const float* info; int P;
  float* out_xyz; float* out_info; int32_t* out_idx;
  int dbg;  // development aid (VMS_DS_DBG): 1 = the stream kernel consumes the ring without arithmetic
  int shared_frame;  // every row selects from the SAME N particles (coords [N, 3], info [N, P]): vms_dist_select_frame
};

struct Local { float x, y, z, d2; };

// bit-exact restatement of TF's op-by-op float32 arithmetic (no FMA contraction, IEEE division, rint = half-even)
struct Box {
  bool has;
  float bx, by, bz, ix, iy, iz;  // lengths and their float32 reciprocals
  float hx, hy, hz;              // 0.4999 x length: below it the candidate image count is certainly exact
  __device__ __forceinline__ void set(float x, float y, float z) {
    bx = x; by = y; bz = z;
    ix = __frcp_rn(x); iy = __frcp_rn(y); iz = __frcp_rn(z);
    hx = 0.4999f * fabsf(x); hy = 0.4999f * fabsf(y); hz = 0.4999f * fabsf(z);
  }
};

// x - L * rint(x / L) exactly as TF's separate IEEE div / round / mul / sub ops give it, without dividing in the common
// case.  k = rint(x * (1/L)) is a CANDIDATE image count; l = x - L k (rounded product, rounded difference: the very value
// TF computes when k is right).  If |l| <= 0.4999 L and |x / L| < 256, the true quotient Q satisfies |Q - k| < 0.49995
// (the computed l is within 2^-22 |x| of x - k L), hence fl(Q) is within 2^-24 |Q| of that and rint(fl(x / L)) = k: the
// candidate is the exact answer.  Only otherwise (a particle within 0.01 % of the half-box plane) is the division carried
// out.  (The first version tested the fractional part of the quotient: 12 instructions per axis.)
__device__ __forceinline__ float wrap_axis(float x, float L, float invL, float halfL) {
  const float q = __fmul_rn(x, invL);
  float k = rintf(q);
  float l = __fsub_rn(x, __fmul_rn(L, k));
  if (!(fabsf(l) <= halfL) || !(fabsf(q) < 256.f)) {
    k = rintf(__fdiv_rn(x, L));
    l = __fsub_rn(x, __fmul_rn(L, k));
  }
  return l;
}

__device__ __forceinline__ Local local_of_v(float cx, float cy, float cz, float rx, float ry, float rz, const Box& bo) {
  Local l;
  l.x = __fsub_rn(cx, rx);
  l.y = __fsub_rn(cy, ry);
  l.z = __fsub_rn(cz, rz);
  if (bo.has) {
    l.x = wrap_axis(l.x, bo.bx, bo.ix, bo.hx);
    l.y = wrap_axis(l.y, bo.by, bo.iy, bo.hy);
    l.z = wrap_axis(l.z, bo.bz, bo.iz, bo.hz);
  }
  l.d2 = __fadd_rn(__fadd_rn(__fmul_rn(l.x, l.x), __fmul_rn(l.y, l.y)), __fmul_rn(l.z, l.z));
  return l;
}

// ---- the hot loop's unit of work: d^2 of FOUR consecutive particles (48 contiguous bytes = three 16-byte words).
// The streaming kernels issue tens of instructions per particle (ncu, round 2: 57 at 58 % issue utilisation with DRAM at
// 51 %), so this routine is written for instruction count: rint() is the magic-number addition (two full-rate FADDs instead
// of a quarter-rate FRND; exact half-even for |q| < 2^22, and |q| < 256 is checked anyway), the "candidate image count
// is certainly exact" test of wrap_axis is evaluated branch-free -- per axis the largest |l| and |q| of the four
// particles (FMNMX), six comparisons per quad -- and the exact IEEE-division path is an out-of-line call taken by a quad
// only when that predicate fires (a particle within 0.01 % of a half-box plane).
__device__ __noinline__ float4 quad_d2_exact(float4 f0, float4 f1, float4 f2, float rx, float ry, float rz, float bx, float by,
                                             float bz) {
  const float cx[4] = {f0.x, f0.w, f1.z, f2.y}, cy[4] = {f0.y, f1.x, f1.w, f2.z}, cz[4] = {f0.z, f1.y, f2.x, f2.w};
  float d2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float lx = __fsub_rn(cx[u], rx), ly = __fsub_rn(cy[u], ry), lz = __fsub_rn(cz[u], rz);
    lx = __fsub_rn(lx, __fmul_rn(bx, rintf(__fdiv_rn(lx, bx))));
    ly = __fsub_rn(ly, __fmul_rn(by, rintf(__fdiv_rn(ly, by))));
    lz = __fsub_rn(lz, __fmul_rn(bz, rintf(__fdiv_rn(lz, bz))));
    d2[u] = __fadd_rn(__fadd_rn(__fmul_rn(lx, lx), __fmul_rn(ly, ly)), __fmul_rn(lz, lz));
  }
  return make_float4(d2[0], d2[1], d2[2], d2[3]);
}

__device__ __forceinline__ float wrap_fast(float x, float L, float invL, float& q_out) {
  const float kMagic = 12582912.f;  // 1.5 * 2^23: (q + kMagic) - kMagic = rint(q), ties to even, for |q| < 2^22
  const float q = __fmul_rn(x, invL);
  const float k = __fsub_rn(__fadd_rn(q, kMagic), kMagic);
  q_out = q;
  return __fsub_rn(x, __fmul_rn(L, k));
}

__device__ __forceinline__ void quad_d2(const float4 f0, const float4 f1, const float4 f2, float rx, float ry, float rz,
                                        const Box& bo, float (&d2)[4]) {
  float lx[4] = {__fsub_rn(f0.x, rx), __fsub_rn(f0.w, rx), __fsub_rn(f1.z, rx), __fsub_rn(f2.y, rx)};
  float ly[4] = {__fsub_rn(f0.y, ry), __fsub_rn(f1.x, ry), __fsub_rn(f1.w, ry), __fsub_rn(f2.z, ry)};
  float lz[4] = {__fsub_rn(f0.z, rz), __fsub_rn(f1.y, rz), __fsub_rn(f2.x, rz), __fsub_rn(f2.w, rz)};
  if (bo.has) {
    float wx[4], wy[4], wz[4];
    float mx = 0.f, my = 0.f, mz = 0.f, mq = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float qx, qy, qz;
      wx[u] = wrap_fast(lx[u], bo.bx, bo.ix, qx);
      wy[u] = wrap_fast(ly[u], bo.by, bo.iy, qy);
      wz[u] = wrap_fast(lz[u], bo.bz, bo.iz, qz);
      mx = fmaxf(mx, fabsf(wx[u])); my = fmaxf(my, fabsf(wy[u])); mz = fmaxf(mz, fabsf(wz[u]));
      mq = fmaxf(mq, fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz))));
    }
    // (fmaxf drops NaNs: a NaN coordinate gives a NaN d^2 on both paths, which no comparison below accepts)
    const bool bad = (mx > bo.hx) | (my > bo.hy) | (mz > bo.hz) | !(mq < 256.f);
    if (bad) {
      const float4 e = quad_d2_exact(f0, f1, f2, rx, ry, rz, bo.bx, bo.by, bo.bz);
      d2[0] = e.x; d2[1] = e.y; d2[2] = e.z; d2[3] = e.w;
      return;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { lx[u] = wx[u]; ly[u] = wy[u]; lz[u] = wz[u]; }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    d2[u] = __fadd_rn(__fadd_rn(__fmul_rn(lx[u], lx[u]), __fmul_rn(ly[u], ly[u])), __fmul_rn(lz[u], lz[u]));
}

// PREFILTER arithmetic for the hot loop: d^2 of four consecutive particles with contracted operations (FFMA for the
// image count's rounding, the wrap and the sum of squares: 17 instead of 32 instructions per particle).  Not TF's
// arithmetic -- it only decides which particles get the exact evaluation afterwards.  |d2_approx - d2_exact| is bounded in
// the kernel (see `slack`); an image count that differs from the exact one can only happen within 0.01 % of a half-box
// plane, where both values are ~ (L / 2)^2, far above any threshold the prefilter is used with.
__device__ __forceinline__ void quad_d2_approx(const float4 f0, const float4 f1, const float4 f2, float rx, float ry, float rz,
                                               const Box& bo, float (&d2)[4]) {
  const float kMagic = 12582912.f;
  bool far[4];
  float lx[4] = {f0.x - rx, f0.w - rx, f1.z - rx, f2.y - rx};
  float ly[4] = {f0.y - ry, f1.x - ry, f1.w - ry, f2.z - ry};
  float lz[4] = {f0.z - rz, f1.y - rz, f2.x - rz, f2.w - rz};
  if (bo.has) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float kx = fmaf(lx[u], bo.ix, kMagic) - kMagic, ky = fmaf(ly[u], bo.iy, kMagic) - kMagic,
                  kz = fmaf(lz[u], bo.iz, kMagic) - kMagic;
      lx[u] = fmaf(-bo.bx, kx, lx[u]);
      ly[u] = fmaf(-bo.by, ky, ly[u]);
      lz[u] = fmaf(-bo.bz, kz, lz[u]);
      // the error bound assumes |image count| <= 1 (a particle at most 1.5 box lengths from the site): anything further
      // away is handed to the exact evaluation unconditionally (d2 = -1 passes every threshold)
      far[u] = fmaxf(fabsf(kx), fmaxf(fabsf(ky), fabsf(kz))) > 1.5f;
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) far[u] = false;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float v = fmaf(lz[u], lz[u], fmaf(ly[u], ly[u], lx[u] * lx[u]));
    d2[u] = far[u] ? -1.f : v;
  }
}

__device__ __forceinline__ Local local_of(const float* __restrict__ c, float rx, float ry, float rz, bool has_box,
                                          float bx, float by, float bz) {
  Local l;
  l.x = __fsub_rn(c[0], rx);
  l.y = __fsub_rn(c[1], ry);
  l.z = __fsub_rn(c[2], rz);
  if (has_box) {
    l.x = __fsub_rn(l.x, __fmul_rn(bx, rintf(__fdiv_rn(l.x, bx))));
    l.y = __fsub_rn(l.y, __fmul_rn(by, rintf(__fdiv_rn(l.y, by))));
    l.z = __fsub_rn(l.z, __fmul_rn(bz, rintf(__fdiv_rn(l.z, bz))));
  }
  l.d2 = __fadd_rn(__fadd_rn(__fmul_rn(l.x, l.x), __fmul_rn(l.y, l.y)), __fmul_rn(l.z, l.z));
  return l;
}

__device__ __forceinline__ unsigned long long make_key(float d2, unsigned idx) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | idx;
}

__device__ void bitonic_sort(unsigned long long* keys, int n_pow2) {
  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < n_pow2 / 2; t += DT) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

// The same network for up to 64 keys in ONE warp (two keys per lane, exchanges by shuffle): the usual C3 row ends with ~60
// candidates, and 21 CTA barriers for 32 compare-exchanges each were ~8 % of the kernel's stall samples.  Call from warp 0
// between two CTA barriers; keys[n_pow2 .. 64) are not read.
__device__ __forceinline__ void warp_sort64(unsigned long long* keys, int n_pow2) {
  const int lane = threadIdx.x;  // < 32
  unsigned long long r0 = lane < n_pow2 ? keys[lane] : ~0ull;
  unsigned long long r1 = lane + 32 < n_pow2 ? keys[lane + 32] : ~0ull;
#pragma unroll
  for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride == 32) {  // size 64, ascending: element `lane` against element `lane + 32`
        const unsigned long long lo = r0 < r1 ? r0 : r1, hi = r0 < r1 ? r1 : r0;
        r0 = lo; r1 = hi;
      } else {
        const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, r0, stride), p1 = __shfl_xor_sync(0xffffffffu, r1, stride);
        const bool lower = (lane & stride) == 0;
        const bool up0 = (lane & size) == 0, up1 = ((lane + 32) & size) == 0;
        r0 = (lower == up0) ? (r0 < p0 ? r0 : p0) : (r0 < p0 ? p0 : r0);
        r1 = (lower == up1) ? (r1 < p1 ? r1 : p1) : (r1 < p1 ? p1 : r1);
      }
    }
  }
  if (lane < n_pow2) keys[lane] = r0;
  if (lane + 32 < n_pow2) keys[lane + 32] = r1;
}

// Shared state of a CTA (both kernels)
struct RowSmem {
  unsigned long long keys[kCap];
  unsigned hist[1024];
  unsigned warp_cnt[DT / 32];
  unsigned s_count, s_prefix, s_krem, s_running, s_thr;
};

// Sample threshold (exact top_k indices in one pass): sm.hist holds a 1024-bin histogram of the d^2 bit patterns (>> 21:
// 8 exponent + 2 mantissa bits) of a sample of the row; warp 0 finds the first bin where the running count reaches k and
// publishes the largest bit pattern of that bin in sm.s_thr (+inf when the sample holds fewer than k particles).  At
// least k sample particles lie at or below it, so it bounds the row's k-th smallest d^2 from above.  The threshold only
// prunes: order and values come from the exact keys, so results stay bit-identical.  Call from warp 0 between two CTA
// barriers.
__device__ __forceinline__ void sample_threshold(RowSmem& sm, int k) {
  const int tid = threadIdx.x;  // < 32; lane l scans bins [32 l, 32 l + 32)
  unsigned run = 0;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) run += sm.hist[tid * 32 + i];
  unsigned incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (tid >= o) incl += t;
  }
  const unsigned excl = incl - run;
  if (excl < (unsigned)k && incl >= (unsigned)k) {  // exactly one lane: the k-th smallest falls in its bins
    unsigned acc = excl;
    int bin = tid * 32;
    for (int i = 0; i < 32; ++i) {
      acc += sm.hist[tid * 32 + i];
      if (acc >= (unsigned)k) { bin = tid * 32 + i; break; }
    }
    sm.s_thr = ((unsigned)(bin + 1) << 21) - 1u;
  }
  const unsigned total_s = __shfl_sync(0xffffffffu, incl, 31);
  if (tid == 0 && total_s < (unsigned)k) sm.s_thr = 0x7f800000u;
}

// Everything after the streaming pass: `sm.keys[0 .. n_in)` hold the candidates the pass kept (n_in may exceed kCap: then
// the buffer is incomplete).  `have_k`: the candidate set is known to contain the k nearest particles (the streaming pass
// used a threshold >= the k-th smallest d^2), so the exact top_k indices come out of the sort even for beyond-cutoff slots.
// Falls back to an exact 4-pass radix select over the row (re-read from L2) when the buffer overflowed or when the exact
// indices of beyond-cutoff fill slots are wanted and the candidates do not cover them.
__device__ void finish_row(const DistSelParams& p, RowSmem& sm, int64_t b, int64_t start, int n, float rx, float ry, float rz,
                           bool has_box, float bx, float by, float bz, int n_in, bool have_k) {
  unsigned long long* keys = sm.keys;
  unsigned* hist = sm.hist;
  unsigned* warp_cnt = sm.warp_cnt;
  unsigned& s_count = sm.s_count;
  unsigned& s_prefix = sm.s_prefix;
  unsigned& s_krem = sm.s_krem;
  unsigned& s_running = sm.s_running;
  const float* crow = p.coords + start * 3;
  const int k = p.k;
  int n_list;  // number of valid keys in `keys`
  const bool fast = (n_in <= kCap) && (n_in >= k || p.out_idx == nullptr || have_k);
  if (fast) {
    n_list = n_in;
    if (n_in > 2 * k && n_in > 128) {
      // many more candidates than outputs (the sample threshold keeps ~k N / 1024 of them): instead of sorting them all,
      // histogram their d^2 on the top 10 bits, find the bin holding the k-th smallest, keep only keys up to that bin
      // (~1.3 k of them) and sort those.  Keys are read into registers before the compaction overwrites the buffer.
      __syncthreads();
      for (int i = threadIdx.x; i < 1024; i += DT) hist[i] = 0;
      __syncthreads();
      unsigned long long mine[kCap / DT];
#pragma unroll
      for (int u = 0; u < kCap / DT; ++u) {
        const int i = threadIdx.x + u * DT;
        mine[u] = i < n_in ? keys[i] : ~0ull;
        if (i < n_in) atomicAdd(&hist[(unsigned)(mine[u] >> 53)], 1u);
      }
      if (threadIdx.x == 0) s_count = 0;
      __syncthreads();
      if (threadIdx.x < 32) sample_threshold(sm, k);
      __syncthreads();
      const unsigned thr = sm.s_thr;
#pragma unroll
      for (int u = 0; u < kCap / DT; ++u) {
        if ((unsigned)(mine[u] >> 32) <= thr) {
          const unsigned pos = atomicAdd(&s_count, 1u);
          keys[pos] = mine[u];
        }
      }
      __syncthreads();
      n_list = (int)s_count;  // >= k
    }
  } else {
    // ---- exact selection of the k smallest (d^2, index) keys
    __syncthreads();
    if (n <= k) {
      for (int i = threadIdx.x; i < n; i += DT) {
        const Local l = local_of(crow + (size_t)i * 3, rx, ry, rz, has_box, bx, by, bz);
        keys[i] = make_key(l.d2, (unsigned)i);
      }
      n_list = n;
    } else {
      // 4-pass MSB radix select on the d^2 bit pattern (non-negative floats order like unsigned ints)
      if (threadIdx.x == 0) { s_prefix = 0; s_krem = (unsigned)k; }
      for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[threadIdx.x] = 0;  // DT == 256
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int i = threadIdx.x; i < n; i += DT) {
          const unsigned bits = __float_as_uint(local_of(crow + (size_t)i * 3, rx, ry, rz, has_box, bx, by, bz).d2);
          const bool match = pass == 0 || (bits >> (shift + 8)) == prefix;
          if (match) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
          unsigned rem = s_krem, bkt = 0;
          for (; bkt < 255; ++bkt) {
            if (hist[bkt] >= rem) break;
            rem -= hist[bkt];
          }
          s_prefix = (prefix << 8) | bkt;
          s_krem = rem;  // how many of the k fall in the chosen bucket (>= 1)
        }
        __syncthreads();
      }
      const unsigned T = s_prefix;  // bit pattern of the k-th smallest d^2
      const unsigned m_ties = s_krem;  // number of elements equal to T to take, lowest indices first
      if (threadIdx.x == 0) { s_count = 0; s_running = 0; }
      __syncthreads();
      for (int c0 = 0; c0 < n; c0 += DT) {
        const int i = c0 + threadIdx.x;
        unsigned bits = 0xffffffffu;
        if (i < n) bits = __float_as_uint(local_of(crow + (size_t)i * 3, rx, ry, rz, has_box, bx, by, bz).d2);
        const bool less = i < n && bits < T;
        const bool tie = i < n && bits == T;
        const unsigned bal = __ballot_sync(0xffffffffu, tie);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        unsigned rank = s_running + __popc(bal & ((1u << lane) - 1u));
        unsigned total = 0;
        for (int w = 0; w < DT / 32; ++w) {
          if (w < warp) rank += warp_cnt[w];
          total += warp_cnt[w];
        }
        if (less || (tie && rank < m_ties)) {
          const unsigned pos = atomicAdd(&s_count, 1u);
          keys[pos] = ((unsigned long long)bits << 32) | (unsigned)i;  // pos < k <= kCap by construction
        }
        __syncthreads();
        if (threadIdx.x == 0) s_running += total;
      }
      __syncthreads();
      n_list = (int)s_count;  // == k
    }
    // float32.max padding rows (mappings.py:417-426): d^2 = +inf, indices n, n+1, ...
    for (int j = n_list + threadIdx.x; j < k; j += DT) keys[j] = make_key(INFINITY, (unsigned)(n + (j - n_list)));
    if (n_list < k) n_list = k;
  }
  __syncthreads();

  // ---- sort candidates; key order == tf.math.top_k order (ascending d^2, ties by lower index)
  int n_pow2 = 1;
  while (n_pow2 < n_list) n_pow2 <<= 1;
  for (int j = n_list + threadIdx.x; j < n_pow2; j += DT) keys[j] = ~0ull;
  if (n_pow2 > 64) {
    bitonic_sort(keys, n_pow2);
  } else {
    __syncthreads();
    if (threadIdx.x < 32 && n_pow2 > 1) warp_sort64(keys, n_pow2);
    __syncthreads();
  }

  // ---- emit the first k
  float* oxyz = p.out_xyz + b * (int64_t)k * 3;
  for (int j = threadIdx.x; j < k; j += DT) {
    float ox = 0.f, oy = 0.f, oz = 0.f;
    bool keep = false;
    unsigned idx = 0;
    if (j < n_list) {
      const unsigned long long key = keys[j];
      idx = (unsigned)(key & 0xffffffffu);
      const float d2 = __uint_as_float((unsigned)(key >> 32));
      keep = (idx < (unsigned)n) && (d2 <= p.sq_cut);
      if (keep) {
        const Local l = local_of(crow + (size_t)idx * 3, rx, ry, rz, has_box, bx, by, bz);
        ox = l.x; oy = l.y; oz = l.z;
      }
    }
    oxyz[j * 3] = ox; oxyz[j * 3 + 1] = oy; oxyz[j * 3 + 2] = oz;
    if (p.out_idx) p.out_idx[b * (int64_t)k + j] = (int32_t)idx;
    if (p.out_info) {
      float* oi = p.out_info + (b * (int64_t)k + j) * p.P;
      const float* ii = p.info + (start + idx) * p.P;
      for (int c = 0; c < p.P; ++c) oi[c] = keep ? ii[c] : 0.f;
    }
  }
}


template <bool PREF>
__global__ void __launch_bounds__(DT, 6) dist_select_kernel(const DistSelParams p) {
  __shared__ RowSmem sm;
  unsigned long long* keys = sm.keys;
  unsigned& s_count = sm.s_count;

  const int64_t b = blockIdx.x;
  const int64_t start = p.shared_frame ? 0 : (p.row_splits ? p.row_splits[b] : b * p.N);
  const int64_t n64 = (p.row_splits && !p.shared_frame) ? p.row_splits[b + 1] - start : p.N;
  const int n = (int)n64;
  const float* crow = p.coords + start * 3;
  const float rx = p.ref[b * 3], ry = p.ref[b * 3 + 1], rz = p.ref[b * 3 + 2];
  const bool has_box = p.box != nullptr;
  float bx = 1.f, by = 1.f, bz = 1.f;
  if (has_box) {
    const float* bp = p.box + (p.box_per_row ? b * 3 : 0);
    bx = bp[0]; by = bp[1]; bz = bp[2];
  }
  const int k = p.k;

  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();

  // ---- pass A: stream the row once, keep within-cutoff candidates.  Loads first, arithmetic after: the first version
  // (one particle per iteration, a shared atomic in the loop body) spent 58 % of its samples waiting on the row's loads.
  Box bo;
  bo.has = has_box;
  bo.set(bx, by, bz);
  // exact top_k indices wanted and the row is long: the first kSample particles are a sample whose d^2 histogram gives a
  // threshold >= the k-th smallest d^2 of the row (see sample_threshold), so ONE streaming pass collects every particle
  // that can appear in the top k (round 1 re-read such rows five times)
  constexpr int kSample = 1024;
  const bool want_k = p.out_idx != nullptr && n >= kSample && k <= 512;
  unsigned thr_bits = 0u;
  // the tiled call on an aligned row: the sample IS the first round of the stream (kSample = 4 DT: one quad per thread, the
  // same exact arithmetic) -- its d^2 stay in registers, are pushed once the threshold is known, and the stream starts at
  // the second round (the first version evaluated the sample particle by particle and then streamed it again)
  static_assert(kSample == 4 * DT, "the sample is one quad per thread");
  const bool reuse_sample = want_k && !PREF && (reinterpret_cast<uintptr_t>(crow) & 15u) == 0;
  if (want_k) {
    for (int i = threadIdx.x; i < 1024; i += DT) sm.hist[i] = 0;
    __syncthreads();
    float sd2[4] = {0.f, 0.f, 0.f, 0.f};
    if (reuse_sample) {
      const float4* c4 = reinterpret_cast<const float4*>(crow);
      const size_t g = threadIdx.x;
      const float4 f0 = __ldg(c4 + 3 * g), f1 = __ldg(c4 + 3 * g + 1), f2 = __ldg(c4 + 3 * g + 2);
      quad_d2(f0, f1, f2, rx, ry, rz, bo, sd2);
#pragma unroll
      for (int u = 0; u < 4; ++u) atomicAdd(&sm.hist[__float_as_uint(sd2[u]) >> 21], 1u);
    } else {
      for (int i = threadIdx.x; i < kSample; i += DT) {
        const Local l = local_of_v(__ldg(crow + (size_t)i * 3), __ldg(crow + (size_t)i * 3 + 1), __ldg(crow + (size_t)i * 3 + 2),
                                   rx, ry, rz, bo);
        atomicAdd(&sm.hist[__float_as_uint(l.d2) >> 21], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) sample_threshold(sm, k);
    __syncthreads();
    thr_bits = sm.s_thr;
    if (reuse_sample) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (sd2[u] <= p.sq_cut || __float_as_uint(sd2[u]) <= thr_bits) {
          const unsigned pos = atomicAdd(&s_count, 1u);
          if (pos < (unsigned)kCap) keys[pos] = make_key(sd2[u], (unsigned)(4 * threadIdx.x + u));
        }
      }
    }
  }
  auto consider = [&](float cx, float cy, float cz, int i) {
    const Local l = local_of_v(cx, cy, cz, rx, ry, rz, bo);
    if (l.d2 <= p.sq_cut || __float_as_uint(l.d2) <= thr_bits) {
      const unsigned pos = atomicAdd(&s_count, 1u);
      if (pos < (unsigned)kCap) keys[pos] = make_key(l.d2, (unsigned)i);
    }
  };
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(crow) & 15u) == 0) {
    // a thread takes 4 consecutive particles = three 16-byte loads (48 contiguous bytes), two groups in flight
    const float4* c4 = reinterpret_cast<const float4*>(crow);
    const int n4 = n / 4;
    // With an L2-resident frame the kernel is issue-bound (ncu: issue-active 74 %), and TF's unfused float32 op order costs
    // 32 instructions per particle.  So the stream runs a contracted PREFILTER (17 instructions) against a
    // threshold widened by a proven error bound, records the indices of the few particles that pass (within the cutoff or
    // under the sample threshold: ~5 % of a C3 row), and only those are evaluated in TF's arithmetic afterwards: the
    // accepted set, every key and every output bit are exactly those of the exact loop.
    const float T = fmaxf(p.sq_cut, want_k ? __uint_as_float(thr_bits) : 0.f);
    const float Lmax = has_box ? fmaxf(fabsf(bx), fmaxf(fabsf(by), fabsf(bz))) : 0.f;
    const float Lmin = has_box ? fminf(fabsf(bx), fminf(fabsf(by), fabsf(bz))) : INFINITY;
    // with equal image counts |k| <= 1 (enforced in quad_d2_approx), per axis |w_approx - w_exact| <= 2^-24 L + 2^-23 |w|
    // (the exact path rounds L k, then the difference); delta = 2^-19 Lmax is 30 x that.  The sums of squares differ by
    // <= 2^-21 d2.  slack = 4 x the resulting bound on |d2_approx - d2_exact|.
    const float delta = has_box ? 1.9073486e-6f * Lmax : 2.4e-7f * sqrtf(T);
    const float slack = 4.f * (6.f * sqrtf(T) * delta + 3.f * delta * delta + 4.8e-7f * T);
    const float T_pre = T + slack;
    // Used for the shared-frame entry only: with one frame per row in HBM the kernel is bound by the stream and its
    // phase structure, not by issue slots (measured at C3: 0.189 ms with the prefilter, 0.179 ms without; with an
    // L2-resident frame 0.105 / 0.065 ms against 0.122 / 0.085 ms).
    const bool prefilter = PREF && T_pre < 0.24f * Lmin * Lmin && T_pre < 1e30f && p.dbg != 2;  // (0.49 L)^2: image counts agree
    if (prefilter) {
#pragma unroll 2
      for (int g = threadIdx.x; g < n4; g += DT) {
        const float4 f0 = __ldg(c4 + 3 * (size_t)g), f1 = __ldg(c4 + 3 * (size_t)g + 1), f2 = __ldg(c4 + 3 * (size_t)g + 2);
        float d2[4];
        quad_d2_approx(f0, f1, f2, rx, ry, rz, bo, d2);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (d2[u] <= T_pre) {
            const unsigned pos = atomicAdd(&s_count, 1u);
            if (pos < (unsigned)kCap) keys[pos] = (unsigned long long)(unsigned)(4 * g + u);
          }
        }
      }
      __syncthreads();
      // exact evaluation of the recorded particles (read into registers first: the list is compacted in place)
      const unsigned n_c = s_count;
      if (n_c <= (unsigned)kCap) {
        unsigned long long mine[kCap / DT];
#pragma unroll
        for (int u = 0; u < kCap / DT; ++u) {
          const unsigned i = threadIdx.x + u * DT;
          mine[u] = ~0ull;
          if (i < n_c) {
            const unsigned idx = (unsigned)keys[i];
            const Local l = local_of_v(__ldg(crow + (size_t)idx * 3), __ldg(crow + (size_t)idx * 3 + 1),
                                       __ldg(crow + (size_t)idx * 3 + 2), rx, ry, rz, bo);
            if (l.d2 <= p.sq_cut || __float_as_uint(l.d2) <= thr_bits) mine[u] = make_key(l.d2, idx);
          }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kCap / DT; ++u) {
          if (mine[u] != ~0ull) {
            const unsigned pos = atomicAdd(&s_count, 1u);
            keys[pos] = mine[u];
          }
        }
      }
      // (more than kCap recorded: s_count > kCap sends finish_row to its exact radix selection over the row)
    } else {
#pragma unroll 2
      for (int g = threadIdx.x + (reuse_sample ? DT : 0); g < n4; g += DT) {
        const float4 f0 = __ldg(c4 + 3 * (size_t)g), f1 = __ldg(c4 + 3 * (size_t)g + 1), f2 = __ldg(c4 + 3 * (size_t)g + 2);
        float d2[4];
        quad_d2(f0, f1, f2, rx, ry, rz, bo, d2);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (d2[u] <= p.sq_cut || __float_as_uint(d2[u]) <= thr_bits) {
            const unsigned pos = atomicAdd(&s_count, 1u);
            if (pos < (unsigned)kCap) keys[pos] = make_key(d2[u], (unsigned)(4 * g + u));
          }
        }
      }
    }
    done = n4 * 4;
  }
  {
    constexpr int U = 4;
    for (int base = done; base < n; base += DT * U) {
      float cx[U], cy[U], cz[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * DT + threadIdx.x;
        const bool ok = i < n;
        cx[u] = ok ? __ldg(crow + (size_t)i * 3) : 0.f;
        cy[u] = ok ? __ldg(crow + (size_t)i * 3 + 1) : 0.f;
        cz[u] = ok ? __ldg(crow + (size_t)i * 3 + 2) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * DT + threadIdx.x;
        if (i < n) consider(cx[u], cy[u], cz[u], i);
      }
    }
  }
  __syncthreads();
  finish_row(p, sm, b, start, n, rx, ry, rz, has_box, bx, by, bz, (int)sm.s_count, want_k);
}


// ------------------------------------------------------------------------------------------------ streaming kernel
// Alternative path for dense inputs whose row stride is a multiple of 16 bytes (VMS_DISTSEL_STREAM=1; see DESIGN.md for
// the measurements): PERSISTENT CTAs, each walking rows blockIdx.x, blockIdx.x + gridDim.x, ...; a row arrives as chunks
// of kChunk particles through a kStages-deep ring of cp.async.bulk copies (the 1-D TMA path; one `full` mbarrier per stage
// signalled by the copy, one `empty` mbarrier per stage on which every warp arrives after reading it), issued by one thread
// kStages chunks ahead ACROSS row boundaries, so the copy engine keeps streaming while a row is being sorted and emitted.
// There is no CTA barrier inside a row: warps drift apart by up to kStages - 1 chunks.  Threads read particles from shared
// memory as 16-byte words (4 particles = 48 bytes per thread and chunk; a quarter warp covers all 32 banks).
// The first chunk of a row doubles as the sample of the exact-top_k threshold (sample_threshold).
// Measured (4096 rows x 10,000 particles): the ring alone (no arithmetic) streams at 0.77 of the HBM copy peak; with the
// arithmetic the kernel reaches 0.55 -- below the one-CTA-per-row kernel's 0.61, which therefore stays the default.
constexpr int kChunk = 1024;  // particles per ring stage (12 KB)
constexpr int kStages = 4;

__device__ __forceinline__ unsigned ds_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"  // suspend-time hint: the warp sleeps in hardware
        "selp.u32 %0, 1, 0, P1;\n"                                       // instead of spinning through issue slots
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(2000u)
        : "memory");
  }
}

__global__ void __launch_bounds__(DT, 3) dist_select_stream_kernel(const DistSelParams p) {
  extern __shared__ __align__(128) unsigned char dsm[];
  float* ring = reinterpret_cast<float*>(dsm);                                   // [kStages][kChunk * 3]
  RowSmem& sm = *reinterpret_cast<RowSmem*>(dsm + (size_t)kStages * kChunk * 12);
  __shared__ __align__(8) unsigned long long full[kStages], empty[kStages];
  const int tid = threadIdx.x;
  const int n = (int)p.N;
  const int cpr = (n + kChunk - 1) / kChunk;  // chunks per row
  const int my_rows = (int)((p.B - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int total = my_rows * cpr;  // (all loop counters are 32-bit and incremental: 64-bit division is a subroutine)
  const int k = p.k;
  const bool has_box = p.box != nullptr;
  const bool want_k = p.out_idx != nullptr && n > k;  // exact indices of beyond-cutoff slots: sample threshold
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(ds_smem_u32(&full[s])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(ds_smem_u32(&empty[s])), "r"(DT / 32) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const unsigned full0 = ds_smem_u32(&full[0]), empty0 = ds_smem_u32(&empty[0]);
  int i_row = 0, i_ch = 0, i_st = 0;  // producer cursor (thread 0): next chunk to issue
  auto issue = [&]() {  // thread 0: the next chunk of this CTA's sequence -> the next stage
    const int64_t row = (int64_t)blockIdx.x + (int64_t)i_row * gridDim.x;
    const int ch = i_ch, stg = i_st;
    const int cnt = min(kChunk, n - ch * kChunk);
    const unsigned bytes = (unsigned)cnt * 12u;
    const unsigned bar = full0 + 8u * (unsigned)stg;
    if (++i_ch == cpr) { i_ch = 0; ++i_row; }
    if (++i_st == kStages) i_st = 0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the stage's generic-proxy reads are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     ds_smem_u32(ring + (size_t)stg * kChunk * 3)),
                 "l"(p.coords + (row * (int64_t)n + (int64_t)ch * kChunk) * 3), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (tid == 0)
    for (int q = 0; q < kStages && q < total; ++q) issue();

  float rx = 0.f, ry = 0.f, rz = 0.f;
  Box bo;
  bo.has = has_box;
  bo.set(1.f, 1.f, 1.f);
  unsigned thr_bits = 0u;  // keep d^2 whose bit pattern is <= thr_bits (in addition to d^2 <= cutoff^2)
  int r_loc = 0, ch = 0, st = 0;
  unsigned parity = 0;  // phase parity of the stage barriers: flips every kStages chunks
#pragma unroll 1
  for (int q = 0; q < total; ++q) {
    const int64_t b = (int64_t)blockIdx.x + (int64_t)r_loc * gridDim.x;
    const int cnt = min(kChunk, n - ch * kChunk);
    if (ch == 0) {  // new row
      rx = __ldg(p.ref + b * 3); ry = __ldg(p.ref + b * 3 + 1); rz = __ldg(p.ref + b * 3 + 2);
      if (has_box) {
        const float* bp = p.box + (p.box_per_row ? b * 3 : 0);
        bo.set(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2));
      }
      if (tid == 0) sm.s_count = 0;
      if (want_k)
        for (int i = tid; i < 1024; i += DT) sm.hist[i] = 0;
      thr_bits = 0u;
      __syncthreads();
    }
    mbar_wait_parity(full0 + 8u * (unsigned)st, parity);
    // this thread's 4 particles of the chunk: 48 contiguous bytes of the stage
    const float4* c4 = reinterpret_cast<const float4*>(ring + (size_t)st * kChunk * 3);
    const int g = tid;  // kChunk / 4 == DT groups
    float d2[4];
    const int base = ch * kChunk + 4 * g;
    const int nv = min(4, cnt - 4 * g);  // valid particles of the group (<= 0: none)
    if (nv > 0) {
      if (p.dbg == 1) { d2[0] = d2[1] = d2[2] = d2[3] = 1e30f + c4[3 * g].x; } else
      quad_d2(c4[3 * g], c4[3 * g + 1], c4[3 * g + 2], rx, ry, rz, bo, d2);
    }
    if (ch == 0 && want_k) {
      // sample threshold: histogram of the chunk's d^2 on the top 10 bits, first bin where the running count reaches k
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nv) atomicAdd(&sm.hist[__float_as_uint(d2[u]) >> 21], 1u);
      __syncthreads();
      if (tid < 32) sample_threshold(sm, k);
      __syncthreads();
      thr_bits = sm.s_thr;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < nv && (d2[u] <= p.sq_cut || __float_as_uint(d2[u]) <= thr_bits)) {
        const unsigned pos = atomicAdd(&sm.s_count, 1u);
        if (pos < (unsigned)kCap) sm.keys[pos] = make_key(d2[u], (unsigned)(base + u));
      }
    }
    // release the stage: one arrival per warp on its `empty` barrier; thread 0 refills the stage once all warps have
    // released it
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(empty0 + 8u * (unsigned)st) : "memory");
    if (tid == 0 && q + kStages < total) {
      mbar_wait_parity(empty0 + 8u * (unsigned)st, parity);
      issue();
    }
    if (ch == cpr - 1) {
      __syncthreads();  // every candidate of the row is in the key buffer
      float bx = bo.bx, by = bo.by, bz = bo.bz;
      finish_row(p, sm, b, b * (int64_t)n, n, rx, ry, rz, has_box, bx, by, bz, (int)sm.s_count, want_k);
      __syncthreads();
      ch = 0;
      ++r_loc;
    } else {
      ++ch;
    }
    if (++st == kStages) { st = 0; parity ^= 1u; }
  }
}

}  // namespace vms

using namespace vms;

static vms_status dist_select_impl(const float* coords, const int64_t* row_splits, int64_t B, int64_t N, const float* ref,
                                   const float* box, int box_per_row, float cutoff_sq, int k, const float* info, int P,
                                   float* out_xyz, float* out_info, int32_t* out_idx, int shared_frame, vms_stream stream) {
  VMS_RANGE("vms_dist_select");
  VMS_REQUIRE(B >= 0 && N >= 0, VMS_ERR_SHAPE, "dist_select: bad shape");
  VMS_REQUIRE(k >= 1 && k <= kCap, VMS_ERR_INVALID_ARG, "dist_select: max_included must be in [1, %d], got %d", kCap, k);
  VMS_REQUIRE(N < (1LL << 31) && B < (1LL << 31), VMS_ERR_SHAPE, "dist_select: too many particles / rows");
  VMS_REQUIRE(B == 0 || (ref && out_xyz && (coords || (N == 0 && !row_splits))), VMS_ERR_INVALID_ARG,
              "dist_select: NULL pointer");
  VMS_REQUIRE((out_info == nullptr) || (info != nullptr && P >= 1), VMS_ERR_INVALID_ARG,
              "dist_select: particle_info required for out_info");
  if (B == 0) return VMS_OK;
  DistSelParams p = {coords, row_splits, B, N, ref, box, box_per_row, cutoff_sq, k, info, P, out_xyz, out_info, out_idx, 0,
                     shared_frame};
  if (const char* e = getenv("VMS_DS_DBG")) p.dbg = atoi(e);
  // VMS_DISTSEL_STREAM=1: dense rows whose stride is a multiple of 16 bytes stream through the TMA ring (persistent CTAs);
  // default: one CTA per row with direct 16-byte loads, which measured faster
  static int stream_on = -1;
  if (stream_on < 0) {
    const char* e = getenv("VMS_DISTSEL_STREAM");
    stream_on = (e && e[0] == '1') ? 1 : 0;
  }
  const bool aligned = !shared_frame && !row_splits && N % 4 == 0 && N >= kChunk && k <= 512 && (reinterpret_cast<uintptr_t>(coords) & 15u) == 0;
  if (stream_on && aligned) {
    const size_t smem = (size_t)kStages * kChunk * 12 + sizeof(RowSmem) + 128;
    static bool attr_set = false;
    if (!attr_set) {
      VMS_CUDA(cudaFuncSetAttribute(dist_select_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    int per_sm = (int)((size_t)max_smem_optin() / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 3) per_sm = 3;
    const int64_t grid = B < (int64_t)sm_count() * per_sm ? B : (int64_t)sm_count() * per_sm;
    dist_select_stream_kernel<<<(unsigned)grid, DT, smem, as_stream(stream)>>>(p);
    VMS_LAUNCH_CHECK("dist_select_stream_kernel");
    return VMS_OK;
  }
  static bool carve_set = false;
  if (!carve_set) {  // 20 KB of static shared memory per CTA: ask for the largest carve-out so that occupancy is register-bound
    cudaFuncSetAttribute(dist_select_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(dist_select_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaGetLastError();
    carve_set = true;
  }
  // (the prefilter variant is a separate instantiation: its code must not cost the tiled call registers)
  if (shared_frame) dist_select_kernel<true><<<(unsigned)B, DT, 0, as_stream(stream)>>>(p);
  else dist_select_kernel<false><<<(unsigned)B, DT, 0, as_stream(stream)>>>(p);
  VMS_LAUNCH_CHECK("dist_select_kernel");
  return VMS_OK;
}

extern "C" vms_status vms_dist_select(const float* coords, const int64_t* row_splits, int64_t B, int64_t N,
                                      const float* ref, const float* box, int box_per_row, float cutoff_sq, int k,
                                      const float* info, int P, float* out_xyz, float* out_info, int32_t* out_idx,
                                      vms_stream stream) {
  return dist_select_impl(coords, row_splits, B, N, ref, box, box_per_row, cutoff_sq, k, info, P, out_xyz, out_info, out_idx, 0,
                          stream);
}

extern "C" vms_status vms_dist_select_frame(const float* coords, int64_t N, const float* ref, int64_t B, const float* box,
                                            int box_per_row, float cutoff_sq, int k, const float* info, int P, float* out_xyz,
                                            float* out_info, int32_t* out_idx, vms_stream stream) {
  return dist_select_impl(coords, nullptr, B, N, ref, box, box_per_row, cutoff_sq, k, info, P, out_xyz, out_info, out_idx, 1,
                          stream);
}
