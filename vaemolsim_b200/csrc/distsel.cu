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
// Design (B200): one CTA per reference row streams that row's N coordinates exactly ONCE in the common case.
// The streaming pass keeps only candidates with d^2 <= cutoff^2 (all that can ever be non-zero in the output) as
// 64-bit keys (d^2 bits << 32 | index) in shared memory; the key order IS the top_k order, so a shared-memory
// bitonic sort of the handful of candidates finishes the row.  Only when the caller asks for the exact top_k
// indices of beyond-cutoff fill slots, or more than kCap particles fall inside the cutoff, does the row fall back
// to an exact 4-pass radix select over d^2 (re-reading the row from L2).
#include "common.cuh"
#include <math.h>

namespace vms {

constexpr int DT = 256;     // threads per CTA
constexpr int kCap = 2048;  // candidate capacity (keys) per row

struct DistSelParams {
  const float* coords; const int64_t* row_splits; int64_t B, N;
  const float* ref; const float* box; int box_per_row;
  float sq_cut; int k;
  const float* info; int P;
  float* out_xyz; float* out_info; int32_t* out_idx;
};

struct Local { float x, y, z, d2; };

// bit-exact restatement of TF's op-by-op float32 arithmetic (no FMA contraction, IEEE division, rint = half-even)
struct Box {
  bool has;
  float bx, by, bz, ix, iy, iz;  // lengths and their float32 reciprocals
};

// rint(x / L) exactly as IEEE division would give it, without dividing in the common case: q = x * (1/L) is within
// 2^-23 |q| of the true quotient, so rint(q) can differ from rint(x / L) only when q sits that close to a half-integer;
// only then is the division carried out (the XU pipe was 37 % busy with three MUFU.RCP per particle).
__device__ __forceinline__ float image_count(float x, float L, float invL) {
  const float q = __fmul_rn(x, invL);
  float k = rintf(q);
  if (fabsf(fabsf(q - k) - 0.5f) <= 1e-6f * fabsf(q) + 1e-7f || !(fabsf(q) < 4194304.f)) k = rintf(__fdiv_rn(x, L));
  return k;
}

__device__ __forceinline__ Local local_of_v(float cx, float cy, float cz, float rx, float ry, float rz, const Box& bo) {
  Local l;
  l.x = __fsub_rn(cx, rx);
  l.y = __fsub_rn(cy, ry);
  l.z = __fsub_rn(cz, rz);
  if (bo.has) {
    l.x = __fsub_rn(l.x, __fmul_rn(bo.bx, image_count(l.x, bo.bx, bo.ix)));
    l.y = __fsub_rn(l.y, __fmul_rn(bo.by, image_count(l.y, bo.by, bo.iy)));
    l.z = __fsub_rn(l.z, __fmul_rn(bo.bz, image_count(l.z, bo.bz, bo.iz)));
  }
  l.d2 = __fadd_rn(__fadd_rn(__fmul_rn(l.x, l.x), __fmul_rn(l.y, l.y)), __fmul_rn(l.z, l.z));
  return l;
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

__global__ void __launch_bounds__(DT) dist_select_kernel(const DistSelParams p) {
  __shared__ unsigned long long keys[kCap];
  __shared__ unsigned hist[256];
  __shared__ unsigned warp_cnt[DT / 32];
  __shared__ unsigned s_count, s_prefix, s_krem, s_running;

  const int64_t b = blockIdx.x;
  const int64_t start = p.row_splits ? p.row_splits[b] : b * p.N;
  const int64_t n64 = p.row_splits ? p.row_splits[b + 1] - start : p.N;
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
  bo.has = has_box; bo.bx = bx; bo.by = by; bo.bz = bz;
  bo.ix = __frcp_rn(bx); bo.iy = __frcp_rn(by); bo.iz = __frcp_rn(bz);
  auto consider = [&](float cx, float cy, float cz, int i) {
    const Local l = local_of_v(cx, cy, cz, rx, ry, rz, bo);
    if (l.d2 <= p.sq_cut) {
      const unsigned pos = atomicAdd(&s_count, 1u);
      if (pos < (unsigned)kCap) keys[pos] = make_key(l.d2, (unsigned)i);
    }
  };
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(crow) & 15u) == 0) {
    // a thread takes 4 consecutive particles = three 16-byte loads (48 contiguous bytes), two groups in flight
    const float4* c4 = reinterpret_cast<const float4*>(crow);
    const int n4 = n / 4;
#pragma unroll 2
    for (int g = threadIdx.x; g < n4; g += DT) {
      const float4 f0 = __ldg(c4 + 3 * (size_t)g), f1 = __ldg(c4 + 3 * (size_t)g + 1), f2 = __ldg(c4 + 3 * (size_t)g + 2);
      consider(f0.x, f0.y, f0.z, 4 * g);
      consider(f0.w, f1.x, f1.y, 4 * g + 1);
      consider(f1.z, f1.w, f2.x, 4 * g + 2);
      consider(f2.y, f2.z, f2.w, 4 * g + 3);
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
  const int n_in = (int)s_count;
  int n_list;  // number of valid keys in `keys`
  const bool fast = (n_in <= kCap) && (n_in >= k || p.out_idx == nullptr);
  if (fast) {
    n_list = n_in;
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
  if (n_pow2 > 1) bitonic_sort(keys, n_pow2);
  else __syncthreads();

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

}  // namespace vms

using namespace vms;

extern "C" vms_status vms_dist_select(const float* coords, const int64_t* row_splits, int64_t B, int64_t N,
                                      const float* ref, const float* box, int box_per_row, float cutoff_sq, int k,
                                      const float* info, int P, float* out_xyz, float* out_info, int32_t* out_idx,
                                      vms_stream stream) {
  VMS_REQUIRE(B >= 0 && N >= 0, VMS_ERR_SHAPE, "dist_select: bad shape");
  VMS_REQUIRE(k >= 1 && k <= kCap, VMS_ERR_INVALID_ARG, "dist_select: max_included must be in [1, %d], got %d", kCap, k);
  VMS_REQUIRE(N < (1LL << 31) && B < (1LL << 31), VMS_ERR_SHAPE, "dist_select: too many particles / rows");
  VMS_REQUIRE(B == 0 || (ref && out_xyz && (coords || (N == 0 && !row_splits))), VMS_ERR_INVALID_ARG,
              "dist_select: NULL pointer");
  VMS_REQUIRE((out_info == nullptr) || (info != nullptr && P >= 1), VMS_ERR_INVALID_ARG,
              "dist_select: particle_info required for out_info");
  if (B == 0) return VMS_OK;
  DistSelParams p = {coords, row_splits, B, N, ref, box, box_per_row, cutoff_sq, k, info, P, out_xyz, out_info, out_idx};
  dist_select_kernel<<<(unsigned)B, DT, 0, as_stream(stream)>>>(p);
  VMS_LAUNCH_CHECK("dist_select_kernel");
  return VMS_OK;
}
