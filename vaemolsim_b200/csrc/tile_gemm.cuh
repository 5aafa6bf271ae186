// tile_gemm.cuh -- CTA-level building blocks shared by the fused ELBO kernel (elbo_fused.cu) and the fused MC kernel
// (mc_fused.cu): a 32-row tile of configurations lives in dynamic shared memory, 512 threads work on it.
//   outer_gemm  FP32 FFMA "outer-product" GEMM for wide layers (forward, input-gradient and weight-gradient forms)
//   rowdot      thin layers (<= 16 outputs) as per-row dot products with a warp-shuffle reduction
//   Epi         POD epilogue descriptor (bias / activation / mask / partial-gradient accumulate)
//   cp_async4   4-byte global -> shared copies for weight staging
#pragma once
#include "common.cuh"

namespace vms {

constexpr int FR = 32;            // rows per tile
constexpr int FT = 512;           // threads per CTA (16 warps: the phases are latency-bound chains, TLP is what hides them)
constexpr int FW = FT / 32;       // warps per CTA
constexpr int kMaxThin = 16;      // widest "thin" layer (2 dx, 2 dz, conditioner inputs)

// ------------------------------------------------------------------------------------------------ async staging
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ GEMM routines
// The kernel is a long chain of small phases, each executed once per tile: instruction FETCH, not issue, bounded the
// first version (ncu: stall_no_instruction 1.8 per issue with every GEMM call site inlined and its epilogue fully
// unrolled).  So the routines below are __noinline__, take shared-memory operands as OFFSETS into the dynamic
// shared array (the compiler keeps LDS/STS addressing across the call) and a small POD epilogue descriptor instead
// of a lambda: six GEMM instantiations serve the 12 call sites.
struct Epi {
  int kind;    // 0: out = act(v + bias[j]);  1: out = aux > 0 ? v : 0 (relu');  2: out = v (1 - aux^2) (tanh');
               // 3: global partial gradient g[i * si + j * sj] (= or +=)
  int out, ld;            // kinds 0-2: shared offset / pitch of out[i][j]
  int bias, relu;         // kind 0: shared offset of bias[j] (-1: none), relu flag
  int aux, ld_aux;        // kinds 1-2
  float* g; int si, sj, first;  // kind 3
};

__device__ __forceinline__ void epi_apply(const Epi& e, float* sm, int i, int j, float v) {
  if (e.kind == 0) {
    if (e.bias >= 0) v += sm[e.bias + j];
    sm[e.out + i * e.ld + j] = e.relu ? fmaxf(v, 0.f) : v;
  } else if (e.kind == 1) {
    sm[e.out + i * e.ld + j] = sm[e.aux + i * e.ld_aux + j] > 0.f ? v : 0.f;
  } else if (e.kind == 2) {
    const float h = sm[e.aux + i * e.ld_aux + j];
    sm[e.out + i * e.ld + j] = v * (1.f - h * h);
  } else {
    float* d = e.g + i * e.si + j * e.sj;
    *d = e.first ? v : *d + v;
  }
}

// out[i][j] = sum_t S[t * sSt + i] * L[t * sLt + j * sLj]          (S, L: offsets into shared memory)
//   S : i contiguous, rows 16-byte aligned; a warp reads its slab of TI rows as 128-bit broadcasts
//   L : lanes own j = j0 + 32 c, c < TJ (conflict-free when sLj = 1)
// Work items (i-slab, j-group) are dealt to the 16 warps.  The inner loop carries no predicates: out-of-range lanes
// re-read column J - 1 and slabs may read a few rows past I (operands are padded); their results are discarded.
// SPLIT: the contraction is halved between warps w and w + 8, the upper half hands its partial sums over through
// `part` (pitch ldp, normally the destination itself) -- this doubles the slab height a warp can afford, which is what
// moves the loop from shared-memory-bandwidth-bound (TI = 4: 3 FFMA per wavefront) to FFMA-bound (TI = 8: 4.8).
// Contains __syncthreads() when SPLIT: call from uniform control flow.
template <int TI, int TJ, bool SPLIT>
__device__ __noinline__ void outer_gemm(int S, int sSt, int L, int sLt, int sLj, int I, int J, int T, int part, int ldp,
                                        const Epi e) {
  static_assert(TI % 4 == 0, "slab height must be a multiple of 4");
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_is = (I + TI - 1) / TI, n_jg = (J + 32 * TJ - 1) / (32 * TJ);
  const int n_items = n_is * n_jg;
  const int stride = SPLIT ? FW / 2 : FW;
  const int ts = SPLIT ? warp / (FW / 2) : 0;
  const int Th = SPLIT ? (T + 1) / 2 : T;
  const int t0 = ts * Th, t1 = min(T, t0 + Th);
  const int n_rounds = (n_items + stride - 1) / stride;
#pragma unroll 1
  for (int round = 0; round < n_rounds; ++round) {
    const int item = round * stride + (SPLIT ? (warp & (FW / 2 - 1)) : warp);
    const bool valid = item < n_items;
    const int is = valid ? item % n_is : 0, jg = valid ? item / n_is : 0;
    const int i0 = is * TI, j0 = jg * 32 * TJ + lane;
    float acc[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int c = 0; c < TJ; ++c) acc[i][c] = 0.f;
    if (valid) {
      int lo[TJ];
#pragma unroll
      for (int c = 0; c < TJ; ++c) lo[c] = L + min(j0 + 32 * c, J - 1) * sLj + t0 * sLt;
      int so = S + i0 + t0 * sSt;
#pragma unroll 2
      for (int t = t0; t < t1; ++t) {
        float l[TJ];
#pragma unroll
        for (int c = 0; c < TJ; ++c) {
          l[c] = sm[lo[c]];
          lo[c] += sLt;
        }
#pragma unroll
        for (int q = 0; q < TI / 4; ++q) {
          const float4 s4 = *reinterpret_cast<const float4*>(sm + so + 4 * q);
#pragma unroll
          for (int c = 0; c < TJ; ++c) {
            acc[4 * q + 0][c] = fmaf(s4.x, l[c], acc[4 * q + 0][c]);
            acc[4 * q + 1][c] = fmaf(s4.y, l[c], acc[4 * q + 1][c]);
            acc[4 * q + 2][c] = fmaf(s4.z, l[c], acc[4 * q + 2][c]);
            acc[4 * q + 3][c] = fmaf(s4.w, l[c], acc[4 * q + 3][c]);
          }
        }
        so += sSt;
      }
    }
    if (SPLIT) {
      if (valid && ts == 1) {
#pragma unroll
        for (int c = 0; c < TJ; ++c) {
          const int j = j0 + 32 * c;
          if (j < J) {
#pragma unroll
            for (int i = 0; i < TI; ++i)
              if (i0 + i < I) sm[part + (i0 + i) * ldp + j] = acc[i][c];
          }
        }
      }
      __syncthreads();
    }
    if (valid && ts == 0) {
#pragma unroll
      for (int c = 0; c < TJ; ++c) {
        const int j = j0 + 32 * c;
        if (j < J) {
#pragma unroll
          for (int i = 0; i < TI; ++i)
            if (i0 + i < I) epi_apply(e, sm, i0 + i, j, SPLIT ? acc[i][c] + sm[part + (i0 + i) * ldp + j] : acc[i][c]);
        }
      }
    }
  }
}

// Thin outputs: out[r][n] (+)= sum_k X[r * ldx + k] * W[k * sWk + n * sWn] (+ bias[n]),  n < N <= kMaxThin; X, W, out,
// bias are shared-memory offsets.  Warp w owns rows 2 w, 2 w + 1, lanes stride over k, totals by warp shuffle.
static __device__ __noinline__ void rowdot(int X, int ldx, int W, int sWk, int sWn, int Kd, int N, int out, int ldo, int bias,
                                    int accumulate) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int RW = FR / FW;
  constexpr int NB = 4;  // outputs per pass: small code (this routine runs once per call site and tile)
#pragma unroll 1
  for (int n0 = 0; n0 < N; n0 += NB) {
    float acc[RW][NB];
#pragma unroll
    for (int rr = 0; rr < RW; ++rr)
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[rr][n] = 0.f;
#pragma unroll 1
    for (int k = lane; k < Kd; k += 32) {
      float xv[RW];
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) xv[rr] = sm[X + (warp * RW + rr) * ldx + k];
      const int wo = W + k * sWk + n0 * sWn;
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const float w = sm[wo + min(n, N - 1 - n0) * sWn];
#pragma unroll
        for (int rr = 0; rr < RW; ++rr) acc[rr][n] = fmaf(xv[rr], w, acc[rr][n]);
      }
    }
#pragma unroll
    for (int n = 0; n < NB; ++n) {
#pragma unroll
      for (int rr = 0; rr < RW; ++rr) {
        float v = warp_sum(acc[rr][n]);
        if (lane == 0 && n0 + n < N) {
          const int o = out + (warp * RW + rr) * ldo + n0 + n;
          if (bias >= 0) v += sm[bias + n0 + n];
          sm[o] = accumulate ? sm[o] + v : v;
        }
      }
    }
  }
}

__device__ __forceinline__ Epi epi_store(int out, int ld, int bias, int relu) {
  Epi e = {};
  e.kind = 0; e.out = out; e.ld = ld; e.bias = bias; e.relu = relu;
  return e;
}
__device__ __forceinline__ Epi epi_mask(int kind, int out, int ld, int aux, int ld_aux) {
  Epi e = {};
  e.kind = kind; e.out = out; e.ld = ld; e.aux = aux; e.ld_aux = ld_aux;
  return e;
}
__device__ __forceinline__ Epi epi_grad(float* g, int si, int sj, bool first) {
  Epi e = {};
  e.kind = 3; e.g = g; e.si = si; e.sj = sj; e.first = first ? 1 : 0;
  return e;
}

}  // namespace vms
