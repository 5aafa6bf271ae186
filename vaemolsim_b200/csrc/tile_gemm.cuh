// tile_gemm.cuh -- CTA-level building blocks shared by the fused ELBO kernel (elbo_fused.cu) and the fused MC kernel
// (mc_fused.cu): a 32-row tile of configurations lives in dynamic shared memory, 512 threads work on it.
//   outer_gemm  FP32 FFMA "outer-product" GEMM for wide layers (forward, input-gradient and weight-gradient forms)
//   thin_gemm   thin layers (<= 16 outputs): contraction split over the 16 warps, partial sums combined in shared memory
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

// Thin outputs: out[r][n] (+)= sum_k X[r * ldx + k] * W[k * sWk + n * sWn] (+ bias[n]),  n < N <= kMaxThin, FR rows; X, W,
// out, bias, scratch are shared-memory offsets.  The contraction is split 16 ways: warp w takes k in
// [w Kd/16, (w+1) Kd/16), its lanes take the 32 rows (ldx odd => conflict-free), the weight row is a broadcast load
// (128-bit when n is the contiguous index); the 16 partial sums meet in `scratch` (16 x round4(N) x 32 floats) and are
// added in warp order.  (The first version reduced over lanes with shuffles: 60 % of the MC kernel's samples.)
// Contains __syncthreads(): call from uniform control flow; the caller synchronises before using `out`.
static __device__ __noinline__ void thin_gemm(int X, int ldx, int W, int sWk, int sWn, int Kd, int N, int out, int ldo,
                                              int bias, int accumulate, int scratch) {
  extern __shared__ __align__(16) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kc = (Kd + FW - 1) / FW;
  const int k0 = min(Kd, warp * kc), k1 = min(Kd, k0 + kc);
  const int Np = (N + 3) & ~3;
  const int xo = X + lane * ldx;
#pragma unroll 1
  for (int n0 = 0; n0 < N; n0 += 4) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const bool vec = sWn == 1 && ((W + n0) & 3) == 0 && (sWk & 3) == 0 && n0 + 4 <= N;
    if (vec) {
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const float x = sm[xo + k];
        const float4 w = *reinterpret_cast<const float4*>(sm + W + k * sWk + n0);
        a0 = fmaf(x, w.x, a0); a1 = fmaf(x, w.y, a1); a2 = fmaf(x, w.z, a2); a3 = fmaf(x, w.w, a3);
      }
    } else {
      const int c1 = min(n0 + 1, N - 1) * sWn, c2 = min(n0 + 2, N - 1) * sWn, c3 = min(n0 + 3, N - 1) * sWn;
#pragma unroll 4
      for (int k = k0; k < k1; ++k) {
        const float x = sm[xo + k];
        const int wo = W + k * sWk;
        a0 = fmaf(x, sm[wo + n0 * sWn], a0); a1 = fmaf(x, sm[wo + c1], a1);
        a2 = fmaf(x, sm[wo + c2], a2); a3 = fmaf(x, sm[wo + c3], a3);
      }
    }
    float* sc = sm + scratch + (warp * Np + n0) * FR + lane;
    sc[0] = a0; sc[FR] = a1; sc[2 * FR] = a2; sc[3 * FR] = a3;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N * FR; e += FT) {
    const int n = e / FR, r = e - n * FR;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < FW; ++w) v += sm[scratch + (w * Np + n) * FR + r];
    if (bias >= 0) v += sm[bias + n];
    const int o = out + r * ldo + n;
    sm[o] = accumulate ? sm[o] + v : v;
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
