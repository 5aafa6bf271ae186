// gemm_tc.cu -- K2 at scale: Dense forward  out[B, N] = act(x[B, K] W[K, N] + b)  on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulator in TMEM) with float32-grade accuracy through a 3 x TF32 split.
//
// Replaces the same Keras Dense layers as dense.cu (mappings.py:107-121, flows.py:136-152 -- above all the concatenated
// spline heads of flows.py:140-152, [B, 100] x [100, 95], 88 % of the model's FLOPs) when the batch makes the layer a
// real contraction (north_star: "tensor-core (tcgen05) GEMMs only ... where batch x width makes them real
// contractions"): B >= 8192 rows.  Below that the FFMA paths (dense.cu, the fused ELBO kernel) have the shorter
// critical path.
//
// float32 parity.  kind::tf32 reads float32 containers and uses 10 mantissa bits: 1e-3 relative, outside the 1e-5
// budget.  So every operand is split a = a_hi + a_lo (a_hi = cvt.rna.tf32(a), a_lo = cvt.rna.tf32(a - a_hi), both exactly
// representable) and the product is accumulated as a_lo b_hi + a_hi b_lo + a_hi b_hi in the float32 TMEM
// accumulators (leading products and corrections separately, added in the epilogue): the dropped a_lo b_lo term is
// 2^-22 relative.  Three MMAs per k-step.
//
// Structure (one CTA per SM, 256 threads, persistent over 128-row tiles):
//   * W is split once per CTA into shared memory (B operand, N x K "K-major", no swizzle: core matrices of 8 rows x 16
//     bytes, 128 bytes apart along N, (Np/8) * 128 bytes apart along K);
//   * per tile, two threads per row load x with 16-byte loads (all loads first, registers), split it and writes hi / lo in the same canonical
//     layout (A operand, 128 x K); fence.proxy.async hands the tile to the tensor core;
//   * ONE thread issues the 3 * K/8 tcgen05.mma (M = 128, N = Np, K = 8) and commits them to an mbarrier;
//   * the eight warps read their TMEM lane quarter (= 32 rows, alternate 16-column chunks) with tcgen05.ld 32x32b, add bias / apply the activation, stage
//     the tile in shared memory and write it out as ONE contiguous, fully coalesced block (the tile is contiguous in a
//     row-major [B, N] output).
// This first version is not pipelined (load -> MMA -> epilogue per tile are serial inside a CTA); the 148 CTAs overlap
// each other's phases.  TMA loads and a second A stage are the next step.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>

namespace vms {

namespace {

constexpr int TM = 128;     // rows per tile = UMMA M
constexpr int TT = 256;     // threads per CTA: two threads per row for the loads, two warps per TMEM lane quarter
constexpr int kMaxKc = 16;  // 16-byte chunks of one x row held in registers PER THREAD (two threads share a row: K <= 128)

struct TcParams {
  const float* x; int64_t ld_x;
  const float* W; const float* bias;
  float* out; int64_t ld_out;
  int64_t B;
  int K, N, Kp, Np, act;
  unsigned tmem_cols;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// round-to-nearest (ties away) to TF32's 10 mantissa bits -- what cvt.rna.tf32.f32 does for finite values -- with two
// integer instructions: ncu showed cvt.rna.tf32 throttling on its (slow) conversion pipe, 14 % of the kernel's samples
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE ("interleave"), version 1 (Blackwell)
//   bits [0,14) start address >> 4, [16,30) leading-dimension byte offset >> 4 (between the two 16-byte K chunks of one
//   MMA), [32,46) stride byte offset >> 4 (between 8-row core matrices), [46,48) version, [61,64) layout type = 0
__device__ __forceinline__ unsigned long long make_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void mma_tf32(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_wait_parity(unsigned bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TC_DONE;\n"
      "bra TC_WAIT;\n"
      "TC_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(phase)
      : "memory");
}

// element (row, k) of an operand with `rows` rows in the canonical K-major no-swizzle layout (float index)
__device__ __forceinline__ int canon(int row, int k, int rows) { return ((k >> 2) * rows + row) * 4 + (k & 3); }

__global__ void __launch_bounds__(TT, 1) dense_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_base_s;
  __shared__ float s_bias[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Kp = p.Kp, Np = p.Np, K = p.K, N = p.N;
  float* a_hi = sm;                      // [Kp/4][128][4]
  float* a_lo = a_hi + TM * Kp;
  float* b_hi = a_lo + TM * Kp;          // [Kp/4][Np][4]
  float* b_lo = b_hi + Np * Kp;
  float* s_out = a_hi;                   // epilogue staging [128][N], aliases the A tile (free once the MMAs completed)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int n = tid; n < Np; n += TT) s_bias[n] = (p.bias && n < N) ? __ldg(p.bias + n) : 0.f;
  // B operand: W[k][n] -> (n, k) K-major, split, zero padding up to (Np, Kp)
  for (int e = tid; e < Np * Kp; e += TT) {
    const int k = e / Np, n = e - k * Np;
    const float w = (k < K && n < N) ? __ldg(p.W + (size_t)k * N + n) : 0.f;
    const float hi = tf32_rna(w);
    const int o = canon(n, k, Np);
    b_hi[o] = hi;
    b_lo[o] = tf32_rna(w - hi);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem_d = tmem_base_s, tmem_d2 = tmem_base_s + p.tmem_cols / 2;
  // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3, M >> 4
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(Np >> 3) << 17) | ((unsigned)(TM >> 4) << 24);
  const unsigned a_lbo = (TM / 8) * 128, b_lbo = (unsigned)(Np / 8) * 128, sbo = 128;
  unsigned phase = 0;

  const int64_t n_tiles = (p.B + TM - 1) / TM;
  // A-tile loads: two threads per row, each holds its half of the row in registers (all loads issued before any use).
  // The loads of tile i + 1 are issued right after tile i's MMAs, so HBM latency hides behind the tensor core and the
  // epilogue (software pipelining through registers: shared memory has no room for a second A stage).
  const int lr = tid & (TM - 1), half = tid >> 7;
  const int nkc_all = Kp / 4, per = (nkc_all + 1) / 2;
  const int kc0 = half * per, nkc = min(per, nkc_all - kc0);  // this thread's chunks: [kc0, kc0 + nkc)
  float4 v[kMaxKc];
  auto load_rows = [&](int64_t t) {
    const int64_t r0 = t * TM;
    const bool ok = t < n_tiles && r0 + lr < p.B;
    const float* xr = p.x + (r0 + lr) * p.ld_x;
#pragma unroll
    for (int i = 0; i < kMaxKc; ++i) {
      const int kc = kc0 + i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < nkc && ok) {
        if (4 * kc + 3 < K) {
          v[i] = __ldg(reinterpret_cast<const float4*>(xr + 4 * kc));
        } else {
          if (4 * kc < K) v[i].x = __ldg(xr + 4 * kc);
          if (4 * kc + 1 < K) v[i].y = __ldg(xr + 4 * kc + 1);
          if (4 * kc + 2 < K) v[i].z = __ldg(xr + 4 * kc + 2);
        }
      }
    }
  };
  load_rows(blockIdx.x);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    const int nr = (int)min((int64_t)TM, p.B - row0);
    // ---- A tile: split the prefetched row halves, store hi / lo (conflict-free: consecutive rows are consecutive
    // 16-byte slots of a k-chunk)
#pragma unroll
    for (int i = 0; i < kMaxKc; ++i) {
      if (i < nkc) {
        const int kc = kc0 + i;
        float4 h, l;
        h.x = tf32_rna(v[i].x); h.y = tf32_rna(v[i].y); h.z = tf32_rna(v[i].z); h.w = tf32_rna(v[i].w);
        l.x = tf32_rna(v[i].x - h.x); l.y = tf32_rna(v[i].y - h.y);
        l.z = tf32_rna(v[i].z - h.z); l.w = tf32_rna(v[i].w - h.w);
        *reinterpret_cast<float4*>(a_hi + (kc * TM + lr) * 4) = h;
        *reinterpret_cast<float4*>(a_lo + (kc * TM + lr) * 4) = l;
      }
    }
    // generic-proxy writes -> visible to the tensor core (async proxy), then one thread issues the MMAs
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0) {
      const unsigned ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
      for (int k8 = 0; k8 < Kp / 8; ++k8) {
        const unsigned ao = (unsigned)k8 * 2u * a_lbo, bo = (unsigned)k8 * 2u * b_lbo;
        // the leading products a_hi b_hi go to accumulator 1, the two correction products (2^-11 smaller) to accumulator
        // 2: the tensor core's float32 accumulation is not round-to-nearest, and its error grows with the number of
        // accumulation steps at FULL magnitude -- 13 instead of 39 this way (measured: 5.1e-6 of the dot product's scale
        // with one accumulator, 2.3e-6 with two; the FFMA kernel has 2.1e-6).  The epilogue adds the two.
        mma_tf32(tmem_d2, make_desc(al + ao, a_lbo, sbo), make_desc(bh + bo, b_lbo, sbo), idesc, k8 > 0 ? 1u : 0u);
        mma_tf32(tmem_d2, make_desc(ah + ao, a_lbo, sbo), make_desc(bl + bo, b_lbo, sbo), idesc, 1u);
        mma_tf32(tmem_d, make_desc(ah + ao, a_lbo, sbo), make_desc(bh + bo, b_lbo, sbo), idesc, k8 > 0 ? 1u : 0u);
      }
      // completion of everything issued so far -> mbarrier (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar))
                   : "memory");
    }
    load_rows(tile + gridDim.x);  // next tile's rows: in flight while the tensor core and the epilogue work
    mbar_wait_parity(smem_u32(&mbar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // ---- epilogue: warp w owns TMEM lanes 32 w .. 32 w + 31 = tile rows; 32 columns per tcgen05.ld
    {
      // warps w and w + 4 share TMEM lane quarter w & 3 (rows 32 (w & 3) ..) and take alternate 16-column chunks
      const int q = warp & 3, r = 32 * q + lane;
      for (int c0 = 16 * (warp >> 2); c0 < Np; c0 += 32) {
        float v[16], v2[16];
        tmem_ld16(tmem_d + ((unsigned)(32 * q) << 16) + (unsigned)c0, v);
        tmem_ld16(tmem_d2 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v2);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = c0 + i;
          if (n < N) {
            float y = (v[i] + v2[i]) + s_bias[n];
            if (p.act == VMS_ACT_RELU) y = fmaxf(y, 0.f);
            else if (p.act == VMS_ACT_TANH) y = tanhf(y);
            s_out[r * N + n] = y;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (p.ld_out == N) {
      // the tile is one contiguous block of nr * N floats in a row-major [B, N] output
      float* dst = p.out + row0 * N;
      const int total = nr * N;
      if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int t4 = total / 4;
        for (int i = tid; i < t4; i += TT)
          reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_out)[i];
        for (int i = 4 * t4 + tid; i < total; i += TT) dst[i] = s_out[i];
      } else {
        for (int i = tid; i < total; i += TT) dst[i] = s_out[i];
      }
    } else {
      for (int i = tid; i < nr * N; i += TT) {
        const int r = i / N, n = i - r * N;
        p.out[(row0 + r) * p.ld_out + n] = s_out[i];
      }
    }
    __syncthreads();  // s_out aliases the next tile's A operand
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(p.tmem_cols) : "memory");
}

}  // namespace

// Returns true (and the launch status) when the tensor-core path takes the call.
bool dense_forward_tc_try(const float* x, int64_t ld_x, const float* W, const float* b, int64_t B, int K, int N, int act,
                          float* out, int64_t ld_out, cudaStream_t st, vms_status* status) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("VMS_DENSE_TC");
    disabled = (e && e[0] == '0') ? 1 : 0;
  }
  if (disabled) return false;
  if (B < 8192 || K < 16 || N < 16 || x == nullptr) return false;  // below ~64 tiles the FFMA kernel is faster
  if (K > 8 * kMaxKc) return false;
  if (K % 4 != 0 || ld_x % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15u) != 0) return false;
  const int Kp = (K + 7) & ~7, Np = (N + 15) & ~15;
  if (Np > 256) return false;
  const size_t smem = (size_t)(2 * TM * Kp + 2 * Np * Kp) * sizeof(float);
  if ((size_t)TM * N * sizeof(float) > (size_t)2 * TM * Kp * sizeof(float)) return false;  // staging must fit in the A tile
  if (smem + 1024 > (size_t)max_smem_optin()) return false;
  TcParams p = {};
  p.x = x; p.ld_x = ld_x; p.W = W; p.bias = b; p.out = out; p.ld_out = ld_out; p.B = B;
  p.K = K; p.N = N; p.Kp = Kp; p.Np = Np; p.act = act;
  p.tmem_cols = Np <= 16 ? 32u : Np <= 32 ? 64u : Np <= 64 ? 128u : Np <= 128 ? 256u : 512u;  // two accumulators
  cudaError_t e = cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const int64_t n_tiles = (B + TM - 1) / TM;
  const int grid = (int)(n_tiles < sm_count() ? n_tiles : sm_count());
  dense_tc_kernel<<<grid, TT, smem, st>>>(p);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    set_error("launch of dense_tc_kernel failed: %s", cudaGetErrorString(le));
    *status = VMS_ERR_CUDA;
    return true;
  }
  count_launch();
  *status = VMS_OK;
  return true;
}

}  // namespace vms
