// probe.cu -- machine-peak microbenchmarks the roofline fractions of this repo are quoted against (BASELINE.md section 2:
// "FP32 FFMA peak, TF32 tensor peak: not measured yet -- builder must microbenchmark on the box").
//
//   vms_probe_ffma   FP32 FFMA issue peak: every thread runs 8 independent fused-multiply-add chains out of registers, all
//                    SMs filled to 2048 resident threads; FLOPs = 2 x FMAs.  Nominal: #SM x 128 x 2 x clock.
//   vms_probe_mma    tcgen05.mma issue peak for the two operand kinds this library uses: kind::f16 with bfloat16 inputs
//                    (flow_tc.cu / elbo_tcf.cu: 3 x BF16 split) and kind::tf32 (gemm_tc.cu: 3 x TF32 split).  One CTA per SM,
//                    one thread issues back-to-back M x N x K MMAs on operands resident in shared memory (canonical K-major
//                    no-swizzle layout, contents irrelevant) into two alternating TMEM accumulators; nothing is loaded or
//                    stored inside the timed region, so this is the ceiling a kernel built on cta_group::1 MMAs of that
//                    shape can reach, not a GEMM.  The float32-equivalent ceiling of a split product is this number / 6
//                    (bf16) or / 3 (tf32).
// Both are timed with CUDA events on the launching stream after a warm-up launch; the result is the best of `reps`.
#include "common.cuh"

namespace vms {

namespace {

__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
        x7 = x0 + 7.f;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456f) out[0] = s;  // never true for the arguments used: keeps the chains alive
}

// the packed form: fma.rn.f32x2 (SASS FFMA2), two independent float32 FMAs per instruction on a 64-bit register pair
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__global__ void __launch_bounds__(256) ffma2_probe_kernel(float* out, int iters, float a, float b) {
  const float t = threadIdx.x * 1e-3f;
  unsigned long long x[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) x[u] = pack2f(t + u, t + u + 0.5f);
  const unsigned long long aa = pack2f(a, a), bb = pack2f(b, b);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int q = 0; q < 8; ++q) x[q] = ffma2(x[q], aa, bb);
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s ^= x[q];
  if (s == 0x123456789abcdefull) out[0] = 1.f;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long make_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

template <int KIND>  // 0: kind::f16 (bf16 inputs), 1: kind::tf32
__device__ __forceinline__ void mma_issue(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                          unsigned accumulate) {
  if (KIND == 0) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
  }
}

// operands: A [M rows][KS k-steps], B [N rows][KS k-steps] in the canonical K-major no-swizzle layout
// ([k / chunk][rows][16 bytes]); KS k-steps are cycled so the operand fetch pattern is that of a real product
template <int KIND>
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int M, int N, int n_mma, int* err) {
  extern __shared__ __align__(128) unsigned char smb[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_base_s;
  constexpr int KS = 4;                     // k-steps resident
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned a_bytes = (unsigned)M * KS * 32u, b_bytes = (unsigned)N * KS * 32u;  // 32 bytes of K per row and k-step
  for (unsigned e = tid; e < (a_bytes + b_bytes) / 4; e += 128) reinterpret_cast<unsigned*>(smb)[e] = 0u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tm = tmem_base_s;
  if (tid == 0) {
    // instruction descriptor: D = F32; A = B = BF16 (format 1) or TF32 (format 2); K-major; N >> 3, M >> 4
    const unsigned fmt = KIND == 0 ? 1u : 2u;
    const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
    const unsigned a0 = smem_u32(smb), b0 = a0 + a_bytes;
    // one k-step = two 16-byte chunks per row: LBO = chunk-column stride (rows x 16 B), SBO = 8-row group stride (128 B)
    const unsigned a_lbo = (unsigned)M * 16u, b_lbo = (unsigned)N * 16u;
    const unsigned a_step = (unsigned)M * 32u, b_step = (unsigned)N * 32u;
    for (int i = 0; i < n_mma; ++i) {
      const unsigned ks = (unsigned)(i & (KS - 1));
      mma_issue<KIND>(tm + (unsigned)((i & 1) * 256), make_desc(a0 + ks * a_step, a_lbo, 128u),
                      make_desc(b0 + ks * b_step, b_lbo, 128u), idesc, i >= 2 ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&mbar)) : "memory");
    const long long t0 = clock64();
    unsigned ok = 0;
    while (!ok) {
      asm volatile(
          "{\n"
          ".reg .pred P1;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
          "selp.u32 %0, 1, 0, P1;\n"
          "}\n"
          : "=r"(ok)
          : "r"(smem_u32(&mbar)), "r"(0u)
          : "memory");
      if (!ok && clock64() - t0 > 4000000000LL) {
        *err = 1;
        break;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}

template <typename F>
vms_status best_of(F launch, int reps, cudaStream_t st, double* best_ms) {
  cudaEvent_t e0, e1;
  VMS_CUDA(cudaEventCreate(&e0));
  VMS_CUDA(cudaEventCreate(&e1));
  double best = 1e30;
  for (int r = 0; r < reps + 1; ++r) {  // first launch is the warm-up
    VMS_CUDA(cudaEventRecord(e0, st));
    launch();
    VMS_CUDA(cudaEventRecord(e1, st));
    VMS_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    VMS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *best_ms = best;
  return VMS_OK;
}

}  // namespace

}  // namespace vms

using namespace vms;

extern "C" {

vms_status vms_probe_ffma(int iters, int reps, double* tflops, double* ms, vms_stream stream) {
  VMS_REQUIRE(iters >= 1 && reps >= 1 && tflops && ms, VMS_ERR_INVALID_ARG, "probe_ffma: bad arguments");
  float* out = nullptr;
  VMS_CUDA(cudaMalloc(&out, 16));
  const int grid = sm_count() * 8;  // 8 x 256 = 2048 resident threads per SM
  cudaStream_t st = as_stream(stream);
  vms_status s = best_of([&] { ffma_probe_kernel<<<grid, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); count_launch(); }, reps, st, ms);
  cudaFree(out);
  if (s != VMS_OK) return s;
  VMS_LAUNCH_CHECK("ffma_probe_kernel");
  const double fma = (double)grid * 256.0 * (double)iters * 16.0 * 8.0;
  *tflops = 2.0 * fma / (*ms * 1e-3) / 1e12;
  return VMS_OK;
}

vms_status vms_probe_ffma2(int iters, int reps, double* tflops, double* ms, vms_stream stream) {
  VMS_REQUIRE(iters >= 1 && reps >= 1 && tflops && ms, VMS_ERR_INVALID_ARG, "probe_ffma2: bad arguments");
  float* out = nullptr;
  VMS_CUDA(cudaMalloc(&out, 16));
  const int grid = sm_count() * 8;
  cudaStream_t st = as_stream(stream);
  vms_status s = best_of([&] { ffma2_probe_kernel<<<grid, 256, 0, st>>>(out, iters, 0.999f, 1e-3f); count_launch(); }, reps, st, ms);
  cudaFree(out);
  if (s != VMS_OK) return s;
  VMS_LAUNCH_CHECK("ffma2_probe_kernel");
  const double fma = (double)grid * 256.0 * (double)iters * 16.0 * 8.0 * 2.0;  // two FMAs per instruction
  *tflops = 2.0 * fma / (*ms * 1e-3) / 1e12;
  return VMS_OK;
}

vms_status vms_probe_mma(int kind, int M, int N, int n_mma, int reps, double* tflops, double* ms, vms_stream stream) {
  VMS_REQUIRE((kind == 0 || kind == 1) && (M == 64 || M == 128) && N >= 16 && N <= 256 && N % 16 == 0 && n_mma >= 2 &&
                  reps >= 1 && tflops && ms,
              VMS_ERR_INVALID_ARG, "probe_mma: kind in {0 bf16, 1 tf32}, M in {64, 128}, N a multiple of 16 up to 256");
  int* err = nullptr;
  VMS_CUDA(cudaMalloc(&err, sizeof(int)));
  VMS_CUDA(cudaMemset(err, 0, sizeof(int)));
  const size_t smem = (size_t)(M + N) * 4 * 32 + 128;
  auto k0 = mma_probe_kernel<0>;
  auto k1 = mma_probe_kernel<1>;
  VMS_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VMS_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = sm_count();
  cudaStream_t st = as_stream(stream);
  vms_status s = best_of(
      [&] {
        if (kind == 0) k0<<<grid, 128, smem, st>>>(M, N, n_mma, err);
        else k1<<<grid, 128, smem, st>>>(M, N, n_mma, err);
        count_launch();
      },
      reps, st, ms);
  int herr = 0;
  cudaMemcpy(&herr, err, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(err);
  if (s != VMS_OK) return s;
  VMS_LAUNCH_CHECK("mma_probe_kernel");
  VMS_REQUIRE(herr == 0, VMS_ERR_CUDA, "probe_mma: an MMA completion wait ran into its bound");
  const double k_per = kind == 0 ? 16.0 : 8.0;
  *tflops = 2.0 * (double)M * N * k_per * (double)n_mma * grid / (*ms * 1e-3) / 1e12;
  return VMS_OK;
}

}  // extern "C"
