// dense.cu -- K2/K3: small dense layers (FCDeepNN, spline conditioners, pre-masked MADE layers), forward and
// reverse mode, as FP32 FFMA tiled GEMMs.
//
// Replaces Keras Dense at mappings.py:107-121 (`FCDeepNN.build` / `.call` :151-153), flows.py:136-152 and :185-196
// (`SplineBijector` d1 + three heads), and the masked Dense layers inside tfp AutoregressiveNetwork
// (flows.py:454-487, dists.py:301-305), plus TF autodiff through them.
//
// Why FFMA and not tcgen05 here: the reference is float32 and parity is 1e-5 relative; contraction lengths are
// 1..200 and output widths 4..190, and at the named batch (4096 rows / 148 SMs = 28 rows per SM) no CTA owns a
// 128-row MMA tile.  The shapes are latency-bound, not throughput-bound ("GEMMs").
#include "dense.cuh"

namespace vms {

constexpr int BM = 64, BN = 64, BK = 16, GT = 256;

__device__ __forceinline__ float act_grad(float out, int act) {
  return act == VMS_ACT_RELU ? (out > 0.f ? 1.f : 0.f) : (act == VMS_ACT_TANH ? 1.f - out * out : 1.f);
}

__device__ __forceinline__ float load_a(const GemmParams& p, int m, int k) {
  if (m >= p.M || k >= p.K) return 0.f;
  if (p.a_ones) return 1.f;
  if (p.ta) return m == p.a_ones_row ? 1.f : p.A[(int64_t)k * p.lda + m];
  float v = p.A[(int64_t)m * p.lda + k];
  if (p.a_act) v *= act_grad(p.Ao[(int64_t)m * p.ldao + k], p.a_act);
  return v;
}
__device__ __forceinline__ float load_b(const GemmParams& p, int k, int n) {
  if (n >= p.N || k >= p.K) return 0.f;
  if (p.tb) return p.Bm[(int64_t)n * p.ldb + k];
  float v = p.Bm[(int64_t)k * p.ldb + n];
  if (p.b_act) v *= act_grad(p.Bo[(int64_t)k * p.ldbo + n], p.b_act);
  return v;
}

__global__ void __launch_bounds__(GT) gemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tm = (t / 16) * 4, tn = (t % 16) * 4;
  float acc[4][4] = {};
  int k_begin = 0, k_end = p.K;
  if (p.k_per_split > 0) {
    k_begin = blockIdx.z * p.k_per_split;
    k_end = min(p.K, k_begin + p.k_per_split);
  }
  for (int phase = 0; phase < 2; ++phase) {
    if (phase == 1 && (p.K2 <= 0 || blockIdx.z != 0)) break;
    const int kb = phase ? 0 : k_begin, ke = phase ? p.K2 : k_end;
    for (int k0 = kb; k0 < ke; k0 += BK) {
      // stage A tile [BM x BK] and B tile [BK x BN]; thread mapping follows the contiguous axis of each operand
#pragma unroll
      for (int i = 0; i < (BM * BK) / GT; ++i) {
        int e = t + i * GT, m, k;
        if (phase == 0 && p.ta) { m = e % BM; k = e / BM; } else { k = e % BK; m = e / BK; }
        float v;
        if (phase == 0) v = (k0 + k < ke) ? load_a(p, m0 + m, k0 + k) : 0.f;
        else v = (m0 + m < p.M && k0 + k < ke) ? p.A2[(int64_t)(m0 + m) * p.lda2 + k0 + k] : 0.f;
        As[k][m] = v;
      }
#pragma unroll
      for (int i = 0; i < (BN * BK) / GT; ++i) {
        int e = t + i * GT, n, k;
        if (phase == 0 && p.tb) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
        float v;
        if (phase == 0) v = (k0 + k < ke) ? load_b(p, k0 + k, n0 + n) : 0.f;
        else v = (n0 + n < p.N && k0 + k < ke) ? p.B2[(int64_t)(k0 + k) * p.ldb2 + n0 + n] : 0.f;
        Bs[k][n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][tm]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* C = p.C + (p.k_per_split > 0 ? (int64_t)blockIdx.z * p.split_stride : 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.act == VMS_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == VMS_ACT_TANH) v = tanhf(v);
      float* dst = C + (int64_t)m * p.ldc + n;
      *dst = p.accumulate ? *dst + v : v;
    }
  }
}

vms_status gemm_launch(const GemmParams& p, int splits, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return VMS_OK;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, splits > 0 ? splits : 1);
  gemm_kernel<<<grid, GT, 0, st>>>(p);
  VMS_LAUNCH_CHECK("gemm_kernel");
  return VMS_OK;
}

// out[i] (+)= scale * sum_s part[s][i];  fixed summation order => deterministic.  Two destination segments
// (weights then bias) so a [K+1, N] partial lands in separate g_W / g_b buffers.
__global__ void sum_partials_kernel(const float* __restrict__ part, int n_partials, int64_t stride, int64_t n0,
                                    float* out0, int64_t n1, float* out1, float scale, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0 + n1) return;
  float s = 0.f;
  for (int j = 0; j < n_partials; ++j) s += part[(int64_t)j * stride + i];
  s *= scale;
  float* base = i < n0 ? out0 : out1;
  if (!base) return;
  float* dst = i < n0 ? out0 + i : out1 + (i - n0);
  *dst = accumulate ? *dst + s : s;
}

vms_status sum_partials_launch(const float* part, int n_partials, int64_t stride, int64_t n0, float* out0, int64_t n1,
                               float* out1, float scale, int accumulate, cudaStream_t st) {
  int64_t n = n0 + n1;
  if (n <= 0) return VMS_OK;
  sum_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, n_partials, stride, n0, out0, n1, out1, scale,
                                                                  accumulate);
  VMS_LAUNCH_CHECK("sum_partials_kernel");
  return VMS_OK;
}

int dense_splits(int64_t B) {
  int64_t s = (B + 255) / 256;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

__global__ void periodic_kernel(const float* __restrict__ x, int64_t B, int D, const uint8_t* __restrict__ per,
                                float* __restrict__ out) {
  // output layout: [non-periodic ..., cos(periodic ...), sin(periodic ...)]
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int n_p = 0;
  for (int d = 0; d < D; ++d) n_p += per[d] ? 1 : 0;
  const int n_np = D - n_p;
  int inp = 0, ip = 0;
  const float* xr = x + b * D;
  float* o = out + b * (D + n_p);
  for (int d = 0; d < D; ++d) {
    float v = xr[d];
    if (per[d]) {
      o[n_np + ip] = cosf(v);
      o[n_np + n_p + ip] = sinf(v);
      ++ip;
    } else {
      o[inp++] = v;
    }
  }
}

}  // namespace vms

using namespace vms;

extern "C" {

vms_status vms_dense_forward(const float* x, int64_t ld_x, const float* W, const float* b, int64_t B, int K, int N,
                             int act, const float* cond, int64_t ld_c, const float* Wc, int C, float* out,
                             int64_t ld_out, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && K >= 1 && N >= 1, VMS_ERR_SHAPE, "dense_forward: bad shape B=%lld K=%d N=%d", (long long)B, K, N);
  VMS_REQUIRE(B < (1LL << 31), VMS_ERR_SHAPE, "dense_forward: batch too large");
  VMS_REQUIRE(W && out, VMS_ERR_INVALID_ARG, "dense_forward: NULL W / out");
  VMS_REQUIRE(act >= 0 && act <= 2, VMS_ERR_INVALID_ARG, "dense_forward: unknown activation %d", act);
  VMS_REQUIRE((C > 0) == (cond != nullptr && Wc != nullptr) || C == 0, VMS_ERR_INVALID_ARG,
              "dense_forward: conditional_input missing");
  if (B == 0) return VMS_OK;
  GemmParams p = {};
  p.M = (int)B; p.N = N; p.K = K;
  p.A = x; p.lda = ld_x; p.a_ones = (x == nullptr); p.a_ones_row = -1;
  VMS_REQUIRE(x != nullptr || K == 1, VMS_ERR_INVALID_ARG, "dense_forward: NULL x is only valid for the ones input (K=1)");
  p.Bm = W; p.ldb = N;
  if (C > 0) { p.A2 = cond; p.lda2 = ld_c; p.B2 = Wc; p.ldb2 = N; p.K2 = C; }
  p.bias = b; p.act = act; p.C = out; p.ldc = ld_out;
  return gemm_launch(p, 0, as_stream(stream));
}

size_t vms_dense_backward_workspace(int64_t B, int K, int N, int C) {
  size_t rows = (size_t)(K + 1) + (size_t)(C > 0 ? C : 0);
  return (size_t)dense_splits(B) * rows * (size_t)N * sizeof(float);
}

vms_status vms_dense_backward(const float* x, int64_t ld_x, const float* W, int64_t B, int K, int N, int act,
                              const float* out, int64_t ld_out, const float* g_out, int64_t ld_g, const float* cond,
                              int64_t ld_c, const float* Wc, int C, float* g_x, int64_t ld_gx, int accumulate_x,
                              float* g_W, float* g_b, float* g_cond, int64_t ld_gc, float* g_Wc, int accumulate,
                              void* workspace, vms_stream stream) {
  VMS_REQUIRE(B >= 0 && K >= 1 && N >= 1, VMS_ERR_SHAPE, "dense_backward: bad shape");
  VMS_REQUIRE(B < (1LL << 31), VMS_ERR_SHAPE, "dense_backward: batch too large");
  VMS_REQUIRE(g_out, VMS_ERR_INVALID_ARG, "dense_backward: NULL g_out");
  VMS_REQUIRE(act == VMS_ACT_NONE || out, VMS_ERR_INVALID_ARG, "dense_backward: saved output required for act %d", act);
  if (B == 0) return VMS_OK;
  cudaStream_t st = as_stream(stream);
  vms_status s;
  if (g_x) {  // g_x = (g_out * act') @ W^T
    VMS_REQUIRE(W, VMS_ERR_INVALID_ARG, "dense_backward: W required for g_x");
    GemmParams p = {};
    p.M = (int)B; p.N = K; p.K = N;
    p.A = g_out; p.lda = ld_g; p.Ao = out; p.ldao = ld_out; p.a_act = act; p.a_ones_row = -1;
    p.Bm = W; p.ldb = N; p.tb = 1;
    p.C = g_x; p.ldc = ld_gx; p.accumulate = accumulate_x;
    if ((s = gemm_launch(p, 0, st))) return s;
  }
  if (g_cond) {
    VMS_REQUIRE(Wc && C > 0, VMS_ERR_INVALID_ARG, "dense_backward: Wc required for g_cond");
    GemmParams p = {};
    p.M = (int)B; p.N = C; p.K = N;
    p.A = g_out; p.lda = ld_g; p.Ao = out; p.ldao = ld_out; p.a_act = act; p.a_ones_row = -1;
    p.Bm = Wc; p.ldb = N; p.tb = 1;
    p.C = g_cond; p.ldc = ld_gc;
    if ((s = gemm_launch(p, 0, st))) return s;
  }
  if (g_W || g_b) {  // [g_W; g_b] = [x^T; 1^T] @ (g_out * act'), split over the batch, partials summed in order
    VMS_REQUIRE(workspace, VMS_ERR_INVALID_ARG, "dense_backward: workspace required");
    const int splits = dense_splits(B);
    GemmParams p = {};
    p.M = K + 1; p.N = N; p.K = (int)B;
    p.A = x; p.lda = ld_x; p.ta = 1; p.a_ones = (x == nullptr); p.a_ones_row = K;
    p.Bm = g_out; p.ldb = ld_g; p.Bo = out; p.ldbo = ld_out; p.b_act = act;
    p.C = (float*)workspace; p.ldc = N;
    p.k_per_split = (int)((B + splits - 1) / splits);
    p.split_stride = (int64_t)(K + 1) * N;
    if ((s = gemm_launch(p, splits, st))) return s;
    if ((s = sum_partials_launch((const float*)workspace, splits, p.split_stride, (int64_t)K * N, g_W, N, g_b, 1.f,
                                 accumulate, st)))
      return s;
  }
  if (g_Wc) {
    VMS_REQUIRE(workspace && cond && C > 0, VMS_ERR_INVALID_ARG, "dense_backward: cond / workspace required for g_Wc");
    const int splits = dense_splits(B);
    float* ws = (float*)workspace + (size_t)splits * (K + 1) * N;
    GemmParams p = {};
    p.M = C; p.N = N; p.K = (int)B;
    p.A = cond; p.lda = ld_c; p.ta = 1; p.a_ones_row = -1;
    p.Bm = g_out; p.ldb = ld_g; p.Bo = out; p.ldbo = ld_out; p.b_act = act;
    p.C = ws; p.ldc = N;
    p.k_per_split = (int)((B + splits - 1) / splits);
    p.split_stride = (int64_t)C * N;
    if ((s = gemm_launch(p, splits, st))) return s;
    if ((s = sum_partials_launch(ws, splits, p.split_stride, (int64_t)C * N, g_Wc, 0, nullptr, 1.f, accumulate, st)))
      return s;
  }
  return VMS_OK;
}

vms_status vms_periodic_featurise(const float* x, int64_t B, int D, const uint8_t* periodic, float* out,
                                  vms_stream stream) {
  VMS_REQUIRE(B >= 0 && D >= 1, VMS_ERR_SHAPE, "periodic_featurise: bad shape");
  VMS_REQUIRE(x && periodic && out, VMS_ERR_INVALID_ARG, "periodic_featurise: NULL pointer");
  if (B == 0) return VMS_OK;
  periodic_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(x, B, D, periodic, out);
  VMS_LAUNCH_CHECK("periodic_kernel");
  return VMS_OK;
}

vms_status vms_sum_partials(const float* g, int n_partials, int64_t n, float scale, float* out, vms_stream stream) {
  VMS_REQUIRE(g && out && n_partials >= 1 && n >= 0, VMS_ERR_INVALID_ARG, "sum_partials: bad arguments");
  return sum_partials_launch(g, n_partials, n, n, out, 0, nullptr, scale, 0, as_stream(stream));
}

}  // extern "C"
