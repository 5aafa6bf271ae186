// dense.cu -- K2/K3: small dense layers (FCDeepNN, spline conditioners, pre-masked MADE layers), forward and
// reverse mode, as FP32 FFMA GEMMs.
//
// Replaces Keras Dense at mappings.py:107-121 (`FCDeepNN.build` / `.call` :151-153), flows.py:136-152 and :185-196
// (`SplineBijector` d1 + three heads), and the masked Dense layers inside tfp AutoregressiveNetwork
// (flows.py:454-487, dists.py:301-305), plus TF autodiff through them.
//
// Why FFMA and not tcgen05 here: the reference is float32 and parity is 1e-5 relative; contraction lengths are
// 1..200 and output widths 4..190, and at the named batch (4096 rows / 148 SMs = 28 rows per SM) no CTA owns a
// 128-row MMA tile.  The shapes are latency-bound, so the kernels are built for short critical paths: a 32-row tile
// per CTA (128+ CTAs at batch 4096), the WHOLE contraction staged in shared memory in at most two chunks, one
// barrier pair per chunk, 4x4 register tiles fed by 128-bit shared loads.
#include "dense.cuh"

namespace vms {

// ------------------------------------------------------------------------------------------------ row-tile GEMM
constexpr int TM = 32, TN = 64, RT = 128, KC = 112;  // KC * (TM+4 + TN+4) * 4 B = 46.6 KB of static shared memory

__device__ __forceinline__ float act_grad(float out, int act) {
  return act == VMS_ACT_RELU ? (out > 0.f ? 1.f : 0.f) : (act == VMS_ACT_TANH ? 1.f - out * out : 1.f);
}

__global__ void __launch_bounds__(RT) gemm_rowtile_kernel(const RowTileParams p) {
  __shared__ __align__(16) float As[KC][TM + 4];
  __shared__ __align__(16) float Bs[KC][TN + 4];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (int seg = 0; seg < 2; ++seg) {
    const int Kseg = seg == 0 ? p.K : p.K2;
    if (Kseg <= 0) continue;
    const float* A = seg == 0 ? p.A : p.A2;
    const int64_t lda = seg == 0 ? p.lda : p.lda2;
    const float* Bm = seg == 0 ? p.Bm : p.B2;
    const int64_t ldb = seg == 0 ? p.ldb : p.ldb2;
    const bool tb = seg == 0 && p.tb;
    const bool ones = seg == 0 && p.a_ones;
    const int a_act = seg == 0 ? p.a_act : 0;
    for (int k0 = 0; k0 < Kseg; k0 += KC) {
      const int kc = min(KC, Kseg - k0);
      // A chunk: rows are contiguous along k => lanes walk k, warps walk rows; stored k-major for the 4-row loads
      for (int m = warp; m < TM; m += RT / 32) {
        const int gm = m0 + m;
        for (int k = lane; k < kc; k += 32) {
          float v = 0.f;
          if (gm < p.M) {
            if (ones) v = 1.f;
            else {
              v = A[(int64_t)gm * lda + k0 + k];
              if (a_act) v *= act_grad(p.Ao[(int64_t)gm * p.ldao + k0 + k], a_act);
            }
          }
          As[k][m] = v;
        }
      }
      if (!tb) {  // B rows contiguous along n
        for (int k = warp; k < kc; k += RT / 32) {
          const float* src = Bm + (int64_t)(k0 + k) * ldb + n0;
          for (int n = lane; n < TN; n += 32) Bs[k][n] = (n0 + n < p.N) ? src[n] : 0.f;
        }
      } else {  // B'(k, n) = Bm[n * ldb + k]: contiguous along k
        for (int n = warp; n < TN; n += RT / 32) {
          const bool ok = n0 + n < p.N;
          const float* src = Bm + (int64_t)(n0 + n) * ldb + k0;
          for (int k = lane; k < kc; k += 32) Bs[k][n] = ok ? src[k] : 0.f;
        }
      }
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < kc; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(av[i], bv[jn], acc[i][jn]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int n = n0 + tx * 4 + jn;
      if (n >= p.N) continue;
      float v = acc[i][jn];
      if (p.bias) v += p.bias[n];
      if (p.act == VMS_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == VMS_ACT_TANH) v = tanhf(v);
      float* dst = p.C + (int64_t)m * p.ldc + n;
      *dst = p.accumulate ? *dst + v : v;
    }
  }
}

vms_status gemm_rowtile(const RowTileParams& p, cudaStream_t st) {
  if (p.M <= 0 || p.N <= 0) return VMS_OK;
  dim3 grid((p.N + TN - 1) / TN, (p.M + TM - 1) / TM);
  gemm_rowtile_kernel<<<grid, RT, 0, st>>>(p);
  VMS_LAUNCH_CHECK("gemm_rowtile_kernel");
  return VMS_OK;
}

// ------------------------------------------------------------------------------------------------ weight gradient
constexpr int WM = 64, WN = 64, WT = 256, WR = 32;  // output tile 64 x 64, 32 batch rows per stage

__global__ void __launch_bounds__(WT) gemm_wgrad_kernel(const WgradParams p) {
  __shared__ __align__(16) float Xs[WR][WM + 4];
  __shared__ __align__(16) float Gs[WR][WN + 4];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int i0 = blockIdx.y * WM, n0 = blockIdx.x * WN;
  const int ty = t >> 4, tx = t & 15;
  const int64_t rows_per_split = (p.B + p.splits - 1) / p.splits;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(p.B, r_begin + rows_per_split);
  float acc[4][4] = {};
  for (int64_t r0 = r_begin; r0 < r_end; r0 += WR) {
    const int rc = (int)min((int64_t)WR, r_end - r0);
    for (int r = warp; r < WR; r += WT / 32) {
      const bool ok = r < rc;
      const int64_t gr = r0 + r;
      for (int c = lane; c < WM; c += 32) {
        const int i = i0 + c;
        float v = 0.f;
        if (ok && i <= p.Kin) v = (i == p.Kin || p.x == nullptr) ? 1.f : p.x[gr * p.ldx + i];
        Xs[r][c] = v;
      }
      for (int c = lane; c < WN; c += 32) {
        const int n = n0 + c;
        float v = 0.f;
        if (ok && n < p.N) {
          v = p.g[gr * p.ldg + n];
          if (p.act) v *= act_grad(p.out[gr * p.ldo + n], p.act);
        }
        Gs[r][c] = v;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WR; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[r][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Gs[r][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(av[i], bv[jn], acc[i][jn]);
    }
    __syncthreads();
  }
  float* C = p.part + (int64_t)blockIdx.z * p.split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = i0 + ty * 4 + i;
    if (gi > p.Kin) continue;
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int n = n0 + tx * 4 + jn;
      if (n < p.N) C[(int64_t)gi * p.N + n] = acc[i][jn];
    }
  }
}

vms_status gemm_wgrad(const WgradParams& p, cudaStream_t st) {
  if (p.B <= 0 || p.N <= 0) return VMS_OK;
  dim3 grid((p.N + WN - 1) / WN, (p.Kin + 1 + WM - 1) / WM, p.splits);
  gemm_wgrad_kernel<<<grid, WT, 0, st>>>(p);
  VMS_LAUNCH_CHECK("gemm_wgrad_kernel");
  return VMS_OK;
}

// out[i] (+)= scale * sum_s part[s][i];  fixed summation order => deterministic.  Two destination segments
// (weights then bias) so a [K+1, N] partial lands in separate g_W / g_b buffers.
// A CTA covers 32 consecutive outputs with 8 warps: warp g sums the partials s = g, g + 8, ... in ascending order (128-byte
// coalesced rows), the eight sums meet as ((0+1)+(2+3)) + ((4+5)+(6+7)).  (One thread per output walking all partials
// -- 148 dependent L2 round trips on a 43-CTA grid for a coupling block's 11 k gradients -- took 18 us per call.)
constexpr int kSumG = 8;
__global__ void __launch_bounds__(32 * kSumG) sum_partials_kernel(const float* __restrict__ part, int n_partials,
                                                                  int64_t stride, int64_t n0, float* out0, int64_t n1,
                                                                  float* out1, float scale, int accumulate) {
  __shared__ float red[kSumG][33];
  const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const int g = threadIdx.y;
  float s = 0.f;
  if (i < n0 + n1) {
#pragma unroll 4
    for (int j = g; j < n_partials; j += kSumG) s += __ldg(part + (int64_t)j * stride + i);
  }
  red[g][threadIdx.x] = s;
  __syncthreads();
  if (g != 0 || i >= n0 + n1) return;
  const int t = threadIdx.x;
  s = ((red[0][t] + red[1][t]) + (red[2][t] + red[3][t])) + ((red[4][t] + red[5][t]) + (red[6][t] + red[7][t]));
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
  sum_partials_kernel<<<(unsigned)((n + 31) / 32), dim3(32, kSumG), 0, st>>>(part, n_partials, stride, n0, out0, n1, out1,
                                                                            scale, accumulate);
  VMS_LAUNCH_CHECK("sum_partials_kernel");
  return VMS_OK;
}

int dense_splits(int64_t B) {
  int64_t s = (B + 127) / 128;
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
  return vms::dense_forward_impl(x, ld_x, W, b, B, K, N, act, cond, ld_c, Wc, C, out, ld_out, as_stream(stream), true);
}

}  // extern "C"

namespace vms {
// allow_tc = false keeps the FP32 FFMA kernel at every batch size: the unfused ELBO plan uses it so that it stays the
// float32 cross-check of the fused kernel (the 3 x TF32 tensor-core path carries ~5e-6 of the dot product's scale,
// which ill-conditioned splines amplify past the 1e-5 budget of that comparison).
vms_status dense_forward_impl(const float* x, int64_t ld_x, const float* W, const float* b, int64_t B, int K, int N,
                              int act, const float* cond, int64_t ld_c, const float* Wc, int C, float* out,
                              int64_t ld_out, cudaStream_t stream, bool allow_tc) {
  VMS_REQUIRE(B >= 0 && K >= 1 && N >= 1, VMS_ERR_SHAPE, "dense_forward: bad shape B=%lld K=%d N=%d", (long long)B, K, N);
  VMS_REQUIRE(B < (1LL << 31), VMS_ERR_SHAPE, "dense_forward: batch too large");
  VMS_REQUIRE(W && out, VMS_ERR_INVALID_ARG, "dense_forward: NULL W / out");
  VMS_REQUIRE(act >= 0 && act <= 2, VMS_ERR_INVALID_ARG, "dense_forward: unknown activation %d", act);
  VMS_REQUIRE(C == 0 || (cond != nullptr && Wc != nullptr), VMS_ERR_INVALID_ARG,
              "dense_forward: conditional_input missing");
  VMS_REQUIRE(x != nullptr || K == 1, VMS_ERR_INVALID_ARG, "dense_forward: NULL x is only valid for the ones input (K=1)");
  if (B == 0) return VMS_OK;
  if (allow_tc && C == 0 && x != nullptr) {  // large batches: tcgen05 3xTF32 GEMM (gemm_tc.cu); returns false when it does not apply
    vms_status s = VMS_OK;
    if (dense_forward_tc_try(x, ld_x, W, b, B, K, N, act, out, ld_out, stream, &s)) return s;
  }
  RowTileParams p = {};
  p.M = (int)B; p.N = N; p.K = K;
  p.A = x; p.lda = ld_x; p.a_ones = (x == nullptr);
  p.Bm = W; p.ldb = N;
  if (C > 0) { p.A2 = cond; p.lda2 = ld_c; p.B2 = Wc; p.ldb2 = N; p.K2 = C; }
  p.bias = b; p.act = act; p.C = out; p.ldc = ld_out;
  return gemm_rowtile(p, stream);
}
}  // namespace vms

extern "C" {

size_t vms_dense_backward_workspace(int64_t B, int K, int N, int C) {
  size_t rows = (size_t)(K + 1) + (size_t)(C > 0 ? C + 1 : 0);
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
    RowTileParams p = {};
    p.M = (int)B; p.N = K; p.K = N;
    p.A = g_out; p.lda = ld_g; p.Ao = out; p.ldao = ld_out; p.a_act = act;
    p.Bm = W; p.ldb = N; p.tb = 1;
    p.C = g_x; p.ldc = ld_gx; p.accumulate = accumulate_x;
    if ((s = gemm_rowtile(p, st))) return s;
  }
  if (g_cond) {
    VMS_REQUIRE(Wc && C > 0, VMS_ERR_INVALID_ARG, "dense_backward: Wc required for g_cond");
    RowTileParams p = {};
    p.M = (int)B; p.N = C; p.K = N;
    p.A = g_out; p.lda = ld_g; p.Ao = out; p.ldao = ld_out; p.a_act = act;
    p.Bm = Wc; p.ldb = N; p.tb = 1;
    p.C = g_cond; p.ldc = ld_gc;
    if ((s = gemm_rowtile(p, st))) return s;
  }
  const int splits = dense_splits(B);
  if (g_W || g_b) {  // [g_W; g_b] = [x^T; 1^T] @ (g_out * act'), split over the batch, partials summed in order
    VMS_REQUIRE(workspace, VMS_ERR_INVALID_ARG, "dense_backward: workspace required");
    WgradParams p = {};
    p.B = B; p.Kin = K; p.N = N;
    p.x = x; p.ldx = ld_x;
    p.g = g_out; p.ldg = ld_g; p.out = out; p.ldo = ld_out; p.act = act;
    p.part = (float*)workspace; p.split_stride = (int64_t)(K + 1) * N; p.splits = splits;
    if ((s = gemm_wgrad(p, st))) return s;
    if ((s = sum_partials_launch((const float*)workspace, splits, p.split_stride, (int64_t)K * N, g_W, N, g_b, 1.f,
                                 accumulate, st)))
      return s;
  }
  if (g_Wc) {
    VMS_REQUIRE(workspace && cond && C > 0, VMS_ERR_INVALID_ARG, "dense_backward: cond / workspace required for g_Wc");
    float* ws = (float*)workspace + (size_t)splits * (K + 1) * N;
    WgradParams p = {};
    p.B = B; p.Kin = C; p.N = N;
    p.x = cond; p.ldx = ld_c;
    p.g = g_out; p.ldg = ld_g; p.out = out; p.ldo = ld_out; p.act = act;
    p.part = ws; p.split_stride = (int64_t)(C + 1) * N; p.splits = splits;
    if ((s = gemm_wgrad(p, st))) return s;
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
