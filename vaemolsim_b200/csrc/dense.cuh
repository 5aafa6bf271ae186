// dense.cuh -- internal interface of the FFMA GEMMs used by dense.cu and the fused ELBO plan (elbo.cu).
#pragma once
#include "common.cuh"

namespace vms {

// C[M,N] (+)= act( A'[M,K] @ B'[K,N] (+ A2[M,K2] @ B2[K2,N]) + bias ),   one row tile of 32 x 64 per CTA.
//   A'(m,k) = A[m*lda + k] * act'(Ao[m*ldao + k])   (a_act != 0: reverse mode through the layer's activation)
//           = 1                                      (a_ones: the ones((B,1)) conditioner input, flows.py:184-185)
//   B'(k,n) = Bm[k*ldb + n]  or, with tb,  Bm[n*ldb + k]   (weight matrix used transposed for the input gradient)
struct RowTileParams {
  int M, N, K;
  const float* A; int64_t lda; int a_ones;
  const float* Ao; int64_t ldao; int a_act;
  const float* Bm; int64_t ldb; int tb;
  const float* A2; int64_t lda2; const float* B2; int64_t ldb2; int K2;
  const float* bias; int act; int accumulate;
  float* C; int64_t ldc;
};
vms_status gemm_rowtile(const RowTileParams& p, cudaStream_t st);

// Weight + bias gradient of one Dense layer, split over the batch:
//   part[s][i][n] = sum_{r in split s} X'(r,i) * G'(r,n),   i in [0, Kin] (row Kin = ones: the bias gradient),
//   X'(r,i) = x[r*ldx + i] (or 1 when x == NULL),  G'(r,n) = g[r*ldg + n] * act'(out[r*ldo + n]).
// Partials are written to part + s*split_stride as a dense [Kin+1, N] block (the Keras "kernel then bias" order).
struct WgradParams {
  int64_t B; int Kin, N;
  const float* x; int64_t ldx;
  const float* g; int64_t ldg;
  const float* out; int64_t ldo; int act;
  float* part; int64_t split_stride; int splits;
};
vms_status gemm_wgrad(const WgradParams& p, cudaStream_t st);

vms_status sum_partials_launch(const float* part, int n_partials, int64_t stride, int64_t n0, float* out0, int64_t n1,
                               float* out1, float scale, int accumulate, cudaStream_t st);
int dense_splits(int64_t B);

vms_status dense_forward_impl(const float* x, int64_t ld_x, const float* W, const float* b, int64_t B, int K, int N,
                              int act, const float* cond, int64_t ld_c, const float* Wc, int C, float* out,
                              int64_t ld_out, cudaStream_t stream, bool allow_tc);

// gemm_tc.cu: tcgen05 (3 x TF32) Dense forward for large batches
bool dense_forward_tc_try(const float* x, int64_t ld_x, const float* W, const float* b, int64_t B, int K, int N, int act,
                          float* out, int64_t ld_out, cudaStream_t st, vms_status* status);

}  // namespace vms
