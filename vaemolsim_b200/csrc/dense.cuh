// dense.cuh -- internal interface of the FFMA GEMM used by dense.cu and the fused ELBO plan (elbo.cu).
#pragma once
#include "common.cuh"

namespace vms {

struct GemmParams {
  int M, N, K;
  // A operand: value(m,k)
  const float* A; int64_t lda; int ta;          // ta: A[k*lda+m] else A[m*lda+k]
  const float* Ao; int64_t ldao; int a_act;     // multiply by act'(Ao[same index]) (never used with ta)
  int a_ones;                                   // A == 1 everywhere
  int a_ones_row;                               // with ta: row m == a_ones_row reads as 1 (bias gradient), -1 = off
  // B operand: value(k,n)
  const float* Bm; int64_t ldb; int tb;         // tb: B[n*ldb+k] else B[k*ldb+n]
  const float* Bo; int64_t ldbo; int b_act;     // multiply by act'(Bo[same index]) (never used with tb)
  // optional second product accumulated into the same tile (conditional input): A2[M,K2] @ B2[K2,N]
  const float* A2; int64_t lda2; const float* B2; int64_t ldb2; int K2;
  // epilogue
  const float* bias; int act; int accumulate;
  float* C; int64_t ldc;
  // split-K: blockIdx.z handles k in [z*k_per_split, ...); partial z written at C + z*split_stride
  int k_per_split; int64_t split_stride;
};

vms_status gemm_launch(const GemmParams& p, int splits, cudaStream_t st);
vms_status sum_partials_launch(const float* part, int n_partials, int64_t stride, int64_t n0, float* out0, int64_t n1,
                               float* out1, float scale, int accumulate, cudaStream_t st);
int dense_splits(int64_t B);

}  // namespace vms
