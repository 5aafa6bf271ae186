// mlp_stream.cuh -- internal interface of the large-batch FCDeepNN kernels (mlp_stream.cu) used by the ELBO plan.
#pragma once
#include "common.cuh"

namespace vms {

// out [B, Dout] = relu(in [B, Din] @ W0 [Din, H] + b0) @ W1 [H, Dout] + b1   (mappings.py:107-121, one hidden layer)
// reverse mode: g_out [B, Dout] -> optional g_in [B, Din]; per-CTA weight-gradient partials at
// part + cta * part_stride + {o_W0, o_b0, o_W1, o_b1}; every CTA of the grid writes its partial.
struct MlpArgs {
  int64_t B;
  int Din, H, Dout;
  const float *W0, *b0, *W1, *b1;
  const float* in; int64_t ld_in;
  float* out; int64_t ld_out;
  const float* g_out; int64_t ld_g;
  float* g_in; int64_t ld_gin;
  float* part; int64_t part_stride, o_W0, o_b0, o_W1, o_b1;
};

bool mlp_stream_supported(int Din, int H, int Dout);
vms_status mlp_stream_forward(const MlpArgs& a, int grid, cudaStream_t st);
vms_status mlp_stream_backward(const MlpArgs& a, int grid, cudaStream_t st);

}  // namespace vms
