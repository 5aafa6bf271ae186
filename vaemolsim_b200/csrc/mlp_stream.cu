// mlp_stream.cu -- FCDeepNN with one relu hidden layer at large batch: in [B, Din] -> relu(in W0 + b0) [H] -> W1 + b1
// [B, Dout], forward and reverse mode, WITHOUT materialising the hidden layer in HBM.
//
// Replaces Keras Dense x 2 of mappings.py:107-121 / FCDeepNN.call :151-153 (encoder 6 -> 200 -> 4, decoder
// 2 -> 200 -> 12 of tests/test_models.py:161-170) and TF autodiff through them in the large-batch plan (elbo.cu,
// mode 2).  The per-layer kernels of dense.cu write and re-read the [B, 200] hidden layer (and its gradient) four
// times per step: 6.4 KB per configuration against 24 + 16 bytes of real input / output, and their row-tile / split-K
// GEMMs are built for batch 4096.  These layers are too thin for the tensor core (contraction lengths 2, 6; output
// widths 4, 12), so the kernels are FP32 FFMA with every operand of the inner loops a shared-memory broadcast:
//   forward          thread per row, loop over hidden units (weights: 16-byte broadcasts);
//   weight gradients thread per hidden unit, loop over the tile's rows (inputs and output gradients: 16-byte
//                    broadcasts), hidden activation recomputed, accumulators in registers for the CTA's lifetime,
//                    one partial per CTA (summed by the caller in a fixed order: deterministic);
//   input gradient   (decoder only) thread pair per row, loop over hidden units.
#include "mlp_stream.cuh"
#include <math.h>

namespace vms {

namespace {

constexpr int MT = 256;   // threads per CTA
constexpr int TR = 128;   // rows per tile of the backward kernel

// shared-memory weights: w0t [H][DINP] = W0^T with b0 in column Din (the input row carries a 1 there), w1 [H][DOUTP]
template <int DINP, int DOUTP>
__device__ __forceinline__ void stage_weights(const MlpArgs& a, float* w0t, float* w1) {
  for (int e = threadIdx.x; e < a.H * DINP; e += MT) {
    const int j = e / DINP, i = e - j * DINP;
    w0t[e] = i < a.Din ? __ldg(a.W0 + (size_t)i * a.H + j) : (i == a.Din ? __ldg(a.b0 + j) : 0.f);
  }
  for (int e = threadIdx.x; e < a.H * DOUTP; e += MT) {
    const int j = e / DOUTP, n = e - j * DOUTP;
    w1[e] = n < a.Dout ? __ldg(a.W1 + (size_t)j * a.Dout + n) : 0.f;
  }
}

template <int N>
__device__ __forceinline__ void ld_row(const float* s, float (&v)[N]) {
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(s + 4 * q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

template <int DINP, int DOUTP>
__global__ void __launch_bounds__(MT) mlp2_fwd_kernel(const MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* w0t = sm;
  float* w1 = w0t + a.H * DINP;
  stage_weights<DINP, DOUTP>(a, w0t, w1);
  __syncthreads();
  float b1[DOUTP];
#pragma unroll
  for (int n = 0; n < DOUTP; ++n) b1[n] = n < a.Dout ? __ldg(a.b1 + n) : 0.f;
  for (int64_t row = (int64_t)blockIdx.x * MT + threadIdx.x; row < a.B; row += (int64_t)gridDim.x * MT) {
    float x[DINP];
#pragma unroll
    for (int i = 0; i < DINP; ++i) x[i] = i < a.Din ? __ldg(a.in + row * a.ld_in + i) : (i == a.Din ? 1.f : 0.f);
    float acc[DOUTP];
#pragma unroll
    for (int n = 0; n < DOUTP; ++n) acc[n] = b1[n];
#pragma unroll 4
    for (int j = 0; j < a.H; ++j) {
      float w[DINP], u[DOUTP];
      ld_row<DINP>(w0t + j * DINP, w);
      ld_row<DOUTP>(w1 + j * DOUTP, u);
      float pre = 0.f;
#pragma unroll
      for (int i = 0; i < DINP; ++i) pre = fmaf(x[i], w[i], pre);
      const float h = fmaxf(pre, 0.f);
#pragma unroll
      for (int n = 0; n < DOUTP; ++n) acc[n] = fmaf(h, u[n], acc[n]);
    }
    float* o = a.out + row * a.ld_out;
#pragma unroll
    for (int n = 0; n < DOUTP; ++n)
      if (n < a.Dout) o[n] = acc[n];
  }
}

template <int DINP, int DOUTP, bool GIN>
__global__ void __launch_bounds__(MT) mlp2_bwd_kernel(const MlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* s_in = sm;                   // [TR][DINP], column Din = 1
  float* s_g = s_in + TR * DINP;      // [TR][DOUTP]
  float* w0t = s_g + TR * DOUTP;      // [H][DINP]   (input-gradient phase)
  float* w1 = w0t + a.H * DINP;       // [H][DOUTP]
  const int tid = threadIdx.x;
  const int H = a.H;
  if (GIN) stage_weights<DINP, DOUTP>(a, w0t, w1);
  // this thread's hidden unit
  float w0[DINP], w1r[DOUTP], dW0[DINP], dW1[DOUTP];
#pragma unroll
  for (int i = 0; i < DINP; ++i) {
    w0[i] = (tid < H && i < a.Din) ? __ldg(a.W0 + (size_t)i * H + tid) : ((tid < H && i == a.Din) ? __ldg(a.b0 + tid) : 0.f);
    dW0[i] = 0.f;
  }
#pragma unroll
  for (int n = 0; n < DOUTP; ++n) {
    w1r[n] = (tid < H && n < a.Dout) ? __ldg(a.W1 + (size_t)tid * a.Dout + n) : 0.f;
    dW1[n] = 0.f;
  }
  float db1 = 0.f;  // threads MT - DOUTP .. MT - 1 own one output bias each
  const int64_t n_tiles = (a.B + TR - 1) / TR;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TR;
    const int nr = (int)min((int64_t)TR, a.B - row0);
    __syncthreads();  // previous tile fully consumed
    for (int e = tid; e < TR * DINP; e += MT) {
      const int r = e / DINP, i = e - r * DINP;
      s_in[e] = (r < nr && i < a.Din) ? __ldg(a.in + (row0 + r) * a.ld_in + i) : (i == a.Din ? 1.f : 0.f);
    }
    for (int e = tid; e < TR * DOUTP; e += MT) {
      const int r = e / DOUTP, n = e - r * DOUTP;
      s_g[e] = (r < nr && n < a.Dout) ? __ldg(a.g_out + (row0 + r) * a.ld_g + n) : 0.f;
    }
    __syncthreads();
    // ---- weight gradients: thread = hidden unit
    if (tid < H) {
#pragma unroll 2
      for (int r = 0; r < TR; ++r) {
        float x[DINP], g[DOUTP];
        ld_row<DINP>(s_in + r * DINP, x);
        ld_row<DOUTP>(s_g + r * DOUTP, g);
        float pre = 0.f, t = 0.f;
#pragma unroll
        for (int i = 0; i < DINP; ++i) pre = fmaf(x[i], w0[i], pre);
#pragma unroll
        for (int n = 0; n < DOUTP; ++n) t = fmaf(g[n], w1r[n], t);
        const float h = fmaxf(pre, 0.f);
        const float gh = pre > 0.f ? t : 0.f;
#pragma unroll
        for (int n = 0; n < DOUTP; ++n) dW1[n] = fmaf(h, g[n], dW1[n]);
#pragma unroll
        for (int i = 0; i < DINP; ++i) dW0[i] = fmaf(x[i], gh, dW0[i]);
      }
    } else if (tid >= MT - DOUTP) {
      const int n = tid - (MT - DOUTP);
      float s = 0.f;
      for (int r = 0; r < TR; ++r) s += s_g[r * DOUTP + n];
      db1 += s;
    }
    // ---- input gradient: thread pair per row, each half of the hidden units
    if (GIN) {
      const int r = tid >> 1, half = tid & 1;
      float x[DINP], g[DOUTP], gi[DINP];
      ld_row<DINP>(s_in + r * DINP, x);
      ld_row<DOUTP>(s_g + r * DOUTP, g);
#pragma unroll
      for (int i = 0; i < DINP; ++i) gi[i] = 0.f;
      const int j0 = half * ((H + 1) / 2), j1 = half ? H : (H + 1) / 2;
#pragma unroll 2
      for (int j = j0; j < j1; ++j) {
        float w[DINP], u[DOUTP];
        ld_row<DINP>(w0t + j * DINP, w);
        ld_row<DOUTP>(w1 + j * DOUTP, u);
        float pre = 0.f, t = 0.f;
#pragma unroll
        for (int i = 0; i < DINP; ++i) pre = fmaf(x[i], w[i], pre);
#pragma unroll
        for (int n = 0; n < DOUTP; ++n) t = fmaf(g[n], u[n], t);
        const float gh = pre > 0.f ? t : 0.f;
#pragma unroll
        for (int i = 0; i < DINP; ++i) gi[i] = fmaf(gh, w[i], gi[i]);
      }
#pragma unroll
      for (int i = 0; i < DINP; ++i) gi[i] += __shfl_xor_sync(0xffffffffu, gi[i], 1);
      if (half == 0 && r < nr) {
        float* o = a.g_in + (row0 + r) * a.ld_gin;
#pragma unroll
        for (int i = 0; i < DINP; ++i)
          if (i < a.Din) o[i] = gi[i];
      }
    }
  }
  // per-CTA partial in the Keras order of the flat parameter buffer: W0 [Din, H], b0 [H], W1 [H, Dout], b1 [Dout]
  float* part = a.part + (size_t)blockIdx.x * a.part_stride;
  if (tid < H) {
#pragma unroll
    for (int i = 0; i < DINP; ++i) {
      if (i < a.Din) part[a.o_W0 + (size_t)i * H + tid] = dW0[i];
      else if (i == a.Din) part[a.o_b0 + tid] = dW0[i];
    }
#pragma unroll
    for (int n = 0; n < DOUTP; ++n)
      if (n < a.Dout) part[a.o_W1 + (size_t)tid * a.Dout + n] = dW1[n];
  } else if (tid >= MT - DOUTP) {
    const int n = tid - (MT - DOUTP);
    if (n < a.Dout) part[a.o_b1 + n] = db1;
  }
}

template <int DINP, int DOUTP>
vms_status launch_fwd(const MlpArgs& a, int grid, cudaStream_t st) {
  const size_t smem = (size_t)a.H * (DINP + DOUTP) * sizeof(float);
  VMS_CUDA(cudaFuncSetAttribute(mlp2_fwd_kernel<DINP, DOUTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mlp2_fwd_kernel<DINP, DOUTP><<<grid, MT, smem, st>>>(a);
  VMS_LAUNCH_CHECK("mlp2_fwd_kernel");
  return VMS_OK;
}

template <int DINP, int DOUTP>
vms_status launch_bwd(const MlpArgs& a, int grid, cudaStream_t st) {
  const size_t smem = (size_t)(TR + a.H) * (DINP + DOUTP) * sizeof(float);
  if (a.g_in) {
    VMS_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel<DINP, DOUTP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp2_bwd_kernel<DINP, DOUTP, true><<<grid, MT, smem, st>>>(a);
  } else {
    VMS_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel<DINP, DOUTP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp2_bwd_kernel<DINP, DOUTP, false><<<grid, MT, smem, st>>>(a);
  }
  VMS_LAUNCH_CHECK("mlp2_bwd_kernel");
  return VMS_OK;
}

// DINP = round_up(Din + 1, 4) in {4, 8}, DOUTP = round_up(Dout, 4) in {4, 8, 12, 16}
template <bool BWD>
vms_status dispatch(const MlpArgs& a, int grid, cudaStream_t st) {
  const int dinp = (a.Din + 1 + 3) / 4 * 4, doutp = (a.Dout + 3) / 4 * 4;
#define VMS_MLP_CASE(DI, DO) \
  if (dinp == DI && doutp == DO) return BWD ? launch_bwd<DI, DO>(a, grid, st) : launch_fwd<DI, DO>(a, grid, st);
  VMS_MLP_CASE(4, 4) VMS_MLP_CASE(4, 8) VMS_MLP_CASE(4, 12) VMS_MLP_CASE(4, 16)
  VMS_MLP_CASE(8, 4) VMS_MLP_CASE(8, 8) VMS_MLP_CASE(8, 12) VMS_MLP_CASE(8, 16)
#undef VMS_MLP_CASE
  set_error("mlp_stream: unsupported widths Din=%d Dout=%d", a.Din, a.Dout);
  return VMS_ERR_UNSUPPORTED;
}

}  // namespace

bool mlp_stream_supported(int Din, int H, int Dout) { return Din >= 1 && Din <= 7 && Dout >= 1 && Dout <= 16 && H >= 1 && H <= MT - 16; }

vms_status mlp_stream_forward(const MlpArgs& a, int grid, cudaStream_t st) {
  VMS_REQUIRE(mlp_stream_supported(a.Din, a.H, a.Dout) && a.in && a.out && a.W0 && a.b0 && a.W1 && a.b1 && a.B >= 1 && grid >= 1,
              VMS_ERR_INVALID_ARG, "mlp_stream_forward: bad arguments");
  return dispatch<false>(a, grid, st);
}

vms_status mlp_stream_backward(const MlpArgs& a, int grid, cudaStream_t st) {
  VMS_REQUIRE(mlp_stream_supported(a.Din, a.H, a.Dout) && a.in && a.g_out && a.W0 && a.b0 && a.W1 && a.part && a.B >= 1 && grid >= 1,
              VMS_ERR_INVALID_ARG, "mlp_stream_backward: bad arguments");
  return dispatch<true>(a, grid, st);
}

}  // namespace vms
