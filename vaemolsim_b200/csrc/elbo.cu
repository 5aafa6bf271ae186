// elbo.cu -- whole-VAE ELBO forward / forward+backward for the model family of the reference's tests (C1 / C2).
//
// Replaces the Python/TF orchestration of
//   models.py:289-322  VAE.call   (encoder -> sample -> prior -> regulariser -> decoder)
//   models.py:206-229  MappingToDistribution.call (FCDeepNN + distribution layer)
//   dists.py:414-439   FlowedDistribution.call + tfp TransformedDistribution.log_prob
//   flows.py:281-355   RQSSplineRealNVP (chain of RealNVP blocks, density direction = chain inverse)
//   losses.py:58, :253 LogProbLoss (batch mean) and KLDivergenceEstimate
// and TF autodiff (GradientTape) through all of it.
//
// The plan owns every intermediate (activations kept for the backward pass, split-K gradient partials) in HBM --
// sized once for max_batch -- and replays the whole step as ONE CUDA graph per (batch, pointer set): the step is
// ~60 small kernels whose launch overhead would otherwise dominate at the named batch of 4096.
#include "elbo_plan.cuh"
#include "peer.cuh"
#include "flow_tc.cuh"
#include "mlp_stream.cuh"
#include <math.h>
#include <stdlib.h>

namespace vms {

vms_status rqs_prepare();  // rqs.cu

// ------------------------------------------------------------------------------------------------ small kernels
// scalars[0] = loss = nll + w kl, [1] = nll = mean(-logpx), [2] = kl = mean(logq - logpz); fixed-order reduction
__global__ void __launch_bounds__(256) elbo_partial_kernel(const float* __restrict__ logq, const float* __restrict__ logpz,
                                                           const float* __restrict__ logpx, int64_t B,
                                                           float* __restrict__ partial) {
  __shared__ float sh[2][8];
  const int64_t beg = (int64_t)blockIdx.x * 8192, end = min(B, beg + 8192);
  float a = 0.f, c = 0.f;
  for (int64_t i = beg + threadIdx.x; i < end; i += 256) {
    a += logq[i] - logpz[i];
    c += -logpx[i];
  }
  a = warp_sum(a);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, sc = 0.f;
    for (int w = 0; w < 8; ++w) { sa += sh[0][w]; sc += sh[1][w]; }
    partial[2 * blockIdx.x] = sa;
    partial[2 * blockIdx.x + 1] = sc;
  }
}
__global__ void elbo_final_kernel(const float* __restrict__ partial, int nb, int64_t B, float w, float* scalars) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float sa = 0.f, sc = 0.f;
  for (int i = 0; i < nb; ++i) { sa += partial[2 * i]; sc += partial[2 * i + 1]; }
  const float kl = sa / (float)B, nll = sc / (float)B;
  scalars[0] = nll + w * kl;
  scalars[1] = nll;
  scalars[2] = kl;
}

// dst[b, c] (+)= a * src[b, c] over a column block
__global__ void axpy2d_kernel(const float* __restrict__ src, int64_t ld_s, float a, float* dst, int64_t ld_d, int64_t B,
                              int cols, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * cols) return;
  const int64_t b = i / cols;
  const int c = (int)(i - b * cols);
  const float v = a * src[b * ld_s + c];
  float* d = dst + b * ld_d + c;
  *d = accumulate ? *d + v : v;
}
__global__ void fill_kernel(float* p, float v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// decoder head: d/d params of g * sum_d log N(x_d; loc_d, softplus(raw_d)), params = [loc | raw]
__global__ void decoder_head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pd, int64_t B, int D,
                                        float g, float* __restrict__ g_pd) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float loc = pd[b * 2 * D + d], raw = pd[b * 2 * D + D + d];
  const float sc = softplus_tf(raw);
  const float u = x[i] / sc - loc / sc;
  g_pd[b * 2 * D + d] = g * (u / sc);
  g_pd[b * 2 * D + D + d] = g * ((u * u - 1.f) / sc) * sigmoidf_(raw);
}

// encoder head: log q(z|x) terms (explicit z path + parameter path) and the reparameterisation z = eps * s + loc.
// g_z holds d loss / d z from decoder + prior on entry.
__global__ void encoder_head_bwd_kernel(const float* __restrict__ z, const float* __restrict__ pe,
                                        const float* __restrict__ eps, const float* __restrict__ g_z, int64_t B, int D,
                                        float gq, float* __restrict__ g_pe) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int64_t b = i / D;
  const int d = (int)(i - b * D);
  const float loc = pe[b * 2 * D + d], raw = pe[b * 2 * D + D + d];
  const float sc = softplus_tf(raw);
  const float u = z[i] / sc - loc / sc;
  const float gz = g_z[i] + gq * (-u / sc);
  g_pe[b * 2 * D + d] = gq * (u / sc) + gz;
  g_pe[b * 2 * D + D + d] = (gq * ((u * u - 1.f) / sc) + gz * eps[i]) * sigmoidf_(raw);
}

}  // namespace vms

using namespace vms;

namespace {

vms_status alloc(vms_elbo_plan_s* pl, float** p, size_t n_floats) {
  void* v = nullptr;
  VMS_CUDA(cudaMalloc(&v, (n_floats ? n_floats : 4) * sizeof(float)));
  pl->allocs.push_back(v);
  *p = (float*)v;
  return VMS_OK;
}

vms_status dense_fwd(const float* x, int64_t ldx, const float* W, const float* b, int64_t B, int K, int N, int act,
                     float* out, int64_t ldo, cudaStream_t st) {
  return dense_forward_impl(x, ldx, W, b, B, K, N, act, nullptr, 0, nullptr, 0, out, ldo, st, false);
}

// weight + bias gradient partials of one Dense layer written straight into the flat partial-gradient stack:
// [W; b] is a contiguous [K+1, N] block at `off` (the Keras "kernel then bias" order).
vms_status dense_wgrad(vms_elbo_plan_s* pl, const float* x, int64_t ldx, int64_t B, int K, int N, int act,
                       const float* out, int64_t ldo, const float* g_out, int64_t ldg, int64_t off, int splits,
                       cudaStream_t st) {
  WgradParams p = {};
  p.B = B; p.Kin = K; p.N = N;
  p.x = x; p.ldx = ldx;
  p.g = g_out; p.ldg = ldg; p.out = out; p.ldo = ldo; p.act = act;
  p.part = pl->gpart + off; p.split_stride = pl->off.total; p.splits = splits;
  return gemm_wgrad(p, st);
}
vms_status dense_xgrad(const float* W, int64_t B, int K, int N, int act, const float* out, int64_t ldo,
                       const float* g_out, int64_t ldg, float* g_x, int64_t ldgx, int accumulate, cudaStream_t st) {
  RowTileParams p = {};
  p.M = (int)B; p.N = K; p.K = N;
  p.A = g_out; p.lda = ldg; p.Ao = out; p.ldao = ldo; p.a_act = act;
  p.Bm = W; p.ldb = N; p.tb = 1;
  p.C = g_x; p.ldc = ldgx; p.accumulate = accumulate;
  return gemm_rowtile(p, st);
}

#define VMS_TRY(expr)          \
  do {                         \
    vms_status _s = (expr);    \
    if (_s) return _s;         \
  } while (0)

inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// mode 2 (or auto at large batches): the flow blocks run as fused tensor-core kernels
bool plan_uses_tc(const vms_elbo_plan_s* pl, int64_t B) {
  if (!pl->tc_ok || B < 64) return false;
  // auto: above one wave of the fused kernel's tiles; or, when the single fused kernel does not support the shape,
  // from the batch where the tensor-core plan's ~0.23 ms floor beats the per-layer FFMA plan
  return pl->mode == 2 || (pl->mode == 0 && (B >= pl->tc_auto_batch || (!pl->fused && B >= 1024)));
}

FlowTcArgs flow_tc_args(const vms_elbo_plan_s* pl, const float* theta, int i, int64_t B) {
  const vms_elbo_desc& d = pl->d;
  const FlowBlock& fb = pl->blocks[i];
  FlowTcArgs ta = {};
  ta.B = B; ta.dz = d.dz; ta.cs0 = fb.cs0; ta.nc = fb.cs1 - fb.cs0; ta.ts0 = fb.ts0;
  ta.H = d.flow_hidden; ta.K = d.num_bins; ta.bin_min = d.bin_min; ta.bin_max = d.bin_max;
  ta.d1W = theta + fb.off_d1W; ta.d1b = theta + fb.off_d1b; ta.hW = theta + fb.off_hW; ta.hb = theta + fb.off_hb;
  ta.err = pl->tc_err;
  return ta;
}

// CTAs (= weight-gradient partials) of the large-batch MLP backward kernels: two per SM, 128-row tiles
int mlp_grid(int64_t B) {
  const int64_t tiles = (B + 127) / 128;
  return (int)(tiles < 2 * sm_count() ? tiles : 2 * sm_count());
}

MlpArgs mlp_args(const vms_elbo_plan_s* pl, const float* theta, bool encoder, int64_t B) {
  const vms_elbo_desc& d = pl->d;
  const Offsets& o = pl->off;
  MlpArgs ma = {};
  ma.B = B; ma.H = d.hidden;
  ma.Din = encoder ? d.dx : d.dz;
  ma.Dout = encoder ? 2 * d.dz : 2 * d.dx;
  ma.o_W0 = encoder ? o.enc0W : o.dec0W; ma.o_b0 = encoder ? o.enc0b : o.dec0b;
  ma.o_W1 = encoder ? o.enc1W : o.dec1W; ma.o_b1 = encoder ? o.enc1b : o.dec1b;
  ma.W0 = theta + ma.o_W0; ma.b0 = theta + ma.o_b0; ma.W1 = theta + ma.o_W1; ma.b1 = theta + ma.o_b1;
  ma.part = pl->tc_part; ma.part_stride = pl->off.total;
  return ma;
}

vms_status forward_body(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B,
                        float* scalars, cudaStream_t st) {
  const vms_elbo_desc& d = pl->d;
  const Offsets& o = pl->off;
  const int nb = d.num_blocks;
  // encoder: FCDeepNN dx -> H (relu) -> 2 dz ; IndependentNormal: loc | softplus(raw)
  const bool tc = plan_uses_tc(pl, B);
  if (tc) {
    // large-batch plan: both Dense layers in one kernel, the hidden layer never reaches HBM (mlp_stream.cu)
    MlpArgs ma = mlp_args(pl, theta, true, B);
    ma.in = x; ma.ld_in = d.dx; ma.out = pl->pe; ma.ld_out = 2 * d.dz;
    VMS_TRY(mlp_stream_forward(ma, 2 * sm_count(), st));
  } else {
    VMS_TRY(dense_fwd(x, d.dx, theta + o.enc0W, theta + o.enc0b, B, d.dx, d.hidden, VMS_ACT_RELU, pl->he, d.hidden, st));
    VMS_TRY(dense_fwd(pl->he, d.hidden, theta + o.enc1W, theta + o.enc1b, B, d.hidden, 2 * d.dz, VMS_ACT_NONE, pl->pe,
                      2 * d.dz, st));
  }
  float* zbuf = pl->u[nb];
  VMS_TRY(vms_normal_sample_log_prob(pl->pe, 2 * d.dz, 0, d.dz, VMS_SCALE_SOFTPLUS, eps, B, d.dz, zbuf, d.dz, pl->logq,
                                     (vms_stream)st));
  // prior log p(z): chain inverse, block n-1 first (flows.py:323 reverses the list for tfp Chain)
  for (int i = nb - 1; i >= 0; --i) {
    const FlowBlock& fb = pl->blocks[i];
    const float* uin = pl->u[i + 1];
    float* uout = pl->u[i];
    if (plan_uses_tc(pl, B)) {
      // conditioner + spline of the block in ONE tensor-core kernel (flow_tc.cu): hid / raw never reach HBM
      FlowTcArgs ta = flow_tc_args(pl, theta, i, B);
      ta.uin = uin; ta.uout = uout; ta.logpz = pl->logpz; ta.accumulate = (i != nb - 1);
      VMS_TRY(flow_tc_forward(ta, st));
      continue;
    }
    const float* cond = fb.cs1 > fb.cs0 ? uin + fb.cs0 : nullptr;  // NULL => ones((B,1)) (flows.py:184-185)
    VMS_TRY(dense_fwd(cond, d.dz, theta + fb.off_d1W, theta + fb.off_d1b, B, fb.cin, d.flow_hidden, VMS_ACT_TANH,
                      pl->hid[i], d.flow_hidden, st));
    VMS_TRY(dense_fwd(pl->hid[i], d.flow_hidden, theta + fb.off_hW, theta + fb.off_hb, B, d.flow_hidden, fb.ldr,
                      VMS_ACT_NONE, pl->raw[i], fb.ldr, st));
    if (fb.cs1 > fb.cs0)
      VMS_CUDA(cudaMemcpy2DAsync(uout + fb.cs0, d.dz * sizeof(float), uin + fb.cs0, d.dz * sizeof(float),
                                 (fb.cs1 - fb.cs0) * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    vms_rqs_args a = {};
    a.n_rows = B; a.n_dims = fb.dt; a.num_bins = d.num_bins; a.bin_min = d.bin_min; a.bin_max = d.bin_max;
    a.v_in = uin + fb.ts0; a.ld_in = d.dz;
    a.raw_w = pl->raw[i]; a.ld_w = fb.ldr;
    a.raw_h = pl->raw[i] + fb.dt * d.num_bins; a.ld_h = fb.ldr;
    a.raw_s = pl->raw[i] + 2 * fb.dt * d.num_bins; a.ld_s = fb.ldr;
    a.v_out = uout + fb.ts0; a.ld_out = d.dz;
    a.ldj = nullptr; a.ldj_sum = pl->logpz; a.accumulate = (i != nb - 1); a.inverse_dir = 1;
    VMS_TRY(vms_rqs_apply(&a, (vms_stream)st));
  }
  VMS_TRY(vms_std_normal_log_prob(pl->u[0], d.dz, B, d.dz, pl->logpz, nb > 0, (vms_stream)st));
  // decoder
  if (tc) {
    MlpArgs ma = mlp_args(pl, theta, false, B);
    ma.in = zbuf; ma.ld_in = d.dz; ma.out = pl->pd; ma.ld_out = 2 * d.dx;
    VMS_TRY(mlp_stream_forward(ma, 2 * sm_count(), st));
  } else {
    VMS_TRY(dense_fwd(zbuf, d.dz, theta + o.dec0W, theta + o.dec0b, B, d.dz, d.hidden, VMS_ACT_RELU, pl->hd, d.hidden, st));
    VMS_TRY(dense_fwd(pl->hd, d.hidden, theta + o.dec1W, theta + o.dec1b, B, d.hidden, 2 * d.dx, VMS_ACT_NONE, pl->pd,
                      2 * d.dx, st));
  }
  {
    int32_t kind[64], loc[64], sc[64];
    for (int j = 0; j < d.dx; ++j) { kind[j] = VMS_DIST_NORMAL; loc[j] = j; sc[j] = d.dx + j; }
    VMS_TRY(vms_blockwise_log_prob(x, d.dx, pl->pd, 2 * d.dx, B, d.dx, kind, loc, nullptr, sc, VMS_SCALE_SOFTPLUS,
                                   pl->logpx, 0, (vms_stream)st));
  }
  const int nbk = (int)((B + 8191) / 8192);
  elbo_partial_kernel<<<nbk, 256, 0, st>>>(pl->logq, pl->logpz, pl->logpx, B, pl->partial);
  VMS_LAUNCH_CHECK("elbo_partial_kernel");
  elbo_final_kernel<<<1, 32, 0, st>>>(pl->partial, nbk, B, d.kl_weight, scalars ? scalars : pl->scalars);
  VMS_LAUNCH_CHECK("elbo_final_kernel");
  return VMS_OK;
}

vms_status backward_body(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B,
                         float* grad, cudaStream_t st) {
  const vms_elbo_desc& d = pl->d;
  const Offsets& o = pl->off;
  const int nb = d.num_blocks;
  const int splits = dense_splits(B);
  const float invB = 1.0f / (float)B;
  const float g_logpx = -invB, g_logq = d.kl_weight * invB, g_logpz = -d.kl_weight * invB;
  const float* z = pl->u[nb];
  // decoder head and MLP
  decoder_head_bwd_kernel<<<nblk(B * d.dx, 256), 256, 0, st>>>(x, pl->pd, B, d.dx, g_logpx, pl->g_pd);
  VMS_LAUNCH_CHECK("decoder_head_bwd_kernel");
  const bool tc = plan_uses_tc(pl, B);
  if (tc) {
    MlpArgs ma = mlp_args(pl, theta, false, B);
    ma.in = z; ma.ld_in = d.dz; ma.g_out = pl->g_pd; ma.ld_g = 2 * d.dx; ma.g_in = pl->g_z; ma.ld_gin = d.dz;
    VMS_TRY(mlp_stream_backward(ma, mlp_grid(B), st));
  } else {
  VMS_TRY(dense_wgrad(pl, pl->hd, d.hidden, B, d.hidden, 2 * d.dx, VMS_ACT_NONE, nullptr, 0, pl->g_pd, 2 * d.dx,
                      o.dec1W, splits, st));
  VMS_TRY(dense_xgrad(theta + o.dec1W, B, d.hidden, 2 * d.dx, VMS_ACT_NONE, nullptr, 0, pl->g_pd, 2 * d.dx, pl->g_hd,
                      d.hidden, 0, st));
  VMS_TRY(dense_wgrad(pl, z, d.dz, B, d.dz, d.hidden, VMS_ACT_RELU, pl->hd, d.hidden, pl->g_hd, d.hidden, o.dec0W,
                      splits, st));
  VMS_TRY(dense_xgrad(theta + o.dec0W, B, d.dz, d.hidden, VMS_ACT_RELU, pl->hd, d.hidden, pl->g_hd, d.hidden, pl->g_z,
                      d.dz, 0, st));
  }
  // prior
  if (nb == 0) {
    // d/dz [ g_logpz * sum(-0.5 z^2) ] = -g_logpz * z
    axpy2d_kernel<<<nblk(B * d.dz, 256), 256, 0, st>>>(z, d.dz, -g_logpz, pl->g_z, d.dz, B, d.dz, 1);
    VMS_LAUNCH_CHECK("axpy2d_kernel");
  } else {
    float* g_cur = pl->g_ua;
    float* g_nxt = pl->g_ub;
    axpy2d_kernel<<<nblk(B * d.dz, 256), 256, 0, st>>>(pl->u[0], d.dz, -g_logpz, g_cur, d.dz, B, d.dz, 0);
    VMS_LAUNCH_CHECK("axpy2d_kernel");
    fill_kernel<<<nblk(B, 256), 256, 0, st>>>(pl->g_ldj, g_logpz, B);
    VMS_LAUNCH_CHECK("fill_kernel");
    for (int i = 0; i < nb; ++i) {
      const FlowBlock& fb = pl->blocks[i];
      const float* uin = pl->u[i + 1];
      const int nc = fb.cs1 - fb.cs0;
      if (plan_uses_tc(pl, B)) {
        FlowTcArgs ta = flow_tc_args(pl, theta, i, B);
        ta.uin = uin;
        ta.g_cur = g_cur; ta.g_nxt = g_nxt; ta.g_ldj = g_logpz;
        ta.part = pl->tc_part; ta.part_stride = pl->off.total;
        ta.o_d1W = fb.off_d1W; ta.o_d1b = fb.off_d1b; ta.o_hW = fb.off_hW; ta.o_hb = fb.off_hb;
        VMS_TRY(flow_tc_backward(ta, st));
        float* t = g_cur; g_cur = g_nxt; g_nxt = t;
        continue;
      }
      if (nc > 0)
        VMS_CUDA(cudaMemcpy2DAsync(g_nxt + fb.cs0, d.dz * sizeof(float), g_cur + fb.cs0, d.dz * sizeof(float),
                                   nc * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
      vms_rqs_bwd_args a = {};
      a.fwd.n_rows = B; a.fwd.n_dims = fb.dt; a.fwd.num_bins = d.num_bins; a.fwd.bin_min = d.bin_min;
      a.fwd.bin_max = d.bin_max;
      a.fwd.v_in = uin + fb.ts0; a.fwd.ld_in = d.dz;
      a.fwd.raw_w = pl->raw[i]; a.fwd.ld_w = fb.ldr;
      a.fwd.raw_h = pl->raw[i] + fb.dt * d.num_bins; a.fwd.ld_h = fb.ldr;
      a.fwd.raw_s = pl->raw[i] + 2 * fb.dt * d.num_bins; a.fwd.ld_s = fb.ldr;
      a.fwd.inverse_dir = 1;
      a.g_out = g_cur + fb.ts0; a.ld_g_out = d.dz;
      a.g_ldj_sum = pl->g_ldj;
      a.g_in = g_nxt + fb.ts0; a.ld_g_in = d.dz;
      a.g_raw_w = pl->g_raw; a.ld_gw = fb.ldr;
      a.g_raw_h = pl->g_raw + fb.dt * d.num_bins; a.ld_gh = fb.ldr;
      a.g_raw_s = pl->g_raw + 2 * fb.dt * d.num_bins; a.ld_gs = fb.ldr;
      VMS_TRY(vms_rqs_apply_backward(&a, (vms_stream)st));
      // heads: raw = hid @ hW + hb
      VMS_TRY(dense_wgrad(pl, pl->hid[i], d.flow_hidden, B, d.flow_hidden, fb.ldr, VMS_ACT_NONE, nullptr, 0, pl->g_raw,
                          fb.ldr, fb.off_hW, splits, st));
      VMS_TRY(dense_xgrad(theta + fb.off_hW, B, d.flow_hidden, fb.ldr, VMS_ACT_NONE, nullptr, 0, pl->g_raw, fb.ldr,
                          pl->g_hid, d.flow_hidden, 0, st));
      // d1: hid = tanh(cond @ d1W + d1b)
      const float* cond = nc > 0 ? uin + fb.cs0 : nullptr;
      VMS_TRY(dense_wgrad(pl, cond, d.dz, B, fb.cin, d.flow_hidden, VMS_ACT_TANH, pl->hid[i], d.flow_hidden, pl->g_hid,
                          d.flow_hidden, fb.off_d1W, splits, st));
      if (nc > 0)
        VMS_TRY(dense_xgrad(theta + fb.off_d1W, B, fb.cin, d.flow_hidden, VMS_ACT_TANH, pl->hid[i], d.flow_hidden,
                            pl->g_hid, d.flow_hidden, g_nxt + fb.cs0, d.dz, 1, st));
      float* t = g_cur; g_cur = g_nxt; g_nxt = t;
    }
    axpy2d_kernel<<<nblk(B * d.dz, 256), 256, 0, st>>>(g_cur, d.dz, 1.f, pl->g_z, d.dz, B, d.dz, 1);
    VMS_LAUNCH_CHECK("axpy2d_kernel");
  }
  // encoder head + MLP
  encoder_head_bwd_kernel<<<nblk(B * d.dz, 256), 256, 0, st>>>(z, pl->pe, eps, pl->g_z, B, d.dz, g_logq, pl->g_pe);
  VMS_LAUNCH_CHECK("encoder_head_bwd_kernel");
  if (tc) {
    MlpArgs ma = mlp_args(pl, theta, true, B);
    ma.in = x; ma.ld_in = d.dx; ma.g_out = pl->g_pe; ma.ld_g = 2 * d.dz;
    VMS_TRY(mlp_stream_backward(ma, mlp_grid(B), st));
    // every layer's partials are per CTA of the large-batch kernels: fixed-order sums over the flat gradient (the MLP
    // kernels run two CTAs per SM, the coupling-block kernels one)
    const int64_t flow0 = nb > 0 ? pl->blocks[0].off_d1W : o.total;
    VMS_TRY(sum_partials_launch(pl->tc_part, mlp_grid(B), o.total, flow0, grad, 0, nullptr, 1.f, 0, st));
    if (nb > 0)
      VMS_TRY(sum_partials_launch(pl->tc_part + flow0, flow_tc_grid(B), o.total, o.total - flow0, grad + flow0, 0, nullptr,
                                  1.f, 0, st));
    return VMS_OK;
  }
  VMS_TRY(dense_wgrad(pl, pl->he, d.hidden, B, d.hidden, 2 * d.dz, VMS_ACT_NONE, nullptr, 0, pl->g_pe, 2 * d.dz,
                      o.enc1W, splits, st));
  VMS_TRY(dense_xgrad(theta + o.enc1W, B, d.hidden, 2 * d.dz, VMS_ACT_NONE, nullptr, 0, pl->g_pe, 2 * d.dz, pl->g_he,
                      d.hidden, 0, st));
  VMS_TRY(dense_wgrad(pl, x, d.dx, B, d.dx, d.hidden, VMS_ACT_RELU, pl->he, d.hidden, pl->g_he, d.hidden, o.enc0W,
                      splits, st));
  // flat gradient = fixed-order sum of the split partials
  VMS_TRY(sum_partials_launch(pl->gpart, splits, o.total, o.total, grad, 0, nullptr, 1.f, 0, st));
  return VMS_OK;
}

// Run `body` through a cached CUDA graph when the stream allows capture (non-default stream).
template <typename F>
vms_status run_graphed(vms_elbo_plan_s* pl, const std::array<uintptr_t, 10>& key, cudaStream_t st, F body) {
  if (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) return body();
  auto it = pl->graphs.find(key);
  if (it == pl->graphs.end()) {
    const unsigned long long before = vms_launch_count();
    cudaGraph_t g = nullptr;
    VMS_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    vms_status s = body();
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (s || e != cudaSuccess) cudaGetLastError();  // an invalidated capture must not poison later launch checks
    if (s) { if (g) cudaGraphDestroy(g); return s; }
    if (e != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(e)); return VMS_ERR_CUDA; }
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { set_error("graph instantiate failed: %s", cudaGetErrorString(e)); return VMS_ERR_CUDA; }
    if (pl->graphs.size() >= 16) {  // bounded cache
      for (auto& kv : pl->graphs) cudaGraphExecDestroy(kv.second.first);
      pl->graphs.clear();
    }
    // kernels recorded during capture did not run: un-count them, and count them per replay instead
    const int n_kernels = (int)(vms_launch_count() - before);
    count_launch(-n_kernels);
    it = pl->graphs.emplace(key, std::make_pair(ex, n_kernels)).first;
  }
  VMS_CUDA(cudaGraphLaunch(it->second.first, st));
  count_launch(it->second.second);
  return VMS_OK;
}

}  // namespace

// the single fused kernel runs the step unless a mode forces the per-layer plan or the batch is large enough for the
// tensor-core flow kernels (auto)
static bool use_fused(const vms_elbo_plan_s* pl, int64_t B) {
  if (pl->fused && pl->mode == 4) return true;
  return pl->fused && pl->mode == 0 && !(pl->tc_ok && B >= pl->tc_auto_batch);
}

// forward + backward / training steps: the whole-step tensor-core kernel (elbo_tcf.cu) when forced (mode 3) or, in auto
// mode, whenever the batch fits one wave of its 32-row tiles (VMS_TCF_AUTO=0 keeps auto mode on the FFMA fused kernel)
static bool use_tcf(const vms_elbo_plan_s* pl, int64_t B) {
  if (!tcf_available(pl, B)) return false;
  if (pl->mode == 3) return true;
  static int auto_on = -1;
  if (auto_on < 0) {
    const char* e = getenv("VMS_TCF_AUTO");
    auto_on = (e && e[0] == '0') ? 0 : 1;
  }
  // the kernel stays ahead of the per-block tensor-core plan for up to three waves of its one-tile CTAs (tcf_create sizes
  // max_tiles accordingly); an explicit VMS_TC_AUTO_BATCH / vms_elbo_plan_set_tc_auto_batch still wins
  return pl->mode == 0 && auto_on && !(pl->tc_ok && pl->tc_auto_user && B >= pl->tc_auto_batch);
}

extern "C" {

int64_t vms_elbo_param_count(const vms_elbo_desc* desc) {
  if (!desc) return -1;
  return layout(*desc, nullptr).total;
}

vms_status vms_elbo_plan_create(const vms_elbo_desc* desc, vms_elbo_plan* plan) {
  VMS_REQUIRE(desc && plan, VMS_ERR_INVALID_ARG, "elbo_plan_create: NULL argument");
  const vms_elbo_desc& d = *desc;
  VMS_REQUIRE(d.dx >= 1 && d.dx <= 64 && d.dz >= 1 && d.dz <= 64 && d.hidden >= 1, VMS_ERR_SHAPE,
              "elbo_plan_create: dx, dz must be in [1, 64] and hidden >= 1");
  VMS_REQUIRE(d.num_blocks >= 0 && d.max_batch >= 1, VMS_ERR_INVALID_ARG, "elbo_plan_create: bad num_blocks / max_batch");
  VMS_REQUIRE(d.num_blocks == 0 || (d.num_bins >= 2 && d.num_bins <= 64 && d.flow_hidden >= 1 && d.bin_max > d.bin_min),
              VMS_ERR_INVALID_ARG, "elbo_plan_create: bad flow parameters");
  {
    vms_status ps = rqs_prepare();
    if (ps) return ps;
  }
  vms_elbo_plan_s* pl = new vms_elbo_plan_s();
  pl->d = d;
  pl->off = layout(d, &pl->blocks);
  pl->maxB = d.max_batch;
  pl->splits_max = dense_splits(d.max_batch);
  const size_t B = (size_t)d.max_batch;
  int max_ldr = 1;
  for (auto& b : pl->blocks) max_ldr = b.ldr > max_ldr ? b.ldr : max_ldr;
  vms_status s = VMS_OK;
#define A_(ptr, n) if (!s) s = alloc(pl, &(ptr), (n))
  A_(pl->he, B * d.hidden); A_(pl->pe, B * 2 * d.dz); A_(pl->logq, B); A_(pl->logpz, B); A_(pl->logpx, B);
  A_(pl->hd, B * d.hidden); A_(pl->pd, B * 2 * d.dx); A_(pl->scalars, 4); A_(pl->partial, 2 * ((B + 8191) / 8192) + 2);
  pl->u.resize(d.num_blocks + 1); pl->hid.resize(d.num_blocks); pl->raw.resize(d.num_blocks);
  for (int i = 0; i <= d.num_blocks; ++i) A_(pl->u[i], B * d.dz);
  for (int i = 0; i < d.num_blocks; ++i) { A_(pl->hid[i], B * d.flow_hidden); A_(pl->raw[i], B * pl->blocks[i].ldr); }
  A_(pl->g_pd, B * 2 * d.dx); A_(pl->g_hd, B * d.hidden); A_(pl->g_z, B * d.dz); A_(pl->g_pe, B * 2 * d.dz);
  A_(pl->g_he, B * d.hidden); A_(pl->g_ua, B * d.dz); A_(pl->g_ub, B * d.dz);
  A_(pl->g_raw, B * max_ldr); A_(pl->g_hid, B * (d.num_blocks ? d.flow_hidden : 1)); A_(pl->g_ldj, B);
  A_(pl->gpart, (size_t)pl->splits_max * pl->off.total);
  pl->tc_ok = d.num_blocks > 0;
  for (auto& b : pl->blocks) pl->tc_ok = pl->tc_ok && flow_tc_supported(d.dz, b.cin, b.dt, d.flow_hidden, d.num_bins);
  pl->tc_ok = pl->tc_ok && mlp_stream_supported(d.dx, d.hidden, 2 * d.dz) && mlp_stream_supported(d.dz, d.hidden, 2 * d.dx);
  // measured crossover (B200): the single fused kernel wins while its 32-row tiles fit one wave (B <= 32 x #SMs = 4736:
  // 0.149 ms at 4096 against 0.238 ms); from the second wave on the tensor-core plan is faster (0.242 vs 0.313 ms at 6144)
  pl->tc_auto_batch = 32 * (int64_t)sm_count() + 1;
  if (const char* e = getenv("VMS_TC_AUTO_BATCH")) {  // auto mode switches to the tensor-core plan from this batch on
    const long long v = atoll(e);
    if (v >= 64) { pl->tc_auto_batch = v; pl->tc_auto_user = true; }
  }
  if (pl->tc_ok) {
    A_(pl->tc_part, (size_t)2 * sm_count() * pl->off.total);
    float* e = nullptr;
    A_(e, 4);
    pl->tc_err = reinterpret_cast<int*>(e);
    if (!s && cudaMemset(pl->tc_err, 0, 16) != cudaSuccess) s = VMS_ERR_CUDA;
  }
#undef A_
  if (!s) s = fused_create(pl);
  if (!s) s = tcf_create(pl);
  if (s) { vms_elbo_plan_destroy(pl); return s; }
  *plan = pl;
  return VMS_OK;
}

vms_status vms_elbo_plan_destroy(vms_elbo_plan pl) {
  if (!pl) return VMS_OK;
  for (auto& kv : pl->graphs) cudaGraphExecDestroy(kv.second.first);
  for (void* p : pl->allocs) cudaFree(p);
  fused_destroy(pl);
  tcf_destroy(pl);
  delete pl;
  return VMS_OK;
}

static vms_status check_call(vms_elbo_plan pl, const float* theta, const float* x, const float* eps, int64_t B) {
  VMS_REQUIRE(pl, VMS_ERR_INVALID_ARG, "elbo: NULL plan");
  VMS_REQUIRE(theta && x && eps, VMS_ERR_INVALID_ARG, "elbo: NULL tensor pointer");
  VMS_REQUIRE(B >= 1 && B <= pl->maxB, VMS_ERR_SHAPE, "elbo: batch %lld outside [1, max_batch=%lld]", (long long)B,
              (long long)pl->maxB);
  return VMS_OK;
}

vms_status vms_elbo_plan_set_mode(vms_elbo_plan pl, int mode) {
  VMS_REQUIRE(pl, VMS_ERR_INVALID_ARG, "elbo_plan_set_mode: NULL plan");
  VMS_REQUIRE(mode >= 0 && mode <= 4, VMS_ERR_INVALID_ARG,
              "elbo_plan_set_mode: mode must be 0 (auto), 1 (unfused, FFMA), 2 (unfused, tensor-core flow blocks), 3 "
              "(whole-step tensor-core kernel) or 4 (single FFMA fused kernel)");
  VMS_REQUIRE(mode != 4 || pl->fused, VMS_ERR_UNSUPPORTED, "elbo_plan_set_mode: the FFMA fused kernel does not support this shape");
  VMS_REQUIRE(mode != 3 || pl->tcf, VMS_ERR_UNSUPPORTED, "elbo_plan_set_mode: mode 3 does not support this shape");
  VMS_REQUIRE(mode != 2 || pl->tc_ok, VMS_ERR_UNSUPPORTED, "elbo_plan_set_mode: the tensor-core flow kernels do not support this shape");
  pl->mode = mode;
  return VMS_OK;
}

vms_status vms_elbo_plan_tc_status(vms_elbo_plan pl, int* err) {
  VMS_REQUIRE(pl && err, VMS_ERR_INVALID_ARG, "elbo_plan_tc_status: NULL argument");
  *err = 0;
  if (!pl->tc_err) return VMS_OK;
  VMS_CUDA(cudaDeviceSynchronize());
  VMS_CUDA(cudaMemcpy(err, pl->tc_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (*err) VMS_CUDA(cudaMemset(pl->tc_err, 0, sizeof(int)));
  return VMS_OK;
}

int vms_elbo_plan_path(vms_elbo_plan pl, int64_t B) {
  if (!pl) return -1;
  if (use_tcf(pl, B)) return 3;
  if (use_fused(pl, B)) return 0;
  return plan_uses_tc(pl, B) ? 2 : 1;
}

vms_status vms_elbo_plan_set_tc_auto_batch(vms_elbo_plan pl, int64_t batch) {
  VMS_REQUIRE(pl && batch >= 64, VMS_ERR_INVALID_ARG, "elbo_plan_set_tc_auto_batch: NULL plan or batch < 64");
  pl->tc_auto_batch = batch;
  pl->tc_auto_user = true;
  return VMS_OK;
}

vms_status vms_elbo_plan_set_timing(vms_elbo_plan pl, int max_launches) {
  return vms_elbo_plan_set_timing_every(pl, max_launches, 1);
}

vms_status vms_elbo_plan_set_timing_every(vms_elbo_plan pl, int max_launches, int every) {
  VMS_REQUIRE(pl, VMS_ERR_INVALID_ARG, "elbo_plan_set_timing: NULL plan");
  pl->timing_every = every > 1 ? every : 1;
  pl->timing_calls = 0;
  vms_status s = fused_set_timing(pl, max_launches);
  return s ? s : tcf_set_timing(pl, max_launches);
}

vms_status vms_elbo_plan_kernel_ms(vms_elbo_plan pl, double* total_ms, int* launches) {
  VMS_REQUIRE(pl && total_ms && launches, VMS_ERR_INVALID_ARG, "elbo_plan_kernel_ms: NULL argument");
  vms_status s = fused_kernel_ms(pl, total_ms, launches);  // (sets both outputs)
  return s ? s : tcf_kernel_ms(pl, total_ms, launches);     // (adds to them)
}

vms_status vms_elbo_plan_invalidate(vms_elbo_plan pl) {
  VMS_REQUIRE(pl, VMS_ERR_INVALID_ARG, "elbo_plan_invalidate: NULL plan");
  tcf_invalidate(pl);
  return VMS_OK;
}

int vms_elbo_plan_is_fused(vms_elbo_plan pl) { return pl && pl->fused && (pl->mode == 0 || pl->mode == 4) ? 1 : 0; }

vms_status vms_elbo_forward(vms_elbo_plan pl, const float* theta, const float* x, const float* eps, int64_t B, float* z,
                            float* logq, float* logpz, float* logpx, float* scalars, vms_stream stream) {
  VMS_RANGE("vms_elbo_forward");
  vms_status s = check_call(pl, theta, x, eps, B);
  if (s) return s;
  cudaStream_t st = as_stream(stream);
  if (use_fused(pl, B))
    return fused_run(pl, theta, x, eps, B, false, z, logq, logpz, logpx, nullptr, scalars, st);
  std::array<uintptr_t, 10> key = {(uintptr_t)(0 | (pl->mode << 8)), (uintptr_t)B, (uintptr_t)theta, (uintptr_t)x, (uintptr_t)eps, (uintptr_t)z,
                                   (uintptr_t)logq, (uintptr_t)logpz, (uintptr_t)logpx, (uintptr_t)scalars};
  const vms_elbo_desc& d = pl->d;
  return run_graphed(pl, key, st, [&]() -> vms_status {
    VMS_TRY(forward_body(pl, theta, x, eps, B, scalars, st));
    if (z) VMS_CUDA(cudaMemcpyAsync(z, pl->u[d.num_blocks], B * d.dz * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (logq) VMS_CUDA(cudaMemcpyAsync(logq, pl->logq, B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (logpz) VMS_CUDA(cudaMemcpyAsync(logpz, pl->logpz, B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (logpx) VMS_CUDA(cudaMemcpyAsync(logpx, pl->logpx, B * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return VMS_OK;
  });
}

vms_status vms_elbo_forward_backward(vms_elbo_plan pl, const float* theta, const float* x, const float* eps, int64_t B,
                                     float* grad, float* scalars, vms_stream stream) {
  VMS_RANGE("vms_elbo_forward_backward");
  vms_status s = check_call(pl, theta, x, eps, B);
  if (s) return s;
  VMS_REQUIRE(grad, VMS_ERR_INVALID_ARG, "elbo_forward_backward: NULL grad");
  cudaStream_t st = as_stream(stream);
  if (use_tcf(pl, B)) return tcf_run(pl, theta, x, eps, B, grad, scalars, st);
  if (use_fused(pl, B))
    return fused_run(pl, theta, x, eps, B, true, nullptr, nullptr, nullptr, nullptr, grad, scalars, st);
  std::array<uintptr_t, 10> key = {(uintptr_t)(1 | (pl->mode << 8)), (uintptr_t)B, (uintptr_t)theta, (uintptr_t)x, (uintptr_t)eps, (uintptr_t)grad,
                                   (uintptr_t)scalars, 0, 0, 0};
  return run_graphed(pl, key, st, [&]() -> vms_status {
    VMS_TRY(forward_body(pl, theta, x, eps, B, scalars, st));
    return backward_body(pl, theta, x, eps, B, grad, st);
  });
}

/* One DATA-PARALLEL training step: forward + backward, the gradient exchange over NVLink peer memory and Keras Adam.  When the
 * whole-step tensor-core kernel serves the batch, its finish kernel does the exchange itself (two launches per step); every
 * other plan writes its gradient into this rank's slot and launches vms_peer_allreduce_adam.  step / peer_bases as there. */
vms_status vms_elbo_train_step_peer(vms_elbo_plan pl, float* theta, const float* x, const float* eps, int64_t B, float* scalars,
                                    float* m, float* v, int64_t t, double lr, double beta1, double beta2, double eps_adam,
                                    int world, int rank, void* const* peer_bases, unsigned long long step, vms_stream stream) {
  VMS_RANGE("vms_elbo_train_step_peer");
  vms_status s = check_call(pl, theta, x, eps, B);
  if (s) return s;
  VMS_REQUIRE(m && v && t >= 1, VMS_ERR_INVALID_ARG, "elbo_train_step_peer: NULL m / v or t < 1");
  const int64_t P = pl->off.total;
  if (use_tcf(pl, B) && tcf_peer_ok(pl)) {
    PeerArgs a = {};
    if ((s = peer_fill_args(a, world, rank, peer_bases, P, step, 1.0f / (float)world))) return s;
    a.theta = theta; a.m = m; a.v = v;
    a.lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
    a.one_minus_b1 = (float)(1.0 - beta1);
    a.one_minus_b2 = (float)(1.0 - beta2);
    a.eps = (float)eps_adam;
    return tcf_run(pl, theta, x, eps, B, nullptr, scalars, as_stream(stream), nullptr, &a);
  }
  VMS_REQUIRE(peer_bases && rank >= 0 && rank < world && peer_bases[rank], VMS_ERR_INVALID_ARG, "elbo_train_step_peer: bad peers");
  float* slot = (float*)peer_bases[rank] + (int64_t)(step & 1ull) * P;
  if ((s = vms_elbo_forward_backward(pl, theta, x, eps, B, slot, scalars, stream))) return s;
  return vms_peer_allreduce_adam(world, rank, peer_bases, P, step, 1.0f / (float)world, theta, m, v, t, lr, beta1, beta2, eps_adam,
                                 nullptr, stream);
}

/* One training step: ELBO forward + backward + Keras Adam (tests/test_models.py:181) on the flat buffers.  On the fused
 * path the optimiser update rides in the kernel that sums the partial gradients (2 launches per step in all); the
 * unfused path runs vms_elbo_forward_backward followed by vms_adam_step. */
vms_status vms_elbo_train_step(vms_elbo_plan pl, float* theta, const float* x, const float* eps, int64_t B, float* grad,
                               float* scalars, float* m, float* v, int64_t t, double lr, double beta1, double beta2,
                               double eps_adam, vms_stream stream) {
  VMS_RANGE("vms_elbo_train_step");
  vms_status s = check_call(pl, theta, x, eps, B);
  if (s) return s;
  VMS_REQUIRE(grad && m && v && t >= 1, VMS_ERR_INVALID_ARG, "elbo_train_step: NULL grad / m / v or t < 1");
  if (use_tcf(pl, B)) {
    FusedAdam ad;
    ad.theta = theta; ad.m = m; ad.v = v;
    ad.lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
    ad.one_minus_b1 = (float)(1.0 - beta1);
    ad.one_minus_b2 = (float)(1.0 - beta2);
    ad.eps = (float)eps_adam;
    return tcf_run(pl, theta, x, eps, B, grad, scalars, as_stream(stream), &ad);
  }
  if (use_fused(pl, B)) {
    FusedAdam ad;
    ad.theta = theta; ad.m = m; ad.v = v;
    ad.lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
    ad.one_minus_b1 = (float)(1.0 - beta1);
    ad.one_minus_b2 = (float)(1.0 - beta2);
    ad.eps = (float)eps_adam;
    return fused_run(pl, theta, x, eps, B, true, nullptr, nullptr, nullptr, nullptr, grad, scalars, as_stream(stream), &ad);
  }
  s = vms_elbo_forward_backward(pl, theta, x, eps, B, grad, scalars, stream);
  if (s) return s;
  return vms_adam_step(theta, grad, 1, 1.0f, m, v, pl->off.total, t, lr, beta1, beta2, eps_adam, stream);
}

}  // extern "C"
