// elbo_plan.cuh -- plan object shared by the two implementations of the whole-VAE ELBO step:
//   elbo.cu        unfused path: ~60 small kernels replayed as one CUDA graph (any shape)
//   elbo_fused.cu  fused path: ONE persistent kernel, 32-row tiles resident in shared memory (shapes that fit)
#pragma once
#include "dense.cuh"
#include <vector>
#include <map>
#include <array>

namespace vms {

struct FlowBlock {
  int cs0, cs1, ts0, ts1;  // conditioner / transformed column ranges (flows.py:290-306 + tfp RealNVP reverse mask)
  int cin, dt, ldr;        // conditioner input width (>=1: ones input when empty), transformed dims, raw row width
  int64_t off_d1W, off_d1b, off_hW, off_hb;
};

struct Offsets {
  int64_t enc0W, enc0b, enc1W, enc1b, dec0W, dec0b, dec1W, dec1b, total;
};

inline void realnvp_split(int i, int D, int& cs0, int& cs1, int& ts0, int& ts1) {
  if (D == 1) { cs0 = cs1 = 0; ts0 = 0; ts1 = 1; return; }
  if (i % 2 == 0) { int m = D / 2; cs0 = 0; cs1 = m; ts0 = m; ts1 = D; return; }
  int m = D - D / 2;
  cs0 = D - m; cs1 = D; ts0 = 0; ts1 = D - m;
}

inline Offsets layout(const vms_elbo_desc& d, std::vector<FlowBlock>* blocks) {
  Offsets o;
  int64_t p = 0;
  o.enc0W = p; p += (int64_t)d.dx * d.hidden;
  o.enc0b = p; p += d.hidden;
  o.enc1W = p; p += (int64_t)d.hidden * 2 * d.dz;
  o.enc1b = p; p += 2 * d.dz;
  o.dec0W = p; p += (int64_t)d.dz * d.hidden;
  o.dec0b = p; p += d.hidden;
  o.dec1W = p; p += (int64_t)d.hidden * 2 * d.dx;
  o.dec1b = p; p += 2 * d.dx;
  for (int i = 0; i < d.num_blocks; ++i) {
    FlowBlock b;
    realnvp_split(i, d.dz, b.cs0, b.cs1, b.ts0, b.ts1);
    b.cin = b.cs1 - b.cs0 > 0 ? b.cs1 - b.cs0 : 1;
    b.dt = b.ts1 - b.ts0;
    b.ldr = b.dt * (3 * d.num_bins - 1);
    b.off_d1W = p; p += (int64_t)b.cin * d.flow_hidden;
    b.off_d1b = p; p += d.flow_hidden;
    b.off_hW = p; p += (int64_t)d.flow_hidden * b.ldr;
    b.off_hb = p; p += b.ldr;
    if (blocks) blocks->push_back(b);
  }
  o.total = p;
  return o;
}


struct FusedCfg;  // elbo_fused.cu
struct TcfCfg;    // elbo_tcf.cu (whole-step tensor-core kernel, plan mode 3 / auto)

}  // namespace vms

struct vms_elbo_plan_s {
  vms_elbo_desc d;
  vms::Offsets off;
  std::vector<vms::FlowBlock> blocks;
  int64_t maxB;
  int splits_max;
  // forward intermediates
  float *he, *pe, *z, *logq, *logpz, *logpx, *hd, *pd, *scalars, *partial;
  std::vector<float*> u;    // u[i], i = 0..num_blocks: chain-inverse states, u[num_blocks] = z, u[0] = base sample
  std::vector<float*> hid;  // [B, flow_hidden] per block
  std::vector<float*> raw;  // [B, ldr] per block
  // backward scratch
  float *g_pd, *g_hd, *g_z, *g_pe, *g_he, *g_ua, *g_ub, *g_raw, *g_hid, *g_ldj, *gpart;
  std::vector<void*> allocs;
  // graph cache: key = (mode, B, theta, x, eps, out pointers...)
  std::map<std::array<uintptr_t, 10>, std::pair<cudaGraphExec_t, int>> graphs;
  // fused path (elbo_fused.cu): NULL when the shape does not fit; mode 0 = auto (fused when available), 1 = unfused
  vms::FusedCfg* fused = nullptr;
  int mode = 0;
  // tensor-core flow blocks (flow_tc.cu): per-CTA weight-gradient partials [sm_count][flow parameters], error flag,
  // and the batch from which auto mode prefers them over the single fused kernel
  bool tc_ok = false;
  float* tc_part = nullptr;
  int* tc_err = nullptr;
  int64_t tc_auto_batch = 8192;
  bool tc_auto_user = false;  // threshold set by the caller / environment (overrides the whole-step kernel's multi-wave range)
  vms::TcfCfg* tcf = nullptr;  // whole-step tensor-core kernel; NULL when the shape does not fit
  // kernel timing (vms_elbo_plan_set_timing_every): every `timing_every`-th call of the fused paths is measured
  int timing_every = 1;
  unsigned timing_calls = 0;
};

namespace vms {
// elbo_fused.cu
vms_status fused_create(vms_elbo_plan_s* pl);   // sets pl->fused (or leaves NULL when the shape does not fit)
void fused_destroy(vms_elbo_plan_s* pl);
vms_status fused_set_timing(vms_elbo_plan_s* pl, int max_launches);
vms_status fused_kernel_ms(vms_elbo_plan_s* pl, double* total_ms, int* launches);
struct FusedAdam {  // optimiser step folded into the fused path's finishing kernel (vms_elbo_train_step)
  float *theta, *m, *v;
  float lr_t, one_minus_b1, one_minus_b2, eps;
};
// elbo_tcf.cu (plan mode 3 / auto for forward + backward up to one wave of 32-row tiles)
vms_status tcf_create(vms_elbo_plan_s* pl);
void tcf_destroy(vms_elbo_plan_s* pl);
bool tcf_available(const vms_elbo_plan_s* pl, int64_t B);
void tcf_invalidate(vms_elbo_plan_s* pl);  // the packed weight images no longer describe theta
vms_status tcf_set_timing(vms_elbo_plan_s* pl, int max_launches);
vms_status tcf_kernel_ms(vms_elbo_plan_s* pl, double* total_ms, int* launches);  // ADDS to both outputs
struct PeerArgs;
bool tcf_peer_ok(vms_elbo_plan_s* pl);
vms_status tcf_run(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B, float* grad,
                   float* scalars, cudaStream_t st, const FusedAdam* adam = nullptr, const PeerArgs* peer = nullptr);
vms_status fused_run(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B, bool backward,
                     float* z, float* logq, float* logpz, float* logpx, float* grad, float* scalars, cudaStream_t st,
                     const FusedAdam* adam = nullptr);
}  // namespace vms


