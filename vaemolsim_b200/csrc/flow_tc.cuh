// flow_tc.cuh -- internal interface of the tensor-core coupling-block kernels (flow_tc.cu) used by the ELBO plan.
#pragma once
#include "common.cuh"

namespace vms {

// One RealNVP-RQS coupling block in the density direction (chain inverse) over B rows of a [B, dz] chain state:
//   cond = uin[:, cs0:cs0+nc]  (nc == 0: the ones((B, 1)) input of flows.py:184-185)
//   hid  = tanh(cond @ d1W + d1b)                       [B, H]
//   raw  = hid @ hW + hb                                [B, 3K-1]  (widths | heights | slopes, flows.py:140-152)
//   uout[:, ts0] = RQS(raw).inverse(uin[:, ts0]);  logpz (+)= ildj;  uout[:, cs] = uin[:, cs]
// and its reverse mode (raw is recomputed, never stored).  One transformed dimension per block (dt == 1).
struct FlowTcArgs {
  int64_t B;
  int dz, cs0, nc, ts0;
  int H, K;
  float bin_min, bin_max;
  const float *d1W, *d1b, *hW, *hb;
  const float* uin;
  float* uout;
  float* logpz;
  int accumulate;
  // reverse mode: g_cur = d loss / d uout, g_ldj = d loss / d (sum of log-dets) (a constant: losses.py:253 is a mean),
  // g_nxt = d loss / d uin.  Weight-gradient partials: one [d1W | d1b | hW | hb] block per CTA at
  // part + cta * part_stride + o_*; the caller sums the flow_tc_grid(B) partials in a fixed order.
  const float* g_cur;
  float* g_nxt;
  float g_ldj;
  float* part;
  int64_t part_stride, o_d1W, o_d1b, o_hW, o_hb;
  int* err;  // device int, set non-zero if a tensor-core completion wait ran into its bound (results invalid)
};

bool flow_tc_supported(int dz, int cin, int dt, int H, int K);
int flow_tc_grid(int64_t B);
vms_status flow_tc_forward(const FlowTcArgs& a, cudaStream_t st);
vms_status flow_tc_backward(const FlowTcArgs& a, cudaStream_t st);

}  // namespace vms
