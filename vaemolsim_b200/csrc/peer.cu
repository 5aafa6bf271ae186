// peer.cu -- data-parallel training's one exchange step as ONE kernel over NVLink peer memory:
// gradient all-reduce (sum over ranks) fused with the Keras Adam update.
//
// The reference never distributes; the north_star adds "a single NCCL-over-NVLink gradient allreduce per
// step".  The gradient is 44,396 floats (177 KB): an NCCL allreduce of that size is pure latency (~15-25 us: a launch,
// a ring / tree protocol, a second launch for Adam), so here every rank
//   1. publishes "my gradient of step s is in my buffer" by writing s into its flag slot in EVERY peer's buffer,
//   2. spins until all ranks' flags in its OWN buffer reach s,
//   3. reads each parameter's gradient from all peers' buffers directly over NVLink (one-shot all-gather-reduce: each
//      GPU pulls world x 177 KB), sums in rank order -- the same order on every rank, so replicas stay bit-identical --
//      and applies Adam to its local theta / m / v in the same thread.
// No host round trip, no second kernel, and the sum is deterministic.
//
// Failure handling: the wait is bounded in wall-clock time (VMS_PEER_TIMEOUT_MS, default 2000).  A rank that gives up does
// NOT apply an update, poisons every replica's buffer so nobody updates afterwards, and `PeerExchange.check()` raises on
// the host -- a stalled peer can no longer silently desynchronise the replicas.
//
// Buffers: each rank owns one device allocation [2][P] floats (gradient slots, double-buffered by step parity) +
// 64 x uint64 flags, shared with the peers through CUDA IPC (vms_ipc_*; handles travel over torch.distributed, which
// stays the plumbing).  Double buffering is what makes a single barrier per step enough: a rank overwrites slot s & 1
// at step s + 2, after it has passed the barrier of step s + 1, which every peer can only have signalled after its own
// step-s kernel -- the last reader of slot s & 1 -- completed.
// One process per GPU: the kernel waits on OTHER GPUs' kernels, never on another kernel of the same GPU.
#include "peer.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace vms {

__global__ void __launch_bounds__(256) peer_allreduce_adam_kernel(const PeerArgs a) {
  __shared__ int failed;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      // the gradient of this step was written by the preceding kernel on this stream; make it visible system-wide,
      // then raise my flag in every rank's buffer (including my own) and wait for every rank's flag in MY buffer.
      // Bounded (wall-clock nanoseconds, independent of the SM clock): a rank that never arrives must not hang this GPU.
      // ONE block decides for the grid, so all blocks agree.
      peer_raise_flags(a);
      failed = peer_wait_and_decide(a);
    } else {
      // block 0 is always dispatched first and waits on nothing of this grid: no co-residency requirement
      failed = peer_wait_decision(a);
    }
  }
  __syncthreads();
  if (failed) return;  // no update on a failed exchange: theta / m / v keep the last consistent state
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  const float g = peer_pull_sum(a, i);
  if (a.grad_out) a.grad_out[i] = g;
  const float mi = a.m[i] + (g - a.m[i]) * a.one_minus_b1;
  const float vi = a.v[i] + (g * g - a.v[i]) * a.one_minus_b2;
  a.m[i] = mi;
  a.v[i] = vi;
  a.theta[i] = a.theta[i] - a.lr_t * mi / (sqrtf(vi) + a.eps);
}

unsigned long long peer_timeout_ns() {
  static unsigned long long timeout_ms = 0;
  if (!timeout_ms) {
    const char* e = getenv("VMS_PEER_TIMEOUT_MS");
    timeout_ms = e && atoll(e) > 0 ? (unsigned long long)atoll(e) : 2000ull;
  }
  return timeout_ms * 1000000ull;
}

vms_status peer_fill_args(PeerArgs& a, int world, int rank, void* const* peer_bases, int64_t n_params, unsigned long long step,
                          float grad_scale) {
  VMS_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, VMS_ERR_INVALID_ARG,
              "peer exchange: world must be in [1, %d] and 0 <= rank < world", kMaxPeers);
  VMS_REQUIRE(peer_bases && n_params >= 1 && step >= 1, VMS_ERR_INVALID_ARG, "peer exchange: bad arguments");
  a.world = world; a.rank = rank; a.P = n_params; a.step = step; a.grad_scale = grad_scale;
  for (int r = 0; r < world; ++r) {
    VMS_REQUIRE(peer_bases[r], VMS_ERR_INVALID_ARG, "peer exchange: NULL peer buffer %d", r);
    a.base[r] = (float*)peer_bases[r];
  }
  a.timeout_ns = peer_timeout_ns();
  return VMS_OK;
}

}  // namespace vms

using namespace vms;

extern "C" {

size_t vms_peer_buffer_bytes(int64_t n_params) { return (size_t)(2 * n_params) * sizeof(float) + kFlagSlots * sizeof(unsigned long long); }

vms_status vms_ipc_get_handle(void* device_ptr, unsigned char handle[64]) {
  VMS_REQUIRE(device_ptr && handle, VMS_ERR_INVALID_ARG, "ipc_get_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  VMS_CUDA(cudaIpcGetMemHandle(&h, device_ptr));
  memcpy(handle, &h, 64);
  return VMS_OK;
}

vms_status vms_ipc_open_handle(const unsigned char handle[64], void** device_ptr) {
  VMS_REQUIRE(handle && device_ptr, VMS_ERR_INVALID_ARG, "ipc_open_handle: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  VMS_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VMS_OK;
}

vms_status vms_ipc_close_handle(void* device_ptr) {
  if (!device_ptr) return VMS_OK;
  VMS_CUDA(cudaIpcCloseMemHandle(device_ptr));
  return VMS_OK;
}

vms_status vms_peer_allreduce_adam(int world, int rank, void* const* peer_bases, int64_t n_params, unsigned long long step,
                                   float grad_scale, float* theta, float* m, float* v, int64_t t, double lr, double beta1,
                                   double beta2, double eps, float* grad_out, vms_stream stream) {
  VMS_RANGE("vms_peer_allreduce_adam");
  VMS_REQUIRE(theta && m && v && t >= 1, VMS_ERR_INVALID_ARG, "peer_allreduce_adam: bad arguments");
  PeerArgs a = {};
  vms_status fs = peer_fill_args(a, world, rank, peer_bases, n_params, step, grad_scale);
  if (fs) return fs;
  a.theta = theta; a.m = m; a.v = v; a.grad_out = grad_out;
  a.lr_t = (float)(lr * sqrt(1.0 - pow(beta2, (double)t)) / (1.0 - pow(beta1, (double)t)));
  a.one_minus_b1 = (float)(1.0 - beta1);
  a.one_minus_b2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  peer_allreduce_adam_kernel<<<(unsigned)((n_params + 255) / 256), 256, 0, as_stream(stream)>>>(a);
  VMS_LAUNCH_CHECK("peer_allreduce_adam_kernel");
  return VMS_OK;
}

}  // extern "C"
