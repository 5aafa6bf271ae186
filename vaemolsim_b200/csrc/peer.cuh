// peer.cuh -- the peer-memory exchange protocol shared by peer.cu (allreduce + Adam kernel) and elbo_tcf.cu (the finish
// kernel that sums the tile partials, exchanges and updates in one launch).
#pragma once
#include "common.cuh"

namespace vms {

constexpr int kMaxPeers = 8;
constexpr int kFlagSlots = 64;

struct PeerArgs {
  int world, rank;
  float* base[kMaxPeers];  // every rank's buffer as mapped into THIS process ([rank] = own allocation)
  int64_t P;
  unsigned long long step;  // 1, 2, 3, ... (monotonic)
  float grad_scale;
  float *theta, *m, *v;
  float lr_t, one_minus_b1, one_minus_b2, eps;
  float* grad_out;  // optional: the reduced, scaled gradient [P]
  unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long* flags_of(float* base, int64_t P) {
  return reinterpret_cast<unsigned long long*>(base + 2 * P);
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Flag slots of a rank's buffer (uint64 each):  [0, 8) arrival flags written by the ranks;  32 + r: rank r gave up waiting
// at that step (written by r into its own buffer);  40: this rank's grid-wide decision for the current step (2 step + failed,
// written by block 0, read by the other blocks);  41: block counter of the fused finish + exchange kernel (local);
// 48: POISON -- some rank of the job timed out (written by that rank into every buffer): from then on no rank updates its
// parameters, so replicas differ by at most the one step in flight and the host (`PeerExchange.check`) raises.

// Called by ONE thread of block 0 after this rank's flag has been (or is being) raised: waits (bounded) for every rank's
// arrival flag of `step` in this rank's buffer, poisons the job on a time-out, publishes the grid-wide decision and returns
// it (1 = failed).
__device__ __forceinline__ int peer_wait_and_decide(const PeerArgs& a) {
  volatile unsigned long long* mine = flags_of(a.base[a.rank], a.P);
  int bad = mine[48] != 0ull ? 1 : 0;
  const unsigned long long t0 = globaltimer_ns();
  for (int r = 0; r < a.world && !bad; ++r) {
    while (mine[r] < a.step) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > a.timeout_ns || mine[48] != 0ull) {
        bad = 1;
        break;
      }
    }
  }
  if (bad) {
    mine[32 + a.rank] = a.step;
    for (int r = 0; r < a.world; ++r) flags_of(a.base[r], a.P)[48] = a.step;  // poison every replica
  }
  __threadfence_system();
  mine[40] = 2ull * a.step + (unsigned long long)bad;
  return bad;
}

// the other blocks: wait for block 0's decision of this step
__device__ __forceinline__ int peer_wait_decision(const PeerArgs& a) {
  volatile unsigned long long* mine = flags_of(a.base[a.rank], a.P);
  unsigned long long d;
  while ((d = mine[40]) < 2ull * a.step) __nanosleep(32);
  __threadfence();
  return (int)(d & 1ull);
}

__device__ __forceinline__ void peer_raise_flags(const PeerArgs& a) {
  __threadfence_system();
  for (int r = 0; r < a.world; ++r) {
    volatile unsigned long long* f = flags_of(a.base[r], a.P) + a.rank;
    *f = a.step;
  }
  __threadfence_system();
}

// sum of one parameter's gradient over the ranks' slots, in rank order (identical on every rank).  Peer memory over NVLink:
// volatile = never served from this SM's L1; all loads are issued before the first addition.
__device__ __forceinline__ float peer_pull_sum(const PeerArgs& a, int64_t i) {
  const int64_t off = (int64_t)(a.step & 1ull) * a.P + i;
  float val[kMaxPeers];
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r)
    val[r] = r < a.world ? *reinterpret_cast<volatile const float*>(a.base[r] + off) : 0.f;
  float g = 0.f;
#pragma unroll
  for (int r = 0; r < kMaxPeers; ++r) g += val[r];
  return g * a.grad_scale;
}

unsigned long long peer_timeout_ns();
vms_status peer_fill_args(PeerArgs& a, int world, int rank, void* const* peer_bases, int64_t n_params, unsigned long long step,
                          float grad_scale);

}  // namespace vms
