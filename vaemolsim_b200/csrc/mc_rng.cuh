// mc_rng.cuh -- device random streams and lane helpers shared by the fused MC kernels (mc_chain.cu, mc_nb.cu).
#pragma once
#include "common.cuh"
#include <math.h>

namespace vms {
namespace mcdev {

// ------------------------------------------------------------------------------------------------ PCG64 on the device
// mcmc.py:119 draws the accept uniforms as `self._rng.random(size=B)` from NumPy's default generator: PCG64 = a 128-bit
// LCG (state <- state * M + inc mod 2^128, M = 0x2360ED051FC65DA4_4385DF649FCCF645) with the XSL-RR 128/64 output function,
// u = (out >> 11) * 2^-53, filled in chain order, step after step.  Chain c at step k therefore owns draw number
// k * B_global + c of the stream: an LCG jumps ahead in O(log n) multiplications, so every chain derives its own
// sub-sequence (state after c + 1 steps, then an affine jump by B_global per MC step, both exact integer arithmetic) and
// the host's sequential draw + np.log (35 ms per 100 x 65,536 block, the bound of round 1's end-to-end MC number)
// disappears together with its 8-byte-per-proposal upload.  u is bit-identical to NumPy's.  log(u) is CUDA's double log
// (<= 1 ulp) where the reference takes np.log: the decision log_acc >= log u can only differ if the two sides agree to
// ~1e-13 relative; such chain-steps are COUNTED (n_uncertain) and the host re-runs the call on the NumPy stream when
// the counter is non-zero (probability ~1e-6 per 6.5 M proposals), so decisions stay those of mcmc.py:116-120.
struct U128 {
  unsigned long long hi, lo;
};
__device__ __forceinline__ U128 mul128(U128 a, U128 b) {
  U128 r;
  r.lo = a.lo * b.lo;
  r.hi = __umul64hi(a.lo, b.lo) + a.hi * b.lo + a.lo * b.hi;
  return r;
}
__device__ __forceinline__ U128 add128(U128 a, U128 b) {
  U128 r;
  r.lo = a.lo + b.lo;
  r.hi = a.hi + b.hi + (r.lo < a.lo ? 1ull : 0ull);
  return r;
}
// state after `delta` steps (pcg_advance_lcg_128)
__device__ __forceinline__ U128 pcg_advance(U128 state, U128 inc, unsigned long long delta) {
  U128 acc_m = {0ull, 1ull}, acc_p = {0ull, 0ull};
  U128 cur_m = {0x2360ED051FC65DA4ull, 0x4385DF649FCCF645ull}, cur_p = inc;
  while (delta > 0) {
    if (delta & 1ull) {
      acc_m = mul128(acc_m, cur_m);
      acc_p = add128(mul128(acc_p, cur_m), cur_p);
    }
    cur_p = mul128(add128(cur_m, U128{0ull, 1ull}), cur_p);
    cur_m = mul128(cur_m, cur_m);
    delta >>= 1;
  }
  return add128(mul128(acc_m, state), acc_p);
}
// XSL-RR output of a state, as the double NumPy's Generator.random() returns
__device__ __forceinline__ double pcg_uniform(U128 s) {
  const unsigned long long v = s.hi ^ s.lo;
  const unsigned rot = (unsigned)(s.hi >> 58);
  const unsigned long long out = (v >> rot) | (v << ((64u - rot) & 63u));
  return (double)(out >> 11) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ void box_muller(unsigned a, unsigned b, float& n0, float& n1) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.f * logf(u1));
  float s, c;
  sincospif(2.f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// sum over the 4 lanes of a chain; every lane gets the same value: (l0 + l1) + (l2 + l3)
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

}  // namespace mcdev
}  // namespace vms
