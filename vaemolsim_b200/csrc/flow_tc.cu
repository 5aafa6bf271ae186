// flow_tc.cu -- a RealNVP-RQS coupling block at scale: conditioner MLP on the 5th-generation tensor cores (tcgen05,
// kind::f16 on a 3 x BF16 split of every float32 operand, accumulators in TMEM) fused with the spline evaluation,
// forward and reverse mode.
//
// Replaces, for batches where the conditioner is a real contraction ([B, 100] x [100, 95] per block, 88 % of the
// model's FLOPs), the chain  flows.py:184-196 (SplineBijector.call: d1 + the three heads)  ->  flows.py:86-101
// (softmax / softplus activations)  ->  tfp RationalQuadraticSpline.inverse + inverse_log_det_jacobian as driven by
// tfp RealNVP (flows.py:312) inside TransformedDistribution.log_prob (dists.py:414-439), and TF autodiff through it.
//
// Why one kernel per block.  Unfused, a block round-trips hid [B, 100] and raw [B, 95] through HBM in the forward pass
// and again (plus their gradients) in the backward pass: ~3 KB per row per block against 24 bytes of real input /
// output.  Here a CTA owns a 64-row tile: it builds hid in shared memory straight from the conditioner column(s),
// multiplies by the heads matrix on the tensor core, copies raw from TMEM to shared memory and runs the octet spline
// routines of rqs_device.cuh on it.  The backward kernel RECOMPUTES hid and raw (tensor-core time is cheap, HBM is
// not), runs the spline's reverse mode, and feeds the raw-logit gradients to two more tensor-core products:
//   d hid = g_raw @ hW^T          (A = g_raw, K-major; B = the SAME shared-memory copy of hW read MN-major)
//   d hW  = hid^T @ g_raw         (A = the hid tile read MN-major; B = the g_raw tile read MN-major; M = 128)
// so no operand is ever transposed in memory.  The ones column appended to hid folds hb into the forward product and
// makes d hb a row of d hW.  Weight gradients accumulate in registers across the CTA's tiles (the TMEM accumulator is
// drained per tile: tensor-core float32 accumulation is not round-to-nearest, so long in-TMEM sums drift) and leave
// the CTA once, as a per-CTA partial that the caller reduces in a fixed order (deterministic).
//
// float32 parity on 16-bit tensor cores.  a = a1 + a2 + a3 with three bfloat16 parts (8 significant bits each,
// truncation splits: exact for zero and every |a| > 2^-110) and a.b ~ a1 b1 + (a1 b2 + a2 b1) + (a1 b3 + a2 b2 + a3 b1): six
// kind::f16 MMAs per 16-deep k-step, the dropped terms are 2^-24 relative.  The leading product has its own TMEM
// accumulator, the five corrections share a second one (added in the epilogue), as in gemm_tc.cu.
// Why not 3 x TF32 as in gemm_tc.cu: kind::tf32 reads MN-major (transposed) operands only in the 128B-base-32B swizzled
// layout -- measured: a no-swizzle MN-major tf32 descriptor returns zeros -- so the two transposed products would need
// second, differently laid out copies of hid, g_raw and hW, which do not fit 227 KB.  16-bit operands can be read
// K-major and MN-major from the same no-swizzle tile, the split costs 6 bytes per element instead of 8, and the f16
// pipe runs at twice the tf32 rate, so six MMAs cost what three did.
//
// Layouts.  A-type tiles ([64 rows] x [k]) are stored as 16-byte chunks of 8 bf16: [k/8][65][8] (one pad slot per
// chunk column spreads the octets' stores over the banks).  Read K-major that is the canonical no-swizzle layout with
// LBO = 1040 B (chunk stride), SBO = 128 B (8 rows); read MN-major (the "row" index becomes K) it is the canonical
// MN-major no-swizzle layout with SBO = 1040 B (8-element group stride), LBO = 128 B (8 K rows), 256 B per k-step.
// hW is stored [j/8][Rp][8] (n = raw column, k = hidden unit j): K-major for the forward product (LBO = 16 Rp,
// SBO = 128), MN-major for d hid (n = j, k = raw column: SBO = 16 Rp, LBO = 128).
// M = 64 accumulators occupy lanes 0-15 of each 32-lane TMEM quarter (row = 16 quarter + lane).
#include "flow_tc.cuh"
#include "rqs_device.cuh"
#include <math.h>
#include <stdlib.h>

namespace vms {

namespace {

constexpr int FM = 64;               // rows per tile (UMMA M of the row-major products)
constexpr int FT = 512;              // worker threads per CTA: one octet per row in the spline phase
constexpr int FTA = FT + 32;         // + the warp that issues the tensor-core instructions
constexpr unsigned CSB = (FM + 1) * 16;  // bytes per chunk column of an A-type tile (64 rows + one pad slot)
constexpr int kMaxC = 4;             // conditioner columns supported

struct KParams {
  FlowTcArgs a;
  int Hp, R, LDS, cin;
  float scale;
  // shared-memory offsets (bytes)
  int o_hid, o_graw, o_w, o_raw, o_w1, o_b1, o_c, o_v, o_g, o_lp, o_gin;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// a = p1 + p2 + p3 exactly (zero and |a| > 2^-110; tests/test_split_numerics.py), each part a bfloat16 (upper half of a
// float32 pattern)
__device__ __forceinline__ void split3(float a, unsigned& h1, unsigned& h2, unsigned& h3) {
  h1 = __float_as_uint(a) & 0xffff0000u;
  const float r1 = a - __uint_as_float(h1);
  h2 = __float_as_uint(r1) & 0xffff0000u;
  const float r2 = r1 - __uint_as_float(h2);
  h3 = __float_as_uint(r2) & 0xffff0000u;
}
// two bfloat16 (upper halves of e0, e1) packed with e0 in the low half
__device__ __forceinline__ unsigned pack2(unsigned e0, unsigned e1) { return __byte_perm(e0, e1, 0x7632); }
__device__ __forceinline__ float bf_lo(unsigned v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(unsigned v) { return __uint_as_float(v & 0xffff0000u); }

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (see gemm_tc.cu; K-major: LBO = K-chunk stride, SBO = 8-row
// group stride; MN-major: LBO = 8-K-row group stride, SBO = MN-group stride)
__device__ __forceinline__ unsigned long long make_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_addr & 0x3FFFFu) >> 4);
  d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void mma_bf16(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void mma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Bounded wait for a phase of the MMA-completion mbarrier: a wrong descriptor must not hang the GPU.
__device__ __forceinline__ bool mbar_wait_bounded(unsigned bar, unsigned phase) {
  const long long t0 = clock64();
  for (;;) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(phase)
        : "memory");
    if (ok) return true;
    if (clock64() - t0 > 2000000000LL) return false;  // ~1 s
  }
}

__device__ __forceinline__ void tc_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// The 16 worker warps synchronise among themselves on named barrier 1; the MMA warp never takes part, so a worker never
// waits for the issue of tensor-core instructions, only (at an mbarrier) for their completion.
__device__ __forceinline__ void worker_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  asm volatile("bar.sync 1, %0;\n" ::"n"(FT) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// workers -> MMA warp: "the operands of the next product are in shared memory / its TMEM target is free"
__device__ __forceinline__ void operands_ready(int id) {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "n"(FTA) : "memory");
}
__device__ __forceinline__ void wait_operands(int id) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "n"(FTA) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// instruction descriptor: D = F32, A = B = BF16, optional MN-major operands, N >> 3, M >> 4
__device__ __forceinline__ unsigned idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
         ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// One product C = A B with float32-grade accuracy from 3 x BF16 parts, ONE TMEM accumulator: the correction terms go in
// first, smallest first (a1 b3, a3 b1, a2 b2 ~ 2^-16; then a1 b2, a2 b1 ~ 2^-8), the leading a1 b1 last.  The tensor
// core's float32 accumulation is not round-to-nearest and its error grows with the magnitude of the running sum, so
// only the n_k leading steps accumulate at full magnitude.
// a0 / b0: shared-memory address of part 1 at k-step 0; a_part / b_part: bytes between parts; a_step / b_step: bytes
// per k-step; (al, as) / (bl, bs): the operands' LBO / SBO.
// ONE thread issues every MMA of the CTA and the other 511 wait for it at the next barrier, so the issue path is kept
// to an add per operand: the descriptors of the three parts are built once, a k-step / part offset only changes the
// 14-bit start-address field (shared-memory addresses are < 256 KB: no carry out of the field).
__device__ __forceinline__ void issue_product(unsigned acc, unsigned a0, unsigned a_part, unsigned a_step, unsigned al,
                                              unsigned as, unsigned b0, unsigned b_part, unsigned b_step, unsigned bl,
                                              unsigned bs, int n_k, unsigned idesc) {
  const unsigned long long da1 = make_desc(a0, al, as), db1 = make_desc(b0, bl, bs);
  const unsigned long long pa = a_part >> 4, pb = b_part >> 4, sa = a_step >> 4, sb = b_step >> 4;
  const unsigned long long da2 = da1 + pa, da3 = da2 + pa, db2 = db1 + pb, db3 = db2 + pb;
  unsigned first = 0u;
  unsigned long long ka = 0, kb = 0;
#pragma unroll 1
  for (int ks = 0; ks < n_k; ++ks, ka += sa, kb += sb) {
    mma_bf16(acc, da1 + ka, db3 + kb, idesc, first);
    first = 1u;
    mma_bf16(acc, da3 + ka, db1 + kb, idesc, 1u);
    mma_bf16(acc, da2 + ka, db2 + kb, idesc, 1u);
  }
  ka = kb = 0;
#pragma unroll 1
  for (int ks = 0; ks < n_k; ++ks, ka += sa, kb += sb) {
    mma_bf16(acc, da1 + ka, db2 + kb, idesc, 1u);
    mma_bf16(acc, da2 + ka, db1 + kb, idesc, 1u);
  }
  ka = kb = 0;
#pragma unroll 1
  for (int ks = 0; ks < n_k; ++ks, ka += sa, kb += sb) mma_bf16(acc, da1 + ka, db1 + kb, idesc, 1u);
}

// Software pipeline (per CTA, tiles t0, t1, ... of 64 rows; hid / inputs / the raw accumulator are double-buffered).
// 16 worker warps + ONE MMA warp: the workers signal "operands ready" on a named barrier they only ARRIVE at, the MMA warp
// waits there, issues the products and commits them to an mbarrier the workers wait on when they need the result
// (round 1 had worker thread 0 issue: 42-102 MMAs per tile queue up behind the tensor pipe, and the other 511 threads
// waited for that thread at the next CTA barrier -- ncu: 22 % of all stall samples at that barrier).
//   forward  workers:  inputs + hid of tile i+1 -> ready(hid) | wait raw(i) -> TMEM -> smem -> spline(i)
//            MMA warp: wait ready(hid) -> raw(i+1)
//   backward workers:  wait raw(i) -> TMEM -> smem, spline reverse mode(i) -> ready(g_raw) -> inputs + hid of tile i+1
//                      -> ready(hid) -> wait d hid / d hW(i) -> epilogues(i)
//            MMA warp: wait ready(g_raw) -> d hid(i), d hW(i);  wait ready(hid) -> raw(i+1)
template <int RP, bool BWD>
__global__ void __launch_bounds__(FTA, 1) flow_tc_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(128) unsigned char smb[];
  __shared__ __align__(8) unsigned long long mbar[4];  // [0], [1]: raw of buffer 0 / 1;  [2]: d hid;  [3]: d hW
  __shared__ unsigned tmem_base_s;
  const FlowTcArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = a.H, Hp = p.Hp, K = a.K, R = p.R, LDS = p.LDS, nc = a.nc, dz = a.dz, cin = p.cin;
  const int nchH = Hp / 8;
  constexpr int nchR = RP / 8;
  const unsigned hid_sz = (unsigned)nchH * CSB;       // bytes of one part of one hid buffer
  constexpr unsigned graw_sz = (unsigned)nchR * CSB;  // ... of the g_raw tile
  const unsigned w_cs = RP * 16u;                     // chunk-column stride of the heads matrix
  const unsigned w_sz = (unsigned)nchH * w_cs;
  unsigned char* hid = smb + p.o_hid;    // [2 buffers][3 parts]
  unsigned char* graw = smb + p.o_graw;  // [3 parts] (BWD)
  unsigned char* wsm = smb + p.o_w;      // [3 parts]
  float* s_raw = reinterpret_cast<float*>(smb + p.o_raw);  // [FM][LDS] raw spline parameters / d pre-activation (BWD)
  float* s_w1 = reinterpret_cast<float*>(smb + p.o_w1);    // [4][Hp], rows >= cin zero
  float* s_b1 = reinterpret_cast<float*>(smb + p.o_b1);    // [Hp]
  float* s_c = reinterpret_cast<float*>(smb + p.o_c);      // [2][FM][4] conditioner columns (1 in column 0 when nc == 0)
  float* s_v = reinterpret_cast<float*>(smb + p.o_v);      // [2][FM] value to transform
  float* s_g = reinterpret_cast<float*>(smb + p.o_g);      // [2][FM] upstream gradient of the transformed column (BWD)
  float* s_lp = reinterpret_cast<float*>(smb + p.o_lp);    // [2][FM] log-det accumulated so far (FWD, accumulate)
  float* s_gin = reinterpret_cast<float*>(smb + p.o_gin);  // [FM] gradient wrt the spline input

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  const bool worker = tid < FT;  // warp 16 issues the tensor-core instructions and does nothing else
  // heads matrix with the bias as row H (the ones column of hid multiplies it), split, zero padded to [Hp, RP];
  // four independent loads in flight per thread
  for (int e0 = 0; worker && e0 < Hp * RP; e0 += 4 * FT) {
    float w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * FT + tid;
      const int j = e / RP, c = e - j * RP;
      w[u] = 0.f;
      if (e < Hp * RP && c < R) {
        if (j < H) w[u] = __ldg(a.hW + (size_t)j * R + c);
        else if (j == H) w[u] = __ldg(a.hb + c);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * FT + tid;
      if (e < Hp * RP) {
        const int j = e / RP, c = e - j * RP;
        unsigned h1, h2, h3;
        split3(w[u], h1, h2, h3);
        const unsigned o = (unsigned)(j >> 3) * w_cs + (unsigned)c * 16u + (unsigned)(j & 7) * 2u;
        *reinterpret_cast<unsigned short*>(wsm + o) = (unsigned short)(h1 >> 16);
        *reinterpret_cast<unsigned short*>(wsm + w_sz + o) = (unsigned short)(h2 >> 16);
        *reinterpret_cast<unsigned short*>(wsm + 2 * w_sz + o) = (unsigned short)(h3 >> 16);
      }
    }
  }
  for (int e = tid; worker && e < 4 * Hp; e += FT) {
    const int c = e / Hp, j = e - c * Hp;
    s_w1[e] = (c < cin && j < H) ? __ldg(a.d1W + (size_t)c * H + j) : 0.f;
  }
  for (int j = tid; worker && j < Hp; j += FT) s_b1[j] = j < H ? __ldg(a.d1b + j) : 0.f;
  if (BWD)
    for (unsigned e = tid; worker && e < 3 * graw_sz / 4; e += FT) reinterpret_cast<unsigned*>(graw)[e] = 0u;

  const int64_t n_tiles = (a.B + FM - 1) / FM;
  // ---- S1: a tile's inputs.  Threads 0..63 load their row of tile t into registers (pre) one tile ahead of its use and
  // put them into buffer b an iteration later, so the global-load latency hides behind a whole tile of work.
  float pre[6];  // v, cond[0..3], g (BWD) or the log-det accumulated so far (FWD)
  auto load_inputs = [&](int64_t tile) {
    if (tid < FM) {
      const int64_t row = tile * FM + tid;
      const bool ok = tile < n_tiles && row < a.B;
      const float* ur = a.uin + row * dz;
      pre[0] = ok ? __ldg(ur + a.ts0) : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) pre[1 + c] = (c < nc && ok) ? __ldg(ur + a.cs0 + c) : ((nc == 0 && c == 0) ? 1.f : 0.f);
      if (BWD) pre[5] = ok ? __ldg(a.g_cur + row * dz + a.ts0) : 0.f;
      else pre[5] = (ok && a.accumulate) ? a.logpz[row] : 0.f;
    }
  };
  auto put_inputs = [&](int b) {
    if (tid < FM) {
      s_v[b * FM + tid] = pre[0];
      *reinterpret_cast<float4*>(s_c + (b * FM + tid) * 4) = make_float4(pre[1], pre[2], pre[3], pre[4]);
      if (BWD) s_g[b * FM + tid] = pre[5];
      else s_lp[b * FM + tid] = pre[5];
    }
  };
  // ---- S2: hid = tanh(cond @ d1W + d1b), ones column at j = H, split into buffer b (the A operand)
  auto build_hid = [&](int b) {
    const int r = tid & (FM - 1);
    const float4 cn = *reinterpret_cast<const float4*>(s_c + (b * FM + r) * 4);
    const float cnd[4] = {cn.x, cn.y, cn.z, cn.w};
    unsigned char* hb_ = hid + (unsigned)b * 3u * hid_sz;
    for (int jq = tid >> 6; jq < 2 * nchH; jq += FT / FM) {  // jq: group of 4 hidden units = half a chunk
      const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + 4 * jq);
      float pre[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < cin) {
          const float4 w4 = *reinterpret_cast<const float4*>(s_w1 + c * Hp + 4 * jq);
          pre[0] = fmaf(cnd[c], w4.x, pre[0]); pre[1] = fmaf(cnd[c], w4.y, pre[1]);
          pre[2] = fmaf(cnd[c], w4.z, pre[2]); pre[3] = fmaf(cnd[c], w4.w, pre[3]);
        }
      }
      unsigned h1[4], h2[4], h3[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = 4 * jq + q;
        const float h = j < H ? tanhf(pre[q]) : (j == H ? 1.f : 0.f);
        split3(h, h1[q], h2[q], h3[q]);
      }
      const unsigned o = (unsigned)(jq >> 1) * CSB + (unsigned)r * 16u + (unsigned)(jq & 1) * 8u;
      *reinterpret_cast<uint2*>(hb_ + o) = make_uint2(pack2(h1[0], h1[1]), pack2(h1[2], h1[3]));
      *reinterpret_cast<uint2*>(hb_ + hid_sz + o) = make_uint2(pack2(h2[0], h2[1]), pack2(h2[2], h2[3]));
      *reinterpret_cast<uint2*>(hb_ + 2 * hid_sz + o) = make_uint2(pack2(h3[0], h3[1]), pack2(h3[2], h3[3]));
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  };
  tc_sync();
  const unsigned tm = tmem_base_s;
  const unsigned tD2 = tm + 2 * RP, tD3 = tm + 2 * RP + 128;  // raw of buffer b: tm + b * RP
  const unsigned id1 = idesc_bf16(FM, RP, 0, 0);
  const unsigned id2 = idesc_bf16(FM, Hp, 0, 1);
  const unsigned id3 = idesc_bf16(128, RP, 1, 1);
  const unsigned hid_a = smem_u32(hid), graw_a = smem_u32(graw), w_a = smem_u32(wsm);
  // raw = [hid, 1] @ [hW; hb]   (M = 64, N = RP, K = Hp in steps of 16), buffer b
  auto issue_raw = [&](int b) {
    issue_product(tm + (unsigned)b * RP, hid_a + (unsigned)b * 3u * hid_sz, hid_sz, 2u * CSB, CSB, 128, w_a, w_sz, 2u * w_cs,
                  w_cs, 128, Hp / 16, id1);
    mma_commit(smem_u32(&mbar[b]));
  };
  bool failed = false;

  // reverse-mode accumulators, CTA lifetime.  d [hW; hb]: thread (TMEM quarter q = warp & 3, lane) owns hidden unit
  // 32 q + lane and the 16-column chunks sub, sub + 4 (sub = warp >> 2) of the raw columns.
  float acc_w[2][16];
  float acc_b1 = 0.f, acc_w1[kMaxC];
  if (BWD) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc_w[0][i] = acc_w[1][i] = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) acc_w1[c] = 0.f;
  }

  // named barriers: 1 = the workers among themselves; 2, 3 = "hid of tile it is built" (alternating: a worker may be one
  // tile ahead of the MMA warp in the forward kernel); 4 = "g_raw of this tile is written" (backward)
  if (!worker) {
    // ---- the MMA warp: same tile loop, only the products
    int it = 0;
    wait_operands(2);
    if (lane == 0) issue_raw(0);
#pragma unroll 1
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int b = it & 1;
      const bool has_next = tile + gridDim.x < n_tiles;
      if (BWD) {
        // d hid = g_raw @ hW^T (M = 64, N = Hp, K = RP)  and  d [hW; hb] = [hid, 1]^T @ g_raw (M = 128, N = RP, K = 64)
        wait_operands(4);
        if (lane == 0) {
          issue_product(tD2, graw_a, graw_sz, 2u * CSB, CSB, 128, w_a, w_sz, 256u, 128, w_cs, RP / 16, id2);
          mma_commit(smem_u32(&mbar[2]));  // S7a starts on d hid while d hW is still in the pipe
          issue_product(tD3, hid_a + (unsigned)b * 3u * hid_sz, hid_sz, 256u, 128, CSB, graw_a, graw_sz, 256u, 128, CSB,
                        FM / 16, id3);
          mma_commit(smem_u32(&mbar[3]));
        }
        __syncwarp();
      }
      if (has_next) {
        wait_operands(2 + ((it + 1) & 1));
        if (lane == 0) issue_raw(b ^ 1);
        __syncwarp();
      }
    }
  } else {
  // ---- the workers
  // pipeline prologue: tile 0 of this CTA; the loads of tile 1 are in flight from here on
  load_inputs(blockIdx.x);
  put_inputs(0);
  load_inputs((int64_t)blockIdx.x + gridDim.x);
  worker_sync();
  build_hid(0);
  operands_ready(2);

  int it = 0;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int b = it & 1;
    const int64_t row0 = tile * FM;
    const int nr = (int)min((int64_t)FM, a.B - row0);
    const int64_t next = tile + gridDim.x;
    const bool has_next = next < n_tiles;
    if (!BWD && has_next) {
      // next tile's inputs and hid while the tensor core computes this tile's raw parameters
      put_inputs(b ^ 1);
      load_inputs(next + gridDim.x);
      worker_sync();
      build_hid(b ^ 1);
      operands_ready(2 + ((it + 1) & 1));
    }
    if (!mbar_wait_bounded(smem_u32(&mbar[b]), (unsigned)(it >> 1) & 1u)) failed = true;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    // ---- S4: TMEM -> shared memory (rows 16 q + lane live in lanes 0-15 of TMEM quarter q)
    {
      const int q = warp & 3, row = 16 * q + lane;
      for (int c0 = 16 * (warp >> 2); c0 < RP; c0 += 64) {
        float v1[16];
        tmem_ld16(tm + (unsigned)b * RP + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(s_raw + row * LDS + c0 + i) = make_float4(v1[i], v1[i + 1], v1[i + 2], v1[i + 3]);
        }
      }
    }
    worker_sync();
    // ---- S5: the spline, one octet per row
    {
      const int r = tid >> 3, j = tid & 7;
      const bool ok = r < nr;
      const float* rr = s_raw + r * LDS;
      const float v = s_v[b * FM + r];
      if (!BWD) {
        float out, ldj, ldj_all;
        bool writer;
        rqsdev::octet_apply<4, true, false>(rr, rr + K, rr + 2 * K, v, j, K, true, a.bin_min, p.scale, out, ldj, ldj_all,
                                            writer);
        if (ok) {
          float* uo = a.uout + (row0 + r) * dz;
          if (writer) uo[a.ts0] = out;
          if (j < nc) uo[a.cs0 + j] = s_c[(b * FM + r) * 4 + j];
          if (j == 0) a.logpz[row0 + r] = s_lp[b * FM + r] + ldj_all;
        }
      } else {
        float g_in, gw[4], gh[4], gs[4];
        bool writer;
        rqsdev::octet_backward<4, true, false>(rr, rr + K, rr + 2 * K, v, s_g[b * FM + r], ok ? a.g_ldj : 0.f, j, K, true,
                                               a.bin_min, p.scale, g_in, writer, gw, gh, gs);
        if (writer) s_gin[r] = g_in;
        if (4 * j < K) {
          // raw columns: widths 4 j .. 4 j + 3, heights K + 4 j .., slopes 2 K + 4 j .. (slope K - 1 does not exist):
          // half a 16-byte chunk each
          if (4 * j + 3 >= K - 1) gs[3] = 0.f;
#pragma unroll
          for (int arr = 0; arr < 3; ++arr) {
            const float* g4 = arr == 0 ? gw : (arr == 1 ? gh : gs);
            unsigned h1[4], h2[4], h3[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split3(g4[q], h1[q], h2[q], h3[q]);
            const int c = arr * K + 4 * j;
            const unsigned o = (unsigned)(c >> 3) * CSB + (unsigned)r * 16u + (unsigned)((c >> 2) & 1) * 8u;
            *reinterpret_cast<uint2*>(graw + o) = make_uint2(pack2(h1[0], h1[1]), pack2(h1[2], h1[3]));
            *reinterpret_cast<uint2*>(graw + graw_sz + o) = make_uint2(pack2(h2[0], h2[1]), pack2(h2[2], h2[3]));
            *reinterpret_cast<uint2*>(graw + 2 * graw_sz + o) = make_uint2(pack2(h3[0], h3[1]), pack2(h3[2], h3[3]));
          }
        }
      }
    }
    if (BWD) {
      // ---- S6 (MMA warp): d hid and d [hW; hb] from the g_raw tile just written
      operands_ready(4);
      if (has_next) {
        // next tile's inputs and hid behind this tile's gradient products; the MMA warp queues raw(i+1) after them
        put_inputs(b ^ 1);
        load_inputs(next + gridDim.x);
        worker_sync();
        build_hid(b ^ 1);
        operands_ready(2 + ((it + 1) & 1));
      } else {
        worker_sync();  // every octet has read its raw parameters before S7a overwrites them
      }
      if (!mbar_wait_bounded(smem_u32(&mbar[2]), (unsigned)it & 1u)) failed = true;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // ---- S7a: d pre-activation = d hid * (1 - hid^2) -> shared memory (over the raw parameters, no longer needed)
      {
        const int q = warp & 3, row = 16 * q + lane;
        const unsigned char* hb_ = hid + (unsigned)b * 3u * hid_sz;
        for (int c0 = 16 * (warp >> 2); c0 < Hp; c0 += 64) {
          float v1[16];
          tmem_ld16(tD2 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
          if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 16; i += 8) {
              const unsigned o = (unsigned)((c0 + i) >> 3) * CSB + (unsigned)row * 16u;
              const uint4 p1 = *reinterpret_cast<const uint4*>(hb_ + o);
              const uint4 p2 = *reinterpret_cast<const uint4*>(hb_ + hid_sz + o);
              const uint4 p3 = *reinterpret_cast<const uint4*>(hb_ + 2 * hid_sz + o);
              const unsigned w1[4] = {p1.x, p1.y, p1.z, p1.w}, w2[4] = {p2.x, p2.y, p2.z, p2.w},
                             w3[4] = {p3.x, p3.y, p3.z, p3.w};
              float d[8];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float ha = bf_lo(w1[t]) + bf_lo(w2[t]) + bf_lo(w3[t]);
                const float hb2 = bf_hi(w1[t]) + bf_hi(w2[t]) + bf_hi(w3[t]);
                const int j = c0 + i + 2 * t;
                d[2 * t] = j < H ? v1[i + 2 * t] * (1.f - ha * ha) : 0.f;
                d[2 * t + 1] = j + 1 < H ? v1[i + 2 * t + 1] * (1.f - hb2 * hb2) : 0.f;
              }
              *reinterpret_cast<float4*>(s_raw + row * LDS + c0 + i) = make_float4(d[0], d[1], d[2], d[3]);
              *reinterpret_cast<float4*>(s_raw + row * LDS + c0 + i + 4) = make_float4(d[4], d[5], d[6], d[7]);
            }
          }
        }
      }
      // ---- S7b: this tile's d [hW; hb] joins the register accumulators
      if (!mbar_wait_bounded(smem_u32(&mbar[3]), (unsigned)it & 1u)) failed = true;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      {
        const int q = warp & 3, sub = warp >> 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = 16 * (sub + 4 * h);
          if (c0 < RP) {
            float v1[16];
            tmem_ld16(tD3 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc_w[h][i] += v1[i];
          }
        }
      }
      worker_sync();
      // ---- S7c: d d1b / d d1W (column sums over the tile's rows: thread = (hidden unit, quarter of the rows), each
      // with its own CTA-lifetime accumulators), then the gradient wrt the conditioner columns (one octet per row) and g_nxt
      {
        const int j = tid & 127, part = tid >> 7;
        if (j < Hp) {
          float s = 0.f, sc[kMaxC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
          for (int r = part * (FM / 4); r < (part + 1) * (FM / 4); ++r) {
            const float d = s_raw[r * LDS + j];
            const float4 cn = *reinterpret_cast<const float4*>(s_c + (b * FM + r) * 4);
            s += d;
            sc[0] = fmaf(cn.x, d, sc[0]); sc[1] = fmaf(cn.y, d, sc[1]);
            sc[2] = fmaf(cn.z, d, sc[2]); sc[3] = fmaf(cn.w, d, sc[3]);
          }
          acc_b1 += s;
#pragma unroll
          for (int c = 0; c < kMaxC; ++c) acc_w1[c] += sc[c];
        }
      }
      {
        const int r = tid >> 3, l = tid & 7;
        float gc[kMaxC] = {0.f, 0.f, 0.f, 0.f};
        if (nc > 0) {
          const int j0 = l * (Hp / 8), j1 = j0 + Hp / 8;
          for (int j = j0; j < j1; ++j) {
            const float d = s_raw[r * LDS + j];
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
              if (c < nc) gc[c] = fmaf(d, s_w1[c * Hp + j], gc[c]);
          }
#pragma unroll
          for (int c = 0; c < kMaxC; ++c) {
            gc[c] += __shfl_xor_sync(0xffffffffu, gc[c], 1);
            gc[c] += __shfl_xor_sync(0xffffffffu, gc[c], 2);
            gc[c] += __shfl_xor_sync(0xffffffffu, gc[c], 4);
          }
        }
        if (l == 0 && r < nr) {
          float* gn = a.g_nxt + (row0 + r) * dz;
          const float* gcur = a.g_cur + (row0 + r) * dz;
          gn[a.ts0] = s_gin[r];
#pragma unroll
          for (int c = 0; c < kMaxC; ++c)
            if (c < nc) gn[a.cs0 + c] = __ldg(gcur + a.cs0 + c) + gc[c];
        }
      }
    }
    worker_sync();  // s_raw, s_gin and the input buffers of this tile are free again
  }

  if (BWD) {
    float* part = a.part + (size_t)blockIdx.x * a.part_stride;
    {
      // the four row-quarter accumulators of every hidden unit meet in shared memory (fixed order)
      float* red = s_raw;  // [4][Hp][5]
      const int j = tid & 127, pq = tid >> 7;
      if (j < Hp) {
        float* q = red + (pq * Hp + j) * 5;
        q[0] = acc_b1; q[1] = acc_w1[0]; q[2] = acc_w1[1]; q[3] = acc_w1[2]; q[4] = acc_w1[3];
      }
      worker_sync();
      if (tid < H) {
        float t[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) t[k] = ((red[(0 * Hp + tid) * 5 + k] + red[(1 * Hp + tid) * 5 + k]) +
                                            red[(2 * Hp + tid) * 5 + k]) + red[(3 * Hp + tid) * 5 + k];
        part[a.o_d1b + tid] = t[0];
#pragma unroll
        for (int c = 0; c < kMaxC; ++c)
          if (c < cin) part[a.o_d1W + (size_t)c * H + tid] = t[1 + c];
      }
    }
    const int jj = 32 * (warp & 3) + lane, sub = warp >> 2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = 16 * (sub + 4 * h) + i;
        if (n < R) {
          if (jj < H) part[a.o_hW + (size_t)jj * R + n] = acc_w[h][i];
          else if (jj == H) part[a.o_hb + n] = acc_w[h][i];
        }
      }
    }
  }
  if (failed && tid == 0 && a.err) atomicExch(a.err, 1);
  }  // workers
  tc_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

vms_status configure(const FlowTcArgs& a, bool bwd, KParams& p, int& RP, size_t& smem) {
  VMS_REQUIRE(flow_tc_supported(a.dz, a.nc > 0 ? a.nc : 1, 1, a.H, a.K), VMS_ERR_UNSUPPORTED,
              "flow_tc: unsupported block shape (dz=%d nc=%d H=%d K=%d)", a.dz, a.nc, a.H, a.K);
  VMS_REQUIRE(a.B >= 1 && a.uin && a.d1W && a.d1b && a.hW && a.hb, VMS_ERR_INVALID_ARG, "flow_tc: bad arguments");
  p.a = a;
  p.cin = a.nc > 0 ? a.nc : 1;
  p.R = 3 * a.K - 1;
  RP = p.R <= 64 ? 64 : 96;
  p.Hp = round_up(a.H + 1, 16);
  const int mx = p.Hp > RP ? p.Hp : RP;
  p.LDS = ((mx / 4) | 1) * 4;
  // flows.py:92: Python float arithmetic, then one cast to float32
  p.scale = (float)((double)a.bin_max - (double)a.bin_min - (double)a.K * 1e-2);
  int o = 0;  // bytes
  // d hW reads 16 chunk columns of the hid parts (M = 128): the over-read stays inside initialised shared memory
  p.o_hid = o; o += 2 * 3 * (p.Hp / 8) * (int)CSB;
  p.o_graw = o; o += bwd ? 3 * (RP / 8) * (int)CSB : 0;
  p.o_w = o; o += 3 * (p.Hp / 8) * RP * 16;
  o = round_up(o, 16);
  p.o_raw = o; o += FM * p.LDS * 4;
  p.o_w1 = o; o += 4 * p.Hp * 4;
  p.o_b1 = o; o += p.Hp * 4;
  p.o_c = o; o += 2 * FM * 4 * 4;
  p.o_v = o; o += 2 * FM * 4;
  p.o_g = o; o += 2 * FM * 4;
  p.o_lp = o; o += 2 * FM * 4;
  p.o_gin = o; o += FM * 4;
  smem = (size_t)o;
  VMS_REQUIRE(smem + 1024 <= (size_t)max_smem_optin(), VMS_ERR_UNSUPPORTED, "flow_tc: tile does not fit shared memory");
  return VMS_OK;
}

template <int RP, bool BWD>
vms_status launch(const KParams& p, size_t smem, cudaStream_t st) {
  VMS_CUDA(cudaFuncSetAttribute(flow_tc_kernel<RP, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  flow_tc_kernel<RP, BWD><<<flow_tc_grid(p.a.B), FTA, smem, st>>>(p);
  VMS_LAUNCH_CHECK("flow_tc_kernel");
  return VMS_OK;
}

}  // namespace

bool flow_tc_supported(int dz, int cin, int dt, int H, int K) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("VMS_FLOW_TC");
    disabled = (e && e[0] == '0') ? 1 : 0;
  }
  if (disabled) return false;
  if (dt != 1 || dz < 1 || cin < 1 || cin > kMaxC) return false;
  if (K < 4 || K > 32 || K % 4 != 0) return false;
  if (H < 8 || H + 1 > 112) return false;  // Hp <= 112: two accumulators of d hid fit the TMEM map, tiles fit 227 KB
  return true;
}

int flow_tc_grid(int64_t B) {
  const int64_t n_tiles = (B + FM - 1) / FM;
  return (int)(n_tiles < sm_count() ? n_tiles : sm_count());
}

vms_status flow_tc_forward(const FlowTcArgs& a, cudaStream_t st) {
  KParams p = {};
  int RP = 0;
  size_t smem = 0;
  vms_status s = configure(a, false, p, RP, smem);
  if (s) return s;
  VMS_REQUIRE(a.uout && a.logpz, VMS_ERR_INVALID_ARG, "flow_tc_forward: NULL output");
  return RP == 64 ? launch<64, false>(p, smem, st) : launch<96, false>(p, smem, st);
}

vms_status flow_tc_backward(const FlowTcArgs& a, cudaStream_t st) {
  KParams p = {};
  int RP = 0;
  size_t smem = 0;
  vms_status s = configure(a, true, p, RP, smem);
  if (s) return s;
  VMS_REQUIRE(a.g_cur && a.g_nxt && a.part, VMS_ERR_INVALID_ARG, "flow_tc_backward: NULL gradient buffer");
  return RP == 64 ? launch<64, true>(p, smem, st) : launch<96, true>(p, smem, st);
}

}  // namespace vms
