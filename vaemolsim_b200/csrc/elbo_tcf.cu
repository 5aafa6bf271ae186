// elbo_tcf.cu -- the whole ELBO training step of the C2 family as ONE persistent tensor-core kernel (plan mode 3; chosen
// automatically for forward + backward / train steps up to one wave of 32-row tiles).  It is the coupling-block pipeline
// of flow_tc.cu (3 x BF16 split, kind::f16 MMAs, one TMEM accumulator per product, K-major and MN-major reads of the same
// tiles, octet spline routines on the raw parameters) for all blocks of the chain back to back, with the encoder / decoder
// MLPs as in-tile FFMA phases, so that at the named batch (4096 rows = 128 tiles of 32 rows) the step is one launch whose
// heavy products run on the tensor core instead of the FFMA pipe of elbo_fused.cu.
//
// History.  Round 1 (one 64-row tile per CTA, every phase serial, weights staged by generic copies): correct, 0.187 ms
// against 0.146 ms of the FFMA kernel.  Round 2, ncu source view of the 32-row variant (profiles/r02_tcf_v1_hotspots.txt):
// 17 % of the samples in the scattered 4-byte stores of the d[hW; hb] partial and the barrier behind them, 11 % waiting
// for MMA completion, 20 % in MLP phases with run-time widths and per-phase weight staging from global memory, 5 % in the
// generic copy of the heads matrix.  This version: 32 valid rows per tile (the M = 64 products read whatever follows in
// shared memory for rows 32-63; every product is row-wise independent there and the d hW contraction covers rows 0-31
// only), the pre-split heads matrix of the NEXT block arrives by cp.async.bulk behind the current block's SIMT phases,
// encoder / decoder / conditioner weights are resident images brought in once by bulk copies, MLP phases are compiled
// for the widths of the shape, d[hW; hb] leaves the CTA as 16-byte stores into a padded partial layout, and the finish
// kernel that sums the partials and applies Adam also writes the NEXT step's pre-split / transposed weight images, so a
// training loop launches two kernels per step (the pre-pack kernel runs only when the parameters were changed behind the
// plan's back: first step, host-side weight assignment, data-parallel exchange).
//
// Reference lines replaced: the same as elbo.cu (models.py:289-322 VAE.call; mappings.py:107-155 FCDeepNN;
// flows.py:184-207, :281-355 RQSSplineRealNVP; dists.py:414-439; losses.py:58, :253) plus TF autodiff through them.
//
// Per tile (rows r = 0..63, all phases separated by CTA barriers):
//   E1  encoder  pe = relu(x W0 + b0) W1 + b1 (FFMA, thread = (row, 1/8 of the hidden units)), z = eps softplus + loc, log q
//   F   for block i = nb-1 .. 0 (chain inverse): stage the block's pre-split heads matrix; hid = tanh(cond d1W + d1b);
//       raw = [hid, 1] [hW; hb] (tcgen05); TMEM -> smem; spline inverse + log-det; chain state u[i] kept in smem
//   D1  decoder  pd = relu(z W0 + b0) W1 + b1, log p(x | z); the tile's loss terms
//   B1  decoder reverse mode: head, weight gradients (thread = (hidden unit, half of the rows)), d z
//   F'  for block i = 0 .. nb-1: recompute hid / raw, spline reverse mode, d hid = g_raw hW^T and d [hW; hb] =
//       [hid, 1]^T g_raw on the tensor core, epilogues (tanh', conditioner gradient, d d1W / d d1b)
//   E2  encoder head reverse mode and weight gradients
// Every CTA owns exactly one tile, so each weight gradient is written once into the CTA's slot of the partial buffer
// [n_tiles][P] (Keras layout) and a finish kernel sums the slots in a fixed order (+ Adam): deterministic.
//
// The helper functions up to issue_product are the ones of flow_tc.cu (kept in step by hand until both kernels settle).
#include "elbo_plan.cuh"
#include "flow_tc.cuh"
#include "peer.cuh"
#include "rqs_device.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace vms {

namespace {

constexpr int FM = 64;
constexpr int FT = 512;
constexpr unsigned CSB = (FM + 1) * 16;
constexpr int kMaxNb = 8;

struct TBlk {
  int cs0, nc, ts0, cin;
  int off_d1W, off_d1b, off_hW, off_hb;
};

struct TParams {
  int64_t B;
  int dx, dz, H, nb, K, fh, Hp, R, RP, LDS, P;
  int rows;  // valid rows per tile: 64, or 32 (the MMAs stay M = 64; rows beyond are zero operands)
  float bin_min, scale, klw;
  int enc0W, enc0b, enc1W, enc1b, dec0W, dec0b, dec1W, dec1b;
  int DIE, DOE, DID, DOD;  // padded MLP widths: round_up(Din + 1, 4), round_up(Dout, 4)
  TBlk blk[kMaxNb];
  const float *theta, *x, *eps;
  const unsigned short* wpk;  // pre-split heads matrices [nb][3 parts][Hp / 8][RP][8]
  const float* fpk;           // float images: encoder MLP | decoder MLP | conditioner first layers (see tcf_maps)
  int f_enc, f_dec, f_d1, n_enc, n_dec, n_d1;  // offsets / sizes in floats (sizes are multiples of 4)
  float *gpart, *spart;
  int P2;               // floats per CTA partial: theta layout + per block a padded [fh + 1][RP] matrix for d [hW; hb]
  int poff_h[kMaxNb];   // offset of that matrix in the partial
  int* err;
  int o_hid, o_graw, o_w, o_raw, o_u, o_gu, o_x, o_eps, o_pe, o_pd, o_gpd, o_gpe, o_gz, o_lp, o_gin, o_w1, o_b1, o_mlpe,
      o_mlpd;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void split3(float a, unsigned& h1, unsigned& h2, unsigned& h3) {
  h1 = __float_as_uint(a) & 0xffff0000u;
  const float r1 = a - __uint_as_float(h1);
  h2 = __float_as_uint(r1) & 0xffff0000u;
  const float r2 = r1 - __uint_as_float(h2);
  h3 = __float_as_uint(r2) & 0xffff0000u;
}
__device__ __forceinline__ unsigned pack2(unsigned e0, unsigned e1) { return __byte_perm(e0, e1, 0x7632); }
__device__ __forceinline__ float bf_lo(unsigned v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(unsigned v) { return __uint_as_float(v & 0xffff0000u); }

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
    if (clock64() - t0 > 2000000000LL) return false;
  }
}

__device__ __forceinline__ void tc_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// 1-D bulk asynchronous copies global -> shared (the TMA path without a tensor map), completion on an mbarrier
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ unsigned idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
         ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// see flow_tc.cu: one product from 3 x BF16 parts into ONE accumulator, small terms first
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

// ------------------------------------------------------------------------------------------------ packed weight images
// The kernel never reads theta directly: every weight arrives in shared memory by a bulk copy of an image laid out for
// its consumer.  pack_map[i] (built on the host, tcf_create) says where parameter i lives:
//   bit 31 clear: float image fpk[m]                 (MLP images: W0^T rows with the bias as column Din, W1 rows, b1;
//                                                      conditioner first layers [nb][4][Hp] and biases [nb][Hp])
//   bit 31 set:   heads image, part 1 at wpk[m & 0x7fffffff], parts 2 / 3 one / two `part` strides further
//                 ([blk][3 parts][j / 8][RP][8] bfloat16, the bias hb as row j = fh: the shared-memory layout of the MMA)
// Padding entries of the images are zeroed once at plan creation.  tcf_prepack_kernel rebuilds the images from theta; the
// finish kernel updates them in the same thread that applies Adam, so consecutive training steps never run the pre-pack.
__device__ __forceinline__ void pack_store(float w, unsigned m, float* __restrict__ fpk, unsigned short* __restrict__ wpk,
                                           size_t part) {
  if (m & 0x80000000u) {
    unsigned h1, h2, h3;
    split3(w, h1, h2, h3);
    unsigned short* dst = wpk + (m & 0x7fffffffu);
    dst[0] = (unsigned short)(h1 >> 16);
    dst[part] = (unsigned short)(h2 >> 16);
    dst[2 * part] = (unsigned short)(h3 >> 16);
  } else {
    fpk[m] = w;
  }
}

__global__ void __launch_bounds__(256) tcf_prepack_kernel(const float* __restrict__ theta, const unsigned* __restrict__ pack_map,
                                                          int P, float* __restrict__ fpk, unsigned short* __restrict__ wpk,
                                                          size_t part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) pack_store(__ldg(theta + i), __ldg(pack_map + i), fpk, wpk, part);
}

// ------------------------------------------------------------------------------------------------ MLP phases
// One two-layer MLP image in shared memory: w0t [H][DI] = W0^T with b0 in column Din (zero padded), w1 [H][DO] (zero
// padded), b1 [DO].  DI, DO are compile-time (multiples of 4): rows are read as 16-byte broadcasts.

// out [rows][DO] = relu(in W0 + b0) W1 + b1;  scratch: [FT / rows][rows][DO] floats
template <int DI, int DO>
__device__ void mlp_forward(const float* wm, int H, const float* in, int ldi, int Din, float* scratch, float* out, int rows) {
  const float* w0t = wm;
  const float* w1 = wm + H * DI;
  const float* b1 = w1 + H * DO;
  const int ng = FT / rows;  // thread = (row, one of ng groups of hidden units); a warp shares its group: broadcasts
  const int r = threadIdx.x & (rows - 1), g = threadIdx.x / rows;
  float x[DI], acc[DO];
#pragma unroll
  for (int i = 0; i < DI; ++i) x[i] = i < Din ? in[r * ldi + i] : (i == Din ? 1.f : 0.f);
#pragma unroll
  for (int n = 0; n < DO; ++n) acc[n] = 0.f;
#pragma unroll 2
  for (int j = g; j < H; j += ng) {
    float pre = 0.f;
#pragma unroll
    for (int q = 0; q < DI / 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(w0t + j * DI + 4 * q);
      pre = fmaf(x[4 * q], w.x, pre); pre = fmaf(x[4 * q + 1], w.y, pre);
      pre = fmaf(x[4 * q + 2], w.z, pre); pre = fmaf(x[4 * q + 3], w.w, pre);
    }
    const float h = fmaxf(pre, 0.f);
#pragma unroll
    for (int q = 0; q < DO / 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(w1 + j * DO + 4 * q);
      acc[4 * q] = fmaf(h, w.x, acc[4 * q]); acc[4 * q + 1] = fmaf(h, w.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(h, w.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(h, w.w, acc[4 * q + 3]);
    }
  }
#pragma unroll
  for (int q = 0; q < DO / 4; ++q)
    *reinterpret_cast<float4*>(scratch + (g * rows + r) * DO + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2],
                                                                                   acc[4 * q + 3]);
  __syncthreads();
  for (int e = threadIdx.x; e < rows * DO; e += FT) {
    const int r2 = e / DO, n = e - r2 * DO;
    float s = b1[n];
    for (int k = 0; k < ng; ++k) s += scratch[(k * rows + r2) * DO + n];
    out[e] = s;
  }
  __syncthreads();
}

// reverse mode: gout [rows][DO] (zero rows beyond the batch) -> this CTA's weight-gradient partial (global, Keras layout),
// optional input gradient gin [rows][4] (Din <= 4).  scratch: max(H * 24, (FT / rows) * rows * 4) floats
template <int DI, int DO>
__device__ void mlp_backward(const float* wm, int H, const float* in, int ldi, int Din, int Dout, const float* gout,
                             float* __restrict__ part, int oW0, int ob0, int oW1, int ob1, float* scratch, float* gin,
                             int rows) {
  const float* w0t = wm;
  const float* w1 = wm + H * DI;
  const int tid = threadIdx.x;
  {  // weight gradients: thread = (hidden unit j, half of the rows); inputs and output gradients are broadcasts
    const int j = tid & 255, half = tid >> 8;
    float w0[DI], w1r[DO], dW0[DI], dW1[DO];
#pragma unroll
    for (int i = 0; i < DI; ++i) { w0[i] = j < H ? w0t[j * DI + i] : 0.f; dW0[i] = 0.f; }
#pragma unroll
    for (int n = 0; n < DO; ++n) { w1r[n] = j < H ? w1[j * DO + n] : 0.f; dW1[n] = 0.f; }
    if (j < H) {
      for (int r = half * (rows / 2); r < (half + 1) * (rows / 2); ++r) {
        float x[DI], g[DO];
#pragma unroll
        for (int i = 0; i < DI; ++i) x[i] = i < Din ? in[r * ldi + i] : (i == Din ? 1.f : 0.f);
#pragma unroll
        for (int q = 0; q < DO / 4; ++q) {
          const float4 g4 = *reinterpret_cast<const float4*>(gout + r * DO + 4 * q);
          g[4 * q] = g4.x; g[4 * q + 1] = g4.y; g[4 * q + 2] = g4.z; g[4 * q + 3] = g4.w;
        }
        float pre = 0.f, t = 0.f;
#pragma unroll
        for (int i = 0; i < DI; ++i) pre = fmaf(x[i], w0[i], pre);
#pragma unroll
        for (int n = 0; n < DO; ++n) t = fmaf(g[n], w1r[n], t);
        const float h = fmaxf(pre, 0.f);
        const float gh = pre > 0.f ? t : 0.f;
#pragma unroll
        for (int n = 0; n < DO; ++n) dW1[n] = fmaf(h, g[n], dW1[n]);
#pragma unroll
        for (int i = 0; i < DI; ++i) dW0[i] = fmaf(x[i], gh, dW0[i]);
      }
      if (half == 1) {
#pragma unroll
        for (int i = 0; i < DI; ++i) scratch[j * 24 + i] = dW0[i];
#pragma unroll
        for (int n = 0; n < DO; ++n) scratch[j * 24 + 8 + n] = dW1[n];
      }
    }
    __syncthreads();
    if (j < H && half == 0) {
#pragma unroll
      for (int i = 0; i < DI; ++i) {
        const float v = dW0[i] + scratch[j * 24 + i];
        if (i < Din) part[oW0 + (size_t)i * H + j] = v;
        else if (i == Din) part[ob0 + j] = v;
      }
#pragma unroll
      for (int n = 0; n < DO; ++n)
        if (n < Dout) part[oW1 + (size_t)j * Dout + n] = dW1[n] + scratch[j * 24 + 8 + n];
    }
    if (tid >= FT - 16 && tid - (FT - 16) < Dout) {  // output bias: column sums of gout
      const int n = tid - (FT - 16);
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s += gout[r * DO + n];
      part[ob1 + n] = s;
    }
    __syncthreads();
  }
  if (gin) {  // input gradient: thread = (row, one of FT / rows groups of hidden units)
    const int ng = FT / rows;
    const int r = tid & (rows - 1), g8 = tid / rows;
    float x[DI], g[DO], gi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < DI; ++i) x[i] = i < Din ? in[r * ldi + i] : (i == Din ? 1.f : 0.f);
#pragma unroll
    for (int n = 0; n < DO; ++n) g[n] = gout[r * DO + n];
    for (int j = g8; j < H; j += ng) {
      float pre = 0.f, t = 0.f;
      float w0[DI];
#pragma unroll
      for (int q = 0; q < DI / 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(w0t + j * DI + 4 * q);
        w0[4 * q] = w.x; w0[4 * q + 1] = w.y; w0[4 * q + 2] = w.z; w0[4 * q + 3] = w.w;
      }
#pragma unroll
      for (int i = 0; i < DI; ++i) pre = fmaf(x[i], w0[i], pre);
#pragma unroll
      for (int q = 0; q < DO / 4; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(w1 + j * DO + 4 * q);
        t = fmaf(g[4 * q], w.x, t); t = fmaf(g[4 * q + 1], w.y, t); t = fmaf(g[4 * q + 2], w.z, t); t = fmaf(g[4 * q + 3], w.w, t);
      }
      const float gh = pre > 0.f ? t : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < Din && i < DI) gi[i] = fmaf(gh, w0[i], gi[i]);
    }
    *reinterpret_cast<float4*>(scratch + (g8 * rows + r) * 4) = make_float4(gi[0], gi[1], gi[2], gi[3]);
    __syncthreads();
    for (int e = tid; e < rows * 4; e += FT) {
      float s = 0.f;
      for (int k = 0; k < ng; ++k) s += scratch[k * rows * 4 + e];
      gin[e] = s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int RP, bool EXACT>
__global__ void __launch_bounds__(FT, 1) tcf_kernel(const __grid_constant__ TParams p) {
  // MLP image widths: the C2 shape (dx = 5..7, dz = 1..2) exactly, else padded to the maxima (zero columns)
  constexpr int DIE = 8, DOE = EXACT ? 4 : 16, DID = EXACT ? 4 : 8, DOD = EXACT ? 12 : 16;
  extern __shared__ __align__(128) unsigned char smb[];
  __shared__ __align__(8) unsigned long long mbars[3];  // [0] MMA completion, [1] heads-matrix copies, [2] resident images
  __shared__ unsigned tmem_base_s;
  __shared__ float red[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // programmatic dependent launch: the finish kernel of this step may be scheduled (and run its prologue: parameter, Adam
  // moment and map loads) while the tiles are still at work; it reads the partials only behind griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int H = p.fh, Hp = p.Hp, K = p.K, R = p.R, LDS = p.LDS, dz = p.dz, dx = p.dx, nb = p.nb;
  const int nchH = Hp / 8;
  constexpr int nchR = RP / 8;
  const unsigned hid_sz = (unsigned)nchH * CSB;
  constexpr unsigned graw_sz = (unsigned)nchR * CSB;
  const unsigned w_cs = RP * 16u;
  const unsigned w_sz = (unsigned)nchH * w_cs;
  unsigned char* hid = smb + p.o_hid;
  unsigned char* graw = smb + p.o_graw;
  unsigned char* wsm = smb + p.o_w;
  float* s_raw = reinterpret_cast<float*>(smb + p.o_raw);   // raw parameters / d pre-activation / MLP scratch
  const float* s_mlpe = reinterpret_cast<const float*>(smb + p.o_mlpe);  // encoder image (resident)
  const float* s_mlpd = reinterpret_cast<const float*>(smb + p.o_mlpd);  // decoder image (resident)
  float* s_u = reinterpret_cast<float*>(smb + p.o_u);       // [nb + 1][FM][4] chain states, u[nb] = z
  float* s_gu = reinterpret_cast<float*>(smb + p.o_gu);     // [FM][4] gradient wrt the current chain state
  float* s_x = reinterpret_cast<float*>(smb + p.o_x);       // [FM][8]
  float* s_eps = reinterpret_cast<float*>(smb + p.o_eps);   // [FM][4]
  float* s_pe = reinterpret_cast<float*>(smb + p.o_pe);     // [FM][DOE]
  float* s_pd = reinterpret_cast<float*>(smb + p.o_pd);     // [FM][DOD]
  float* s_gpd = reinterpret_cast<float*>(smb + p.o_gpd);   // [FM][DOD]
  float* s_gpe = reinterpret_cast<float*>(smb + p.o_gpe);   // [FM][DOE]
  float* s_gz = reinterpret_cast<float*>(smb + p.o_gz);     // [FM][4]
  float* s_lp = reinterpret_cast<float*>(smb + p.o_lp);     // [3][FM]: log q, log p(z), log p(x | z)
  float* s_gin = reinterpret_cast<float*>(smb + p.o_gin);   // [FM]
  const float* s_w1 = reinterpret_cast<const float*>(smb + p.o_w1);  // [nb][4][Hp]   (resident image, with s_b1 behind it)
  const float* s_b1 = reinterpret_cast<const float*>(smb + p.o_b1);  // [nb][Hp]
  const int rows = p.rows;  // valid rows of a tile (<= FM)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  const unsigned bar = smem_u32(&mbars[0]), wbar = smem_u32(&mbars[1]), sbar = smem_u32(&mbars[2]);
  const unsigned hid_a = smem_u32(hid), graw_a = smem_u32(graw), w_a = smem_u32(wsm);
  const size_t wpk_blk = 3 * (size_t)nchH * RP * 8;  // bfloat16 elements of one block's pre-split heads image
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbars[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // every element of the A-type tiles the tensor core can read must be finite: rows beyond the tile, the padding column
  // of g_raw (a NaN there would survive the multiplication by the zero padding row of hW)
  for (unsigned e = tid; e < 3 * hid_sz / 4; e += FT) reinterpret_cast<unsigned*>(hid)[e] = 0u;
  for (unsigned e = tid; e < 3 * graw_sz / 4; e += FT) reinterpret_cast<unsigned*>(graw)[e] = 0u;
  tc_sync();
  // ---- bulk copies: the resident images (MLPs, conditioner first layers) and the heads matrix of the first block
  auto load_heads = [&](int i) {  // thread 0 only; the previous reader of the buffer (an MMA) has been waited for
    mbar_expect_tx(wbar, 3u * w_sz);
    for (int q = 0; q < 3; ++q)
      bulk_g2s(w_a + q * w_sz, p.wpk + (size_t)i * wpk_blk + (size_t)q * (wpk_blk / 3), w_sz, wbar);
  };
  if (tid == 0) {
    mbar_expect_tx(sbar, 4u * (unsigned)(p.n_enc + p.n_dec + p.n_d1));
    bulk_g2s(smem_u32(smb + p.o_mlpe), p.fpk + p.f_enc, 4u * (unsigned)p.n_enc, sbar);
    bulk_g2s(smem_u32(smb + p.o_mlpd), p.fpk + p.f_dec, 4u * (unsigned)p.n_dec, sbar);
    bulk_g2s(smem_u32(smb + p.o_w1), p.fpk + p.f_d1, 4u * (unsigned)p.n_d1, sbar);
    load_heads(nb - 1);
  }
  const unsigned tm = tmem_base_s;
  const unsigned tD1 = tm, tD2 = tm + RP, tD3 = tm + RP + 128;
  const unsigned id1 = idesc_bf16(FM, RP, 0, 0);
  const unsigned id2 = idesc_bf16(FM, Hp, 0, 1);
  const unsigned id3 = idesc_bf16(128, RP, 1, 1);
  unsigned phase = 0, wphase = 0;
  bool failed = false;

  const int64_t tile = blockIdx.x;  // one tile per CTA
  const int64_t row0 = tile * rows;
  const int nr = (int)min((int64_t)rows, p.B - row0);
  float* part = p.gpart + (size_t)blockIdx.x * p.P2;
  const float invB = 1.0f / (float)p.B;
  const float g_logpx = -invB, g_logq = p.klw * invB, g_logpz = -p.klw * invB;

  // ---- hid of block i from the conditioner columns of chain state `uin` ([FM][4]); ones column at j = H
  auto build_hid = [&](int i, const float* uin) {
    const TBlk& fb = p.blk[i];
    const int r = tid & (rows - 1);
    float cnd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) cnd[c] = c < fb.nc ? uin[r * 4 + fb.cs0 + c] : ((fb.nc == 0 && c == 0) ? 1.f : 0.f);
    const float* w1 = s_w1 + i * 4 * Hp;
    const float* b1 = s_b1 + i * Hp;
    for (int jq = tid / rows; jq < 2 * nchH; jq += FT / rows) {
      const float4 b4 = *reinterpret_cast<const float4*>(b1 + 4 * jq);
      float pre[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < fb.cin) {
          const float4 w4 = *reinterpret_cast<const float4*>(w1 + c * Hp + 4 * jq);
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
      *reinterpret_cast<uint2*>(hid + o) = make_uint2(pack2(h1[0], h1[1]), pack2(h1[2], h1[3]));
      *reinterpret_cast<uint2*>(hid + hid_sz + o) = make_uint2(pack2(h2[0], h2[1]), pack2(h2[2], h2[3]));
      *reinterpret_cast<uint2*>(hid + 2 * hid_sz + o) = make_uint2(pack2(h3[0], h3[1]), pack2(h3[2], h3[3]));
    }
  };
  // ---- raw = [hid, 1] [hW; hb] on the tensor core, then TMEM -> s_raw
  // wait_w: the heads matrix of this block arrives by a bulk copy that has to be waited for;  prefetch >= 0: block whose
  // heads matrix is fetched into the (single) buffer as soon as this product has consumed the current one
  auto raw_product = [&](bool wait_w, int prefetch) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    if (wait_w) {
      if (!mbar_wait_bounded(wbar, wphase)) failed = true;
      wphase ^= 1;
    }
    tc_sync();
    if (tid == 0) {
      issue_product(tD1, hid_a, hid_sz, 2u * CSB, CSB, 128, w_a, w_sz, 2u * w_cs, w_cs, 128, Hp / 16, id1);
      mma_commit(bar);
    }
    if (!mbar_wait_bounded(bar, phase)) failed = true;
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0 && prefetch >= 0) load_heads(prefetch);
    const int q = warp & 3, row = 16 * q + lane;
    if (16 * q < rows)
      for (int c0 = 16 * (warp >> 2); c0 < RP; c0 += 64) {
        float v1[16];
        tmem_ld16(tD1 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(s_raw + row * LDS + c0 + i) = make_float4(v1[i], v1[i + 1], v1[i + 2], v1[i + 3]);
        }
      }
    tc_sync();
  };

  // ================================================================ T0: the tile's inputs
  for (int e = tid; e < rows * 8; e += FT) {
    const int r = e >> 3, c = e & 7;
    s_x[e] = (r < nr && c < dx) ? __ldg(p.x + (row0 + r) * dx + c) : 0.f;
  }
  for (int e = tid; e < rows * 4; e += FT) {
    const int r = e >> 2, c = e & 3;
    s_eps[e] = (r < nr && c < dz) ? __ldg(p.eps + (row0 + r) * dz + c) : 0.f;
  }
  // ================================================================ E1: encoder
  if (!mbar_wait_bounded(sbar, 0)) failed = true;  // resident images have landed
  __syncthreads();
  mlp_forward<DIE, DOE>(s_mlpe, p.H, s_x, 8, dx, s_raw, s_pe, rows);
  if (tid < rows) {
    float* z = s_u + (nb * FM + tid) * 4;
    float s = 0.f;
    for (int d = 0; d < 4; ++d) z[d] = 0.f;
    for (int d = 0; d < dz; ++d) {
      const float loc = s_pe[tid * DOE + d], sc = softplus_tf(s_pe[tid * DOE + dz + d]);
      const float zz = __fadd_rn(__fmul_rn(s_eps[tid * 4 + d], sc), loc);
      z[d] = zz;
      s += normal_lp(zz, loc, sc);
    }
    s_lp[tid] = s;
    s_lp[FM + tid] = 0.f;
  }
  __syncthreads();
  // ================================================================ F: chain inverse, block nb - 1 first
#pragma unroll 1
  for (int i = nb - 1; i >= 0; --i) {
    const TBlk& fb = p.blk[i];
    const float* uin = s_u + (i + 1) * FM * 4;
    float* uout = s_u + i * FM * 4;
    build_hid(i, uin);
    raw_product(true, i > 0 ? i - 1 : -1);  // (block 0's matrix stays: it is the first one the reverse pass needs)
    if ((tid >> 3) < rows) {  // whole warps (4 rows each)
      const int r = tid >> 3, j = tid & 7;
      const float* rr = s_raw + r * LDS;
      float out, ldj, ldj_all;
      bool writer;
      rqsdev::octet_apply<4, true, false>(rr, rr + K, rr + 2 * K, uin[r * 4 + fb.ts0], j, K, true, p.bin_min, p.scale, out,
                                          ldj, ldj_all, writer);
      if (writer) uout[r * 4 + fb.ts0] = out;
      if (j < fb.nc) uout[r * 4 + fb.cs0 + j] = uin[r * 4 + fb.cs0 + j];
      if (j == 0) s_lp[FM + r] += ldj_all;
    }
    __syncthreads();
  }
  if (tid < rows) {
    float s = s_lp[FM + tid];
    for (int d = 0; d < dz; ++d) s += normal_lp(s_u[tid * 4 + d], 0.f, 1.f);
    s_lp[FM + tid] = s;
  }
  // ================================================================ D1: decoder, loss terms
  __syncthreads();
  const float* zt = s_u + nb * FM * 4;
  mlp_forward<DID, DOD>(s_mlpd, p.H, zt, 4, dz, s_raw, s_pd, rows);
  red[0] = red[1] = red[2] = red[3] = 0.f;  // (same value from every thread; warps 0 / 1 overwrite their slots below)
  __syncthreads();
  if (tid < rows) {
    float s = 0.f;
    for (int d = 0; d < dx; ++d) s += normal_lp(s_x[tid * 8 + d], s_pd[tid * DOD + d], softplus_tf(s_pd[tid * DOD + dx + d]));
    s_lp[2 * FM + tid] = s;
    const bool ok = tid < nr;
    float kl = ok ? s_lp[tid] - s_lp[FM + tid] : 0.f, nll = ok ? -s : 0.f;
    kl = warp_sum(kl);
    nll = warp_sum(nll);
    if (lane == 0) { red[2 * warp] = kl; red[2 * warp + 1] = nll; }
    // decoder head, reverse mode (rows beyond the tile: zero gradient)
    for (int d = 0; d < DOD; ++d) s_gpd[tid * DOD + d] = 0.f;
    if (ok)
      for (int d = 0; d < dx; ++d) {
        const float loc = s_pd[tid * DOD + d], raw = s_pd[tid * DOD + dx + d];
        const float sc = softplus_tf(raw);
        const float u = s_x[tid * 8 + d] / sc - loc / sc;
        s_gpd[tid * DOD + d] = g_logpx * (u / sc);
        s_gpd[tid * DOD + dx + d] = g_logpx * ((u * u - 1.f) / sc) * sigmoidf_(raw);
      }
  }
  __syncthreads();
  if (tid == 0) {
    p.spart[2 * blockIdx.x] = red[0] + red[2];
    p.spart[2 * blockIdx.x + 1] = red[1] + red[3];
  }
  // ================================================================ B1: decoder reverse mode (weights still staged)
  mlp_backward<DID, DOD>(s_mlpd, p.H, zt, 4, dz, 2 * dx, s_gpd, part, p.dec0W, p.dec0b, p.dec1W, p.dec1b, s_raw, s_gz, rows);
  // ================================================================ F': flow reverse mode, block 0 first
  if (tid < rows)
    for (int d = 0; d < 4; ++d) s_gu[tid * 4 + d] = d < dz ? -g_logpz * s_u[tid * 4 + d] : 0.f;
  __syncthreads();
#pragma unroll 1
  for (int i = 0; i < nb; ++i) {
    const TBlk& fb = p.blk[i];
    const float* uin = s_u + (i + 1) * FM * 4;
    build_hid(i, uin);
    raw_product(i > 0, -1);
    if ((tid >> 3) < rows) {  // spline reverse mode, one octet per row (whole warps)
      const int r = tid >> 3, j = tid & 7;
      const bool ok = r < nr;
      const float* rr = s_raw + r * LDS;
      float g_in, gw[4], gh[4], gs[4];
      bool writer;
      rqsdev::octet_backward<4, true, false>(rr, rr + K, rr + 2 * K, uin[r * 4 + fb.ts0], ok ? s_gu[r * 4 + fb.ts0] : 0.f,
                                             ok ? g_logpz : 0.f, j, K, true, p.bin_min, p.scale, g_in, writer, gw, gh, gs);
      if (writer) s_gin[r] = g_in;
      if (4 * j < K) {
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
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_sync();
    if (tid == 0) {
      issue_product(tD2, graw_a, graw_sz, 2u * CSB, CSB, 128, w_a, w_sz, 256u, 128, w_cs, RP / 16, id2);
      issue_product(tD3, hid_a, hid_sz, 256u, 128, CSB, graw_a, graw_sz, 256u, 128, CSB, rows / 16, id3);
      mma_commit(bar);
    }
    if (!mbar_wait_bounded(bar, phase)) failed = true;
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (tid == 0 && i + 1 < nb) load_heads(i + 1);  // both products have consumed this block's matrix
    if (16 * (warp & 3) < rows) {  // d pre-activation = d hid * (1 - hid^2) -> s_raw
      const int q = warp & 3, row = 16 * q + lane;
      for (int c0 = 16 * (warp >> 2); c0 < Hp; c0 += 64) {
        float v1[16];
        tmem_ld16(tD2 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
        if (lane < 16) {
#pragma unroll
          for (int k = 0; k < 16; k += 8) {
            const unsigned o = (unsigned)((c0 + k) >> 3) * CSB + (unsigned)row * 16u;
            const uint4 p1 = *reinterpret_cast<const uint4*>(hid + o);
            const uint4 p2 = *reinterpret_cast<const uint4*>(hid + hid_sz + o);
            const uint4 p3 = *reinterpret_cast<const uint4*>(hid + 2 * hid_sz + o);
            const unsigned w1[4] = {p1.x, p1.y, p1.z, p1.w}, w2[4] = {p2.x, p2.y, p2.z, p2.w},
                           w3[4] = {p3.x, p3.y, p3.z, p3.w};
            float d[8];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float ha = bf_lo(w1[t]) + bf_lo(w2[t]) + bf_lo(w3[t]);
              const float hb2 = bf_hi(w1[t]) + bf_hi(w2[t]) + bf_hi(w3[t]);
              const int j = c0 + k + 2 * t;
              d[2 * t] = j < H ? v1[k + 2 * t] * (1.f - ha * ha) : 0.f;
              d[2 * t + 1] = j + 1 < H ? v1[k + 2 * t + 1] * (1.f - hb2 * hb2) : 0.f;
            }
            *reinterpret_cast<float4*>(s_raw + row * LDS + c0 + k) = make_float4(d[0], d[1], d[2], d[3]);
            *reinterpret_cast<float4*>(s_raw + row * LDS + c0 + k + 4) = make_float4(d[4], d[5], d[6], d[7]);
          }
        }
      }
    }
    {  // d [hW; hb] of this tile -> the CTA's partial, a padded [fh + 1][RP] matrix (rows 16-byte aligned): thread =
       // hidden unit 32 q + lane, 16-column chunks sub, sub + 4, four 16-byte stores per chunk
      const int q = warp & 3, sub = warp >> 2, jj = 32 * q + lane;
      float* dst = part + p.poff_h[i] + (size_t)jj * RP;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = 16 * (sub + 4 * h);
        if (c0 < RP) {
          float v1[16];
          tmem_ld16(tD3 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
          if (jj <= H) {
#pragma unroll
            for (int k = 0; k < 16; k += 4)
              *reinterpret_cast<float4*>(dst + c0 + k) = make_float4(v1[k], v1[k + 1], v1[k + 2], v1[k + 3]);
          }
        }
      }
    }
    tc_sync();
    float cs_b = 0.f, cs_w[4] = {0.f, 0.f, 0.f, 0.f};
    {  // d d1b / d d1W: thread = (hidden unit, quarter of the rows)
      const int j = tid & 127, pq = tid >> 7;
      if (j < Hp) {
        for (int r = pq * (rows / 4); r < (pq + 1) * (rows / 4); ++r) {
          const float d = s_raw[r * LDS + j];
          cs_b += d;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float cv = c < fb.nc ? uin[r * 4 + fb.cs0 + c] : ((fb.nc == 0 && c == 0) ? 1.f : 0.f);
            cs_w[c] = fmaf(cv, d, cs_w[c]);
          }
        }
      }
    }
    {  // gradient wrt the conditioner columns (one octet per row), chain-state gradient update
      const int r = tid >> 3, l = tid & 7;
      float gc[4] = {0.f, 0.f, 0.f, 0.f};
      if (fb.nc > 0 && r < rows) {  // (r < rows is warp-uniform: 4 rows per warp)
        const float* w1 = s_w1 + i * 4 * Hp;
        const int j0 = l * (Hp / 8), j1 = j0 + Hp / 8;
        for (int j = j0; j < j1; ++j) {
          const float d = s_raw[r * LDS + j];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < fb.nc) gc[c] = fmaf(d, w1[c * Hp + j], gc[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          gc[c] += __shfl_xor_sync(0xffffffffu, gc[c], 1);
          gc[c] += __shfl_xor_sync(0xffffffffu, gc[c], 2);
          gc[c] += __shfl_xor_sync(0xffffffffu, gc[c], 4);
        }
      }
      if (l == 0 && r < rows) {
        s_gu[r * 4 + fb.ts0] = s_gin[r];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < fb.nc) s_gu[r * 4 + fb.cs0 + c] += gc[c];
      }
    }
    __syncthreads();  // s_raw (d pre-activation) fully consumed
    {
      float* q5 = s_raw;  // [4][Hp][5]
      const int j = tid & 127, pq = tid >> 7;
      if (j < Hp) {
        float* q = q5 + (pq * Hp + j) * 5;
        q[0] = cs_b; q[1] = cs_w[0]; q[2] = cs_w[1]; q[3] = cs_w[2]; q[4] = cs_w[3];
      }
      __syncthreads();
      if (tid < H) {
        float t[5];
#pragma unroll
        for (int k = 0; k < 5; ++k)
          t[k] = ((q5[(0 * Hp + tid) * 5 + k] + q5[(1 * Hp + tid) * 5 + k]) + q5[(2 * Hp + tid) * 5 + k]) +
                 q5[(3 * Hp + tid) * 5 + k];
        part[fb.off_d1b + tid] = t[0];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < fb.cin) part[fb.off_d1W + (size_t)c * H + tid] = t[1 + c];
      }
      __syncthreads();
    }
  }
  // ================================================================ E2: encoder head and encoder reverse mode
  if (tid < rows) {
    const bool ok = tid < nr;
    const float* z = s_u + (nb * FM + tid) * 4;
    for (int d = 0; d < DOE; ++d) s_gpe[tid * DOE + d] = 0.f;
    if (ok)
      for (int d = 0; d < dz; ++d) {
        const float loc = s_pe[tid * DOE + d], raw = s_pe[tid * DOE + dz + d];
        const float sc = softplus_tf(raw);
        const float u = z[d] / sc - loc / sc;
        const float gzt = (s_gz[tid * 4 + d] + (nb > 0 ? s_gu[tid * 4 + d] : -g_logpz * z[d])) + g_logq * (-u / sc);
        s_gpe[tid * DOE + d] = g_logq * (u / sc) + gzt;
        s_gpe[tid * DOE + dz + d] = (g_logq * ((u * u - 1.f) / sc) + gzt * s_eps[tid * 4 + d]) * sigmoidf_(raw);
      }
  }
  __syncthreads();
  mlp_backward<DIE, DOE>(s_mlpe, p.H, s_x, 8, dx, 2 * dz, s_gpe, part, p.enc0W, p.enc0b, p.enc1W, p.enc1b, s_raw, nullptr, rows);

  if (failed && tid == 0 && p.err) atomicExch(p.err, 1);
  tc_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}

// grad[i] = sum over the tiles' partials in a fixed order (+ Keras Adam + the next step's weight images), and the three
// loss scalars.  Block = 4 groups x 128 parameters: group g sums its quarter of the partials with eight independent
// running sums (all loads of a thread in flight at once), the four group sums meet in shared memory in a fixed order.
constexpr int kFinScal = 1024;  // shared-memory slots for the tiles' scalar partials (2 per tile: up to 512 tiles)
// {loss, nll, kl} from the tiles' partial sums, added in tile order (deterministic).  Call after the barrier that follows the
// staging of `ssc` (block 0); the last thread of the block adds them from shared memory.  Batches of more than 512 tiles
// (never staged) fall back to thread 0 reading global memory.
__device__ __forceinline__ void finish_scalars(bool staged, const float* ssc, const float* __restrict__ spart, int n_part,
                                               int64_t B, float klw, float* __restrict__ scalars) {
  if (blockIdx.x != 0 || scalars == nullptr) return;
  const float* src = staged ? ssc : spart;
  if (threadIdx.x != (staged ? blockDim.x - 1 : 0u)) return;
  float sa = 0.f, sc = 0.f;
  for (int c = 0; c < n_part; ++c) { sa += src[2 * c]; sc += src[2 * c + 1]; }
  const float kl = sa / (float)B, nll = sc / (float)B;
  scalars[0] = nll + klw * kl;
  scalars[1] = nll;
  scalars[2] = kl;
}

struct TcfAdam {
  float *theta, *m, *v;
  float lr_t, one_minus_b1, one_minus_b2, eps;
};
constexpr int kFinP = 128, kFinG = 4;
__global__ void __launch_bounds__(kFinP * kFinG) tcf_finish_kernel(const float* __restrict__ gpart, int n_part, int P, int P2,
                                                                   const int* __restrict__ part_map, float* __restrict__ grad,
                                                                   const float* __restrict__ spart, int64_t B, float klw,
                                                                   float* __restrict__ scalars, const TcfAdam ad,
                                                                   const unsigned* __restrict__ pack_map, float* __restrict__ fpk,
                                                                   unsigned short* __restrict__ wpk, size_t part) {
  __shared__ float sh[kFinG][kFinP];
  __shared__ float ssc[kFinScal];
  const int tx = threadIdx.x & (kFinP - 1), g = threadIdx.x / kFinP;
  const int i = blockIdx.x * kFinP + tx;
  const bool live = i < P;
  // block 0 also owns the three loss scalars: the tiles' partial sums are fetched by the whole block up front and added in
  // tile order by its LAST thread after the first barrier (one thread fetching and adding 2 x 128 values after its own
  // Adam update was the tail of the kernel)
  const bool scal = blockIdx.x == 0 && scalars != nullptr && 2 * n_part <= kFinScal;
  float th = 0.f, mi = 0.f, vi = 0.f;
  unsigned pm = 0u;
  if (live && g == 0 && ad.theta) { th = ad.theta[i]; mi = ad.m[i]; vi = ad.v[i]; pm = __ldg(pack_map + i); }  // in flight beside the partial loads
  const int poff = live ? __ldg(part_map + i) : 0;
  // everything above is independent of the tile kernel; its partials are read behind this (programmatic dependent launch;
  // returns at once when the kernel was launched the ordinary way)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (scal)
    for (int c = threadIdx.x; c < 2 * n_part; c += kFinP * kFinG) ssc[c] = spart[c];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live) {
    const int per = (n_part + kFinG - 1) / kFinG;
    const int c0 = g * per, c1 = min(n_part, c0 + per);
    const float* src = gpart + poff;
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += src[(size_t)(c + u) * P2];
    }
    for (; c < c1; ++c) acc[0] += src[(size_t)c * P2];
  }
  sh[g][tx] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (g == 0 && live) {
    const float t = (sh[0][tx] + sh[1][tx]) + (sh[2][tx] + sh[3][tx]);
    grad[i] = t;
    if (ad.theta) {  // Keras Adam on the flat buffer, same arithmetic as adam_kernel (adam.cu)
      mi = mi + (t - mi) * ad.one_minus_b1;
      vi = vi + (t * t - vi) * ad.one_minus_b2;
      ad.m[i] = mi;
      ad.v[i] = vi;
      th = th - ad.lr_t * mi / (sqrtf(vi) + ad.eps);
      ad.theta[i] = th;
      pack_store(th, pm, fpk, wpk, part);  // the next step's images, written by the optimiser itself
    }
  }
  finish_scalars(scal, ssc, spart, n_part, B, klw, scalars);
}

// Data-parallel step: the finish kernel and the NVLink exchange in ONE launch.  Phase 1 (never blocks): the tile partials are
// summed exactly as above and this rank's gradient lands in its slot of the peer buffer; the block that finishes last raises
// this rank's flag in every peer's buffer.  Phase 2: block 0 waits (bounded) for every rank's flag and decides for the grid,
// the other blocks wait for its decision.  Phase 3: every parameter's gradient is pulled from all ranks' slots over NVLink,
// summed in rank order, scaled, and the Adam update writes theta AND the next step's weight images -- a data-parallel step is
// two launches (tile kernel + this one), like the single-GPU step.  The whole grid must be co-resident (phase 2 waits for
// blocks of phase 1): the launcher checks it and otherwise keeps the separate finish / exchange kernels.
__global__ void __launch_bounds__(kFinP * kFinG) tcf_finish_peer_kernel(const float* __restrict__ gpart, int n_part, int P, int P2,
                                                                        const int* __restrict__ part_map,
                                                                        const float* __restrict__ spart, int64_t B, float klw,
                                                                        float* __restrict__ scalars, const PeerArgs a,
                                                                        const unsigned* __restrict__ pack_map,
                                                                        float* __restrict__ fpk, unsigned short* __restrict__ wpk,
                                                                        size_t part) {
  __shared__ float sh[kFinG][kFinP];
  __shared__ float ssc[kFinScal];
  __shared__ int failed;
  const int tx = threadIdx.x & (kFinP - 1), g = threadIdx.x / kFinP;
  const int i = blockIdx.x * kFinP + tx;
  const bool live = i < P;
  const bool scal = blockIdx.x == 0 && scalars != nullptr && 2 * n_part <= kFinScal;  // see tcf_finish_kernel
  // timeline of the last step in flag slots 56 .. 59 (nanoseconds of %globaltimer: entry of block 0, gradient complete, exchange
  // decided, block 0 done) -- read by scripts/time_dp_step.py
  unsigned long long* trace = flags_of(a.base[a.rank], a.P) + 56;
  if (blockIdx.x == 0 && threadIdx.x == 0) trace[0] = globaltimer_ns();
  float th = 0.f, mi = 0.f, vi = 0.f;
  unsigned pm = 0u;
  if (live && g == 0) { th = a.theta[i]; mi = a.m[i]; vi = a.v[i]; pm = __ldg(pack_map + i); }
  const int poff = live ? __ldg(part_map + i) : 0;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // see tcf_finish_kernel
  if (scal)
    for (int c = threadIdx.x; c < 2 * n_part; c += kFinP * kFinG) ssc[c] = spart[c];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (live) {
    const int per = (n_part + kFinG - 1) / kFinG;
    const int c0 = g * per, c1 = min(n_part, c0 + per);
    const float* src = gpart + poff;
    int c = c0;
    for (; c + 8 <= c1; c += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += src[(size_t)(c + u) * P2];
    }
    for (; c < c1; ++c) acc[0] += src[(size_t)c * P2];
  }
  sh[g][tx] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  __syncthreads();
  if (g == 0 && live) {
    const float t = (sh[0][tx] + sh[1][tx]) + (sh[2][tx] + sh[3][tx]);
    a.base[a.rank][(int64_t)(a.step & 1ull) * a.P + i] = t;  // my slot of this step
  }
  finish_scalars(scal, ssc, spart, n_part, B, klw, scalars);
  __syncthreads();
  if (threadIdx.x == 0) {
    // release pattern: the block's slot stores happen before this barrier; ONE device-scope fence by thread 0 (fences are
    // cumulative) orders them before the counter increment, and the block that observes the final count issues the
    // system-scope fence before it raises the flags (a fence.sys per writing thread cost ~10 us per step)
    unsigned long long* counter = flags_of(a.base[a.rank], a.P) + 41;
    __threadfence();
    const unsigned long long done = atomicAdd(counter, 1ull) + 1ull;
    if (done == (unsigned long long)gridDim.x) {  // the whole gradient of this rank is in its slot
      *counter = 0ull;
      trace[1] = globaltimer_ns();
      peer_raise_flags(a);
    }
    failed = blockIdx.x == 0 ? peer_wait_and_decide(a) : peer_wait_decision(a);
    if (blockIdx.x == 0) trace[2] = globaltimer_ns();
  }
  __syncthreads();
  if (failed || !(g == 0 && live)) return;
  const float gsum = peer_pull_sum(a, i);
  if (a.grad_out) a.grad_out[i] = gsum;
  mi = mi + (gsum - mi) * a.one_minus_b1;
  vi = vi + (gsum * gsum - vi) * a.one_minus_b2;
  a.m[i] = mi;
  a.v[i] = vi;
  th = th - a.lr_t * mi / (sqrtf(vi) + a.eps);
  a.theta[i] = th;
  pack_store(th, pm, fpk, wpk, part);
  if (blockIdx.x == 0 && threadIdx.x == 0) trace[3] = globaltimer_ns();
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct TcfCfg {
  TParams p;
  size_t smem;
  int max_tiles;
  int peer_ok = -1;  // fused finish + exchange kernel usable (grid co-resident): -1 = not probed yet
  bool exact;       // MLP images in the C2 widths (DI, DO) = (8, 4) / (4, 12); else padded to (8, 16) / (8, 16)
  float *gpart, *spart, *fpk;
  unsigned short* wpk;
  unsigned* pack_map;  // [P] device
  int* part_map;       // [P] device
  size_t part;         // bfloat16 elements between the three parts of a block's heads image
  size_t fpk_floats, wpk_elems;
  // the images are valid for the parameters at `pack_theta` as this plan last left them (a fused Adam step wrote them)
  bool pack_valid = false;
  const float* pack_theta = nullptr;
  // optional per-launch timing of the main kernel (bench.py's roofline leg)
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t ev_used = 0;
};

// Sets pl->tcf when the shape is supported (same block shapes as flow_tc, MLP widths dx <= 7, dz <= 3, hidden <= 240).
vms_status tcf_create(vms_elbo_plan_s* pl) {
  pl->tcf = nullptr;
  const vms_elbo_desc& d = pl->d;
  if (d.num_blocks < 1 || d.num_blocks > kMaxNb || !pl->tc_ok) return VMS_OK;
  if (d.dx > 7 || d.dz > 3 || d.hidden > 240 || 2 * d.dx > 16 || 2 * d.dz > 8) return VMS_OK;
  TcfCfg* f = new TcfCfg();
  TParams& p = f->p;
  memset(&p, 0, sizeof(p));
  p.dx = d.dx; p.dz = d.dz; p.H = d.hidden; p.nb = d.num_blocks; p.K = d.num_bins; p.fh = d.flow_hidden;
  p.R = 3 * d.num_bins - 1;
  p.RP = p.R <= 64 ? 64 : 96;
  p.Hp = round_up(d.flow_hidden + 1, 16);
  const int mx = p.Hp > p.RP ? p.Hp : p.RP;
  p.LDS = ((mx / 4) | 1) * 4;
  p.bin_min = d.bin_min;
  p.scale = (float)((double)d.bin_max - (double)d.bin_min - (double)d.num_bins * 1e-2);
  p.klw = d.kl_weight;
  const Offsets& o = pl->off;
  p.P = (int)o.total;
  p.enc0W = (int)o.enc0W; p.enc0b = (int)o.enc0b; p.enc1W = (int)o.enc1W; p.enc1b = (int)o.enc1b;
  p.dec0W = (int)o.dec0W; p.dec0b = (int)o.dec0b; p.dec1W = (int)o.dec1W; p.dec1b = (int)o.dec1b;
  f->exact = round_up(d.dx + 1, 4) == 8 && round_up(2 * d.dz, 4) == 4 && round_up(d.dz + 1, 4) == 4 && round_up(2 * d.dx, 4) == 12;
  p.DIE = 8; p.DOE = f->exact ? 4 : 16;
  p.DID = f->exact ? 4 : 8; p.DOD = f->exact ? 12 : 16;
  for (int i = 0; i < d.num_blocks; ++i) {
    const FlowBlock& b = pl->blocks[i];
    TBlk& t = p.blk[i];
    t.cs0 = b.cs0; t.nc = b.cs1 - b.cs0; t.ts0 = b.ts0; t.cin = b.cin;
    t.off_d1W = (int)b.off_d1W; t.off_d1b = (int)b.off_d1b; t.off_hW = (int)b.off_hW; t.off_hb = (int)b.off_hb;
  }
  const int nchH = p.Hp / 8, nchR = p.RP / 8, H = d.hidden, fh = d.flow_hidden, nb = d.num_blocks;
  // ---- float images and the maps parameter -> image / parameter -> partial
  p.n_enc = round_up(H * (p.DIE + p.DOE) + p.DOE, 4);
  p.n_dec = round_up(H * (p.DID + p.DOD) + p.DOD, 4);
  p.n_d1 = nb * 5 * p.Hp;
  p.f_enc = 0; p.f_dec = p.n_enc; p.f_d1 = p.n_enc + p.n_dec;
  f->fpk_floats = (size_t)p.n_enc + p.n_dec + p.n_d1;
  f->part = (size_t)nchH * p.RP * 8;
  f->wpk_elems = (size_t)nb * 3 * f->part;
  std::vector<unsigned> pmap((size_t)p.P, 0u);
  std::vector<int> qmap((size_t)p.P, 0);
  for (int i = 0; i < p.P; ++i) qmap[i] = i;
  auto mlp_map = [&](int base, int oW0, int ob0, int oW1, int ob1, int Din, int Dout, int DI, int DO) {
    for (int i = 0; i < Din; ++i)
      for (int j = 0; j < H; ++j) pmap[oW0 + (size_t)i * H + j] = (unsigned)(base + j * DI + i);
    for (int j = 0; j < H; ++j) pmap[ob0 + j] = (unsigned)(base + j * DI + Din);
    for (int j = 0; j < H; ++j)
      for (int n = 0; n < Dout; ++n) pmap[oW1 + (size_t)j * Dout + n] = (unsigned)(base + H * DI + j * DO + n);
    for (int n = 0; n < Dout; ++n) pmap[ob1 + n] = (unsigned)(base + H * DI + H * DO + n);
  };
  mlp_map(p.f_enc, p.enc0W, p.enc0b, p.enc1W, p.enc1b, d.dx, 2 * d.dz, p.DIE, p.DOE);
  mlp_map(p.f_dec, p.dec0W, p.dec0b, p.dec1W, p.dec1b, d.dz, 2 * d.dx, p.DID, p.DOD);
  int P2 = round_up(p.P, 4);
  for (int b = 0; b < nb; ++b) {
    const TBlk& t = p.blk[b];
    for (int c = 0; c < t.cin; ++c)
      for (int j = 0; j < fh; ++j) pmap[t.off_d1W + (size_t)c * fh + j] = (unsigned)(p.f_d1 + (b * 4 + c) * p.Hp + j);
    for (int j = 0; j < fh; ++j) pmap[t.off_d1b + j] = (unsigned)(p.f_d1 + nb * 4 * p.Hp + b * p.Hp + j);
    p.poff_h[b] = P2;
    for (int j = 0; j <= fh; ++j)
      for (int n = 0; n < p.R; ++n) {
        const size_t th = j < fh ? (size_t)t.off_hW + (size_t)j * p.R + n : (size_t)t.off_hb + n;
        pmap[th] = 0x80000000u | (unsigned)((size_t)b * 3 * f->part + (size_t)(j >> 3) * p.RP * 8 + (size_t)n * 8 + (j & 7));
        qmap[th] = P2 + j * p.RP + n;
      }
    P2 += (fh + 1) * p.RP;
  }
  p.P2 = P2;
  // ---- shared-memory carve-up (bytes)
  int off = 0;
  auto take = [&](int bytes) { int o0 = off; off += round_up(bytes, 16); return o0; };
  p.o_hid = take(3 * nchH * (int)CSB);
  p.o_graw = take(3 * nchR * (int)CSB);
  p.o_w = take(3 * nchH * p.RP * 16);
  // s_raw doubles as scratch of the MLP phases: 16 row groups x 32 rows x 16 outputs, or hidden x 24 accumulators
  int raw_floats = FM * p.LDS;
  if (raw_floats < 8 * FM * 16) raw_floats = 8 * FM * 16;
  if (raw_floats < d.hidden * 24) raw_floats = d.hidden * 24;
  p.o_raw = take(4 * raw_floats);
  p.o_u = take((d.num_blocks + 1) * FM * 4 * 4);
  p.o_gu = take(FM * 4 * 4);
  p.o_x = take(FM * 8 * 4); p.o_eps = take(FM * 4 * 4);
  p.o_pe = take(FM * p.DOE * 4); p.o_pd = take(FM * p.DOD * 4);
  p.o_gpd = take(FM * p.DOD * 4); p.o_gpe = take(FM * p.DOE * 4);
  p.o_gz = take(FM * 4 * 4); p.o_lp = take(3 * FM * 4); p.o_gin = take(FM * 4);
  p.o_w1 = take(p.n_d1 * 4);  // [nb][4][Hp] followed by [nb][Hp]: one bulk copy
  p.o_b1 = p.o_w1 + nb * 4 * p.Hp * 4;
  p.o_mlpe = take(p.n_enc * 4);
  p.o_mlpd = take(p.n_dec * 4);
  f->smem = (size_t)off;
  if (f->smem + 1024 > (size_t)max_smem_optin()) { delete f; return VMS_OK; }
  // up to THREE waves of one-tile CTAs (14,208 rows on 148 SMs): measured against the per-block tensor-core plan, 0.18 vs
  // 0.27 ms in the second wave, 0.28 vs 0.34 ms in the third; the partial-gradient buffer is sized for the plan's max_batch
  {
    const int64_t want = ((int64_t)pl->maxB + 31) / 32;
    const int64_t cap = 3 * (int64_t)sm_count();
    f->max_tiles = (int)(want < cap ? (want > sm_count() ? want : sm_count()) : cap);
  }
  void *g = nullptr, *s = nullptr, *w = nullptr, *fp = nullptr, *pm = nullptr, *qm = nullptr;
  bool ok = cudaMalloc(&g, (size_t)f->max_tiles * p.P2 * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&s, (size_t)f->max_tiles * 2 * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&w, f->wpk_elems * sizeof(unsigned short)) == cudaSuccess &&
            cudaMalloc(&fp, f->fpk_floats * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&pm, (size_t)p.P * sizeof(unsigned)) == cudaSuccess &&
            cudaMalloc(&qm, (size_t)p.P * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMemset(w, 0, f->wpk_elems * sizeof(unsigned short)) == cudaSuccess &&
       cudaMemset(fp, 0, f->fpk_floats * sizeof(float)) == cudaSuccess &&
       cudaMemcpy(pm, pmap.data(), (size_t)p.P * sizeof(unsigned), cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(qm, qmap.data(), (size_t)p.P * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
  // the conditioner image's ones-column convention: an empty conditioner (cin = 1, nc = 0) multiplies d1W row 0 by 1
  if (!ok) {
    cudaGetLastError();
    for (void* q : {g, s, w, fp, pm, qm})
      if (q) cudaFree(q);
    delete f;
    return VMS_OK;  // this path is simply unavailable; the FFMA fused kernel serves the shape
  }
  f->gpart = (float*)g; f->spart = (float*)s; f->wpk = (unsigned short*)w; f->fpk = (float*)fp;
  f->pack_map = (unsigned*)pm; f->part_map = (int*)qm;
  pl->tcf = f;
  return VMS_OK;
}

void tcf_destroy(vms_elbo_plan_s* pl) {
  if (!pl->tcf) return;
  for (auto& e : pl->tcf->ev) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  cudaFree(pl->tcf->gpart);
  cudaFree(pl->tcf->spart);
  cudaFree(pl->tcf->wpk);
  cudaFree(pl->tcf->fpk);
  cudaFree(pl->tcf->pack_map);
  cudaFree(pl->tcf->part_map);
  delete pl->tcf;
  pl->tcf = nullptr;
}

// valid rows per tile: 32 (default: 128 CTAs at batch 4096), or 64 with VMS_TCF_ROWS=64 (round 1's configuration)
static int tcf_rows() {
  const char* e = getenv("VMS_TCF_ROWS");
  return (e && atoi(e) == 64) ? FM : 32;
}

bool tcf_available(const vms_elbo_plan_s* pl, int64_t B) {
  const int rows = tcf_rows();
  return pl->tcf && (B + rows - 1) / rows <= pl->tcf->max_tiles;
}

void tcf_invalidate(vms_elbo_plan_s* pl) {
  if (pl->tcf) pl->tcf->pack_valid = false;
}

vms_status tcf_set_timing(vms_elbo_plan_s* pl, int max_launches) {
  TcfCfg* f = pl->tcf;
  if (!f) return VMS_OK;
  f->ev_used = 0;
  f->timing = max_launches > 0;
  while ((int)f->ev.size() < max_launches) {
    cudaEvent_t a, b;
    VMS_CUDA(cudaEventCreate(&a));
    VMS_CUDA(cudaEventCreate(&b));
    f->ev.emplace_back(a, b);
  }
  return VMS_OK;
}

vms_status tcf_kernel_ms(vms_elbo_plan_s* pl, double* total_ms, int* launches) {
  TcfCfg* f = pl->tcf;
  if (!f) return VMS_OK;
  for (size_t i = 0; i < f->ev_used; ++i) {
    float ms = 0.f;
    VMS_CUDA(cudaEventSynchronize(f->ev[i].second));
    VMS_CUDA(cudaEventElapsedTime(&ms, f->ev[i].first, f->ev[i].second));
    *total_ms += ms;
  }
  *launches += (int)f->ev_used;
  f->ev_used = 0;
  return VMS_OK;
}

// forward + backward (+ Adam when `adam`): the tile kernel and the finish kernel; the pre-pack kernel only when the
// images are not known to match theta (first call, parameters changed behind the plan's back)
// launch configuration of the finish kernels: programmatic stream serialisation (the kernel may start while the tile kernel
// is still running and waits for it at griddepcontrol.wait); VMS_TCF_PDL=0 launches them the ordinary way
static cudaLaunchConfig_t finish_launch_config(unsigned grid, cudaStream_t st) {
  static cudaLaunchAttribute attr[1];
  static int pdl = -1;
  if (pdl < 0) {
    const char* e = getenv("VMS_TCF_PDL");
    pdl = (e && e[0] == '0') ? 0 : 1;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kFinP * kFinG);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cfg;
}

bool tcf_peer_ok(vms_elbo_plan_s* pl) {
  TcfCfg* f = pl->tcf;
  if (!f) return false;
  if (f->peer_ok < 0) {  // phase 2 of the fused finish + exchange kernel needs the whole grid resident
    int per_sm = 0;
    const bool ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tcf_finish_peer_kernel, kFinP * kFinG, 0) == cudaSuccess;
    const char* e = getenv("VMS_TCF_PEER_FINISH");
    f->peer_ok = (ok && (int64_t)per_sm * sm_count() >= (f->p.P + kFinP - 1) / kFinP && !(e && e[0] == '0')) ? 1 : 0;
  }
  return f->peer_ok == 1;
}

vms_status tcf_run(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B, float* grad,
                   float* scalars, cudaStream_t st, const FusedAdam* adam, const PeerArgs* peer) {
  TcfCfg* f = pl->tcf;
  VMS_REQUIRE(f && tcf_available(pl, B), VMS_ERR_UNSUPPORTED, "elbo (mode 3): shape or batch not supported");
  TParams p = f->p;
  p.B = B; p.theta = theta; p.x = x; p.eps = eps;
  p.wpk = f->wpk; p.fpk = f->fpk; p.gpart = f->gpart; p.spart = f->spart; p.err = pl->tc_err;
  p.rows = tcf_rows();
  const int n_tiles = (int)((B + p.rows - 1) / p.rows);
  if (!(f->pack_valid && f->pack_theta == theta)) {
    tcf_prepack_kernel<<<(p.P + 255) / 256, 256, 0, st>>>(theta, f->pack_map, p.P, f->fpk, f->wpk, f->part);
    VMS_LAUNCH_CHECK("tcf_prepack_kernel");
  }
  const bool timed = f->timing && f->ev_used < f->ev.size() && (pl->timing_calls++ % (unsigned)pl->timing_every) == 0u;
  if (timed) VMS_CUDA(cudaEventRecord(f->ev[f->ev_used].first, st));
#define VMS_TCF_LAUNCH(RPV, EX)                                                                                             \
  do {                                                                                                                      \
    VMS_CUDA(cudaFuncSetAttribute(tcf_kernel<RPV, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));         \
    tcf_kernel<RPV, EX><<<n_tiles, FT, f->smem, st>>>(p);                                                                   \
  } while (0)
  if (p.RP == 64) {
    if (f->exact) VMS_TCF_LAUNCH(64, true); else VMS_TCF_LAUNCH(64, false);
  } else {
    if (f->exact) VMS_TCF_LAUNCH(96, true); else VMS_TCF_LAUNCH(96, false);
  }
#undef VMS_TCF_LAUNCH
  VMS_LAUNCH_CHECK("tcf_kernel");
  if (timed) VMS_CUDA(cudaEventRecord(f->ev[f->ev_used++].second, st));
  if (peer) {  // data-parallel: finish + exchange + Adam + next images in one launch
    cudaLaunchConfig_t cfg = finish_launch_config((p.P + kFinP - 1) / kFinP, st);
    VMS_CUDA(cudaLaunchKernelEx(&cfg, tcf_finish_peer_kernel, (const float*)f->gpart, n_tiles, p.P, p.P2, (const int*)f->part_map,
                                (const float*)f->spart, B, p.klw, scalars ? scalars : pl->scalars, *peer,
                                (const unsigned*)f->pack_map, f->fpk, f->wpk, f->part));
    VMS_LAUNCH_CHECK("tcf_finish_peer_kernel");
    f->pack_valid = peer->theta == theta;
    f->pack_theta = theta;
    return VMS_OK;
  }
  TcfAdam ad = {};
  if (adam) {
    ad.theta = adam->theta; ad.m = adam->m; ad.v = adam->v;
    ad.lr_t = adam->lr_t; ad.one_minus_b1 = adam->one_minus_b1; ad.one_minus_b2 = adam->one_minus_b2; ad.eps = adam->eps;
  }
  cudaLaunchConfig_t cfg = finish_launch_config((p.P + kFinP - 1) / kFinP, st);
  VMS_CUDA(cudaLaunchKernelEx(&cfg, tcf_finish_kernel, (const float*)f->gpart, n_tiles, p.P, p.P2, (const int*)f->part_map, grad,
                              (const float*)f->spart, B, p.klw, scalars ? scalars : pl->scalars, ad,
                              (const unsigned*)f->pack_map, f->fpk, f->wpk, f->part));
  VMS_LAUNCH_CHECK("tcf_finish_kernel");
  // after a fused Adam step the images describe the updated parameters; without it nothing is known about what the caller
  // does to theta next (e.g. the data-parallel exchange updates it in its own kernel)
  f->pack_valid = adam != nullptr && adam->theta == theta;
  f->pack_theta = theta;
  return VMS_OK;
}

}  // namespace vms
