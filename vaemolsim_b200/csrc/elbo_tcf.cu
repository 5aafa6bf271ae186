// elbo_tcf.cu -- EXPERIMENTAL (plan mode 3; never chosen automatically): the whole ELBO training step of the C2 family
// as ONE persistent tensor-core kernel, one 64-row tile per CTA.  It is the plan of DESIGN.md 9a item 1: the
// coupling-block pipeline of flow_tc.cu (3 x BF16 split, kind::f16 MMAs, one TMEM accumulator per product, K-major and
// MN-major reads of the same tiles, octet spline routines on the raw parameters) for all blocks of the chain back to
// back, with the encoder / decoder MLPs of mlp_stream.cu as in-tile phases, so that at the named batch (4096 rows = 64
// tiles) the step is one launch whose heavy products run on the tensor core instead of the FFMA pipe of
// elbo_fused.cu.  Status (round 1, one measurement, scripts/check_elbo_tcf.py at batch 4096): CORRECT -- loss scalars equal
// to 7 digits, flat gradient 1.5e-7 (norm) against the FFMA plan, worst layer 1.6e-6 -- but SLOWER than the FFMA fused
// kernel: 0.187 ms against 0.146 ms forward + backward.  One 64-row tile per CTA keeps only 64 of 148 SMs busy and the
// RealNVP chain makes a tile's phases strictly sequential (block i needs block i + 1's output), so the tensor core
// idles while the SIMT phases run.  What it needs to win: 32 valid rows per tile (128 CTAs, SIMT phases
// halved, MMA cost unchanged), heads matrices staged with cp.async.bulk behind the previous block's epilogue, and two
// tiles per CTA on alternating warp groups.
// AFTER that measurement (no GPU time left in round 1) the row loops were parametrised by `rows` = valid rows per tile:
// rows = 64 is the default and is meant to be the measured code path unchanged; VMS_TCF_ROWS=32 selects 32-row tiles
// (zero-padded M = 64 products, d hW over 2 k-steps) and has NOT run on a device yet.  First thing to do next round:
// `python scripts/check_elbo_tcf.py 4096` with and without VMS_TCF_ROWS=32.
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
  float *gpart, *spart;
  int* err;
  int o_hid, o_graw, o_w, o_raw, o_u, o_gu, o_x, o_eps, o_pe, o_pd, o_gpd, o_gpe, o_gz, o_lp, o_gin, o_w1, o_b1;
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

// ------------------------------------------------------------------------------------------------ prepack
// heads matrix of every block with its bias as row H, split into three bfloat16 parts, in the shared-memory layout of
// the main kernel ([part][j / 8][RP][8]); one launch per step (theta changes every step)
__global__ void tcf_prepack_kernel(const TParams p, unsigned short* __restrict__ wpk) {
  const int blk = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int Hp = p.Hp, RP = p.RP, R = p.R, H = p.fh;
  if (e >= Hp * RP) return;
  const int j = e / RP, c = e - j * RP;
  float w = 0.f;
  if (c < R) {
    if (j < H) w = p.theta[p.blk[blk].off_hW + (size_t)j * R + c];
    else if (j == H) w = p.theta[p.blk[blk].off_hb + c];
  }
  unsigned h1, h2, h3;
  split3(w, h1, h2, h3);
  const size_t part = (size_t)(Hp / 8) * RP * 8;
  unsigned short* dst = wpk + (size_t)blk * 3 * part + (size_t)(j >> 3) * RP * 8 + (size_t)c * 8 + (j & 7);
  dst[0] = (unsigned short)(h1 >> 16);
  dst[part] = (unsigned short)(h2 >> 16);
  dst[2 * part] = (unsigned short)(h3 >> 16);
}

// ------------------------------------------------------------------------------------------------ MLP phases
// Weights of one two-layer MLP staged into `wm`: w0t [H][DI] = W0^T with b0 in column Din, w1 [H][DO] zero padded, b1 [DO]
__device__ void stage_mlp(const float* __restrict__ theta, int oW0, int ob0, int oW1, int ob1, int Din, int H, int Dout, int DI,
                          int DO, float* wm) {
  float* w0t = wm;
  float* w1 = wm + H * DI;
  float* b1 = w1 + H * DO;
  for (int e = threadIdx.x; e < H * DI; e += FT) {
    const int j = e / DI, i = e - j * DI;
    w0t[e] = i < Din ? __ldg(theta + oW0 + (size_t)i * H + j) : (i == Din ? __ldg(theta + ob0 + j) : 0.f);
  }
  for (int e = threadIdx.x; e < H * DO; e += FT) {
    const int j = e / DO, n = e - j * DO;
    w1[e] = n < Dout ? __ldg(theta + oW1 + (size_t)j * Dout + n) : 0.f;
  }
  for (int n = threadIdx.x; n < DO; n += FT) b1[n] = n < Dout ? __ldg(theta + ob1 + n) : 0.f;
  __syncthreads();
}

// row r of `in` ([FM][ldi], Din valid columns) extended by the bias input 1 and zero padded to 8
__device__ __forceinline__ void load_in8(const float* in, int ldi, int Din, int r, float (&x)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = i < Din ? in[r * ldi + i] : (i == Din ? 1.f : 0.f);
}

// out [FM][DO] = relu(in W0 + b0) W1 + b1;  scratch: [8][FM][DO] floats
__device__ void mlp_forward(const float* wm, int H, int DI, int DO, const float* in, int ldi, int Din, float* scratch,
                            float* out, int rows) {
  const float* w0t = wm;
  const float* w1 = wm + H * DI;
  const float* b1 = w1 + H * DO;
  const int ng = FT / rows;  // thread = (row, one of ng groups of hidden units)
  const int r = threadIdx.x & (rows - 1), g = threadIdx.x / rows;
  float x[8], acc[16];
  load_in8(in, ldi, Din, r, x);
#pragma unroll
  for (int n = 0; n < 16; ++n) acc[n] = 0.f;
  for (int j = g; j < H; j += ng) {
    float pre = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < DI) pre = fmaf(x[i], w0t[j * DI + i], pre);
    const float h = fmaxf(pre, 0.f);
#pragma unroll
    for (int n = 0; n < 16; ++n)
      if (n < DO) acc[n] = fmaf(h, w1[j * DO + n], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 16; ++n)
    if (n < DO) scratch[(g * rows + r) * DO + n] = acc[n];
  __syncthreads();
  for (int e = threadIdx.x; e < rows * DO; e += FT) {
    const int r2 = e / DO, n = e - r2 * DO;
    float s = b1[n];
    for (int k = 0; k < ng; ++k) s += scratch[(k * rows + r2) * DO + n];
    out[e] = s;
  }
  __syncthreads();
}

// reverse mode: gout [FM][DO] (zero rows beyond the tile) -> this CTA's weight-gradient partial (global, Keras layout),
// optional input gradient gin [FM][4] (Din <= 4).  scratch: max(H * 24, 8 * FM * 4) floats
__device__ void mlp_backward(const float* wm, int H, int DI, int DO, const float* in, int ldi, int Din, int Dout,
                             const float* gout, float* __restrict__ part, int oW0, int ob0, int oW1, int ob1, float* scratch,
                             float* gin, int rows) {
  const float* w0t = wm;
  const float* w1 = wm + H * DI;
  const int tid = threadIdx.x;
  {  // weight gradients: thread = (hidden unit j, half of the rows)
    const int j = tid & 255, half = tid >> 8;
    float w0[8], w1r[16], dW0[8], dW1[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { w0[i] = (j < H && i < DI) ? w0t[j * DI + i] : 0.f; dW0[i] = 0.f; }
#pragma unroll
    for (int n = 0; n < 16; ++n) { w1r[n] = (j < H && n < DO) ? w1[j * DO + n] : 0.f; dW1[n] = 0.f; }
    if (j < H) {
      for (int r = half * (rows / 2); r < (half + 1) * (rows / 2); ++r) {
        float x[8];
        load_in8(in, ldi, Din, r, x);
        float pre = 0.f, t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) pre = fmaf(x[i], w0[i], pre);
        float g[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) g[n] = n < DO ? gout[r * DO + n] : 0.f;
#pragma unroll
        for (int n = 0; n < 16; ++n) t = fmaf(g[n], w1r[n], t);
        const float h = fmaxf(pre, 0.f);
        const float gh = pre > 0.f ? t : 0.f;
#pragma unroll
        for (int n = 0; n < 16; ++n) dW1[n] = fmaf(h, g[n], dW1[n]);
#pragma unroll
        for (int i = 0; i < 8; ++i) dW0[i] = fmaf(x[i], gh, dW0[i]);
      }
      if (half == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) scratch[j * 24 + i] = dW0[i];
#pragma unroll
        for (int n = 0; n < 16; ++n) scratch[j * 24 + 8 + n] = dW1[n];
      }
    }
    __syncthreads();
    if (j < H && half == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = dW0[i] + scratch[j * 24 + i];
        if (i < Din) part[oW0 + (size_t)i * H + j] = v;
        else if (i == Din) part[ob0 + j] = v;
      }
#pragma unroll
      for (int n = 0; n < 16; ++n)
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
  if (gin) {  // input gradient: thread = (row, 1/8 of the hidden units)
    const int ng = FT / rows;
    const int r = tid & (rows - 1), g8 = tid / rows;
    float x[8], g[16], gi[4] = {0.f, 0.f, 0.f, 0.f};
    load_in8(in, ldi, Din, r, x);
#pragma unroll
    for (int n = 0; n < 16; ++n) g[n] = n < DO ? gout[r * DO + n] : 0.f;
    for (int j = g8; j < H; j += ng) {
      float pre = 0.f, t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < DI) pre = fmaf(x[i], w0t[j * DI + i], pre);
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < DO) t = fmaf(g[n], w1[j * DO + n], t);
      const float gh = pre > 0.f ? t : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < Din) gi[i] = fmaf(gh, w0t[j * DI + i], gi[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) scratch[(g8 * rows + r) * 4 + i] = gi[i];
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
template <int RP>
__global__ void __launch_bounds__(FT, 1) tcf_kernel(const __grid_constant__ TParams p) {
  extern __shared__ __align__(128) unsigned char smb[];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ unsigned tmem_base_s;
  __shared__ float red[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
  float* s_mlp = reinterpret_cast<float*>(smb + p.o_graw);  // MLP weights borrow the g_raw tiles outside the flow's reverse mode
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
  float* s_w1 = reinterpret_cast<float*>(smb + p.o_w1);     // [nb][4][Hp]
  float* s_b1 = reinterpret_cast<float*>(smb + p.o_b1);     // [nb][Hp]
  const int DOE = p.DOE, DOD = p.DOD;
  const int rows = p.rows;  // valid rows of a tile (<= FM)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  // conditioner first layers of every block
  for (int e = tid; e < nb * 4 * Hp; e += FT) {
    const int b = e / (4 * Hp), c = (e / Hp) & 3, j = e % Hp;
    s_w1[e] = (c < p.blk[b].cin && j < H) ? __ldg(p.theta + p.blk[b].off_d1W + (size_t)c * H + j) : 0.f;
  }
  for (int e = tid; e < nb * Hp; e += FT) {
    const int b = e / Hp, j = e - b * Hp;
    s_b1[e] = j < H ? __ldg(p.theta + p.blk[b].off_d1b + j) : 0.f;
  }
  if (rows < FM)  // rows beyond the tile must be finite (and zero) operands of the M = 64 products
    for (unsigned e = tid; e < 3 * hid_sz / 4; e += FT) reinterpret_cast<unsigned*>(hid)[e] = 0u;
  tc_sync();
  const unsigned tm = tmem_base_s;
  const unsigned tD1 = tm, tD2 = tm + RP, tD3 = tm + RP + 128;
  const unsigned id1 = idesc_bf16(FM, RP, 0, 0);
  const unsigned id2 = idesc_bf16(FM, Hp, 0, 1);
  const unsigned id3 = idesc_bf16(128, RP, 1, 1);
  const unsigned bar = smem_u32(&mbar);
  const unsigned hid_a = smem_u32(hid), graw_a = smem_u32(graw), w_a = smem_u32(wsm);
  unsigned phase = 0;
  bool failed = false;

  const int64_t tile = blockIdx.x;  // one tile per CTA
  const int64_t row0 = tile * rows;
  const int nr = (int)min((int64_t)rows, p.B - row0);
  float* part = p.gpart + (size_t)blockIdx.x * p.P;
  const float invB = 1.0f / (float)p.B;
  const float g_logpx = -invB, g_logq = p.klw * invB, g_logpz = -p.klw * invB;

  // ---- heads matrix of block i -> shared memory (generic copy of the pre-split tiles, then visible to the tensor core)
  auto stage_heads = [&](int i) {
    const uint4* src = reinterpret_cast<const uint4*>(p.wpk + (size_t)i * 3 * (size_t)nchH * RP * 8);
    uint4* dst = reinterpret_cast<uint4*>(wsm);
    const int n16 = (int)(3u * w_sz / 16u);
    for (int e = tid; e < n16; e += FT) dst[e] = __ldg(src + e);
  };
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
  auto raw_product = [&]() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_sync();
    if (tid == 0) {
      issue_product(tD1, hid_a, hid_sz, 2u * CSB, CSB, 128, w_a, w_sz, 2u * w_cs, w_cs, 128, Hp / 16, id1);
      mma_commit(bar);
    }
    if (!mbar_wait_bounded(bar, phase)) failed = true;
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
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
  stage_mlp(p.theta, p.enc0W, p.enc0b, p.enc1W, p.enc1b, dx, p.H, 2 * dz, p.DIE, DOE, s_mlp);  // (ends with a barrier)
  mlp_forward(s_mlp, p.H, p.DIE, DOE, s_x, 8, dx, s_raw, s_pe, rows);
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
    stage_heads(i);
    build_hid(i, uin);
    raw_product();
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
  stage_mlp(p.theta, p.dec0W, p.dec0b, p.dec1W, p.dec1b, dz, p.H, 2 * dx, p.DID, DOD, s_mlp);
  const float* zt = s_u + nb * FM * 4;
  mlp_forward(s_mlp, p.H, p.DID, DOD, zt, 4, dz, s_raw, s_pd, rows);
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
  mlp_backward(s_mlp, p.H, p.DID, DOD, zt, 4, dz, 2 * dx, s_gpd, part, p.dec0W, p.dec0b, p.dec1W, p.dec1b, s_raw, s_gz,
               rows);
  // ================================================================ F': flow reverse mode, block 0 first
  for (unsigned e = tid; e < 3 * graw_sz / 4; e += FT) reinterpret_cast<unsigned*>(graw)[e] = 0u;  // (held the MLP weights)
  if (tid < rows)
    for (int d = 0; d < 4; ++d) s_gu[tid * 4 + d] = d < dz ? -g_logpz * s_u[tid * 4 + d] : 0.f;
  __syncthreads();
#pragma unroll 1
  for (int i = 0; i < nb; ++i) {
    const TBlk& fb = p.blk[i];
    const float* uin = s_u + (i + 1) * FM * 4;
    stage_heads(i);
    build_hid(i, uin);
    raw_product();
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
    {  // d [hW; hb] of this tile -> the CTA's partial (thread = hidden unit 32 q + lane, 16-column chunks sub, sub + 4)
      const int q = warp & 3, sub = warp >> 2, jj = 32 * q + lane;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c0 = 16 * (sub + 4 * h);
        if (c0 < RP) {
          float v1[16];
          tmem_ld16(tD3 + ((unsigned)(32 * q) << 16) + (unsigned)c0, v1);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int n = c0 + k;
            if (n < R) {
              if (jj < H) part[fb.off_hW + (size_t)jj * R + n] = v1[k];
              else if (jj == H) part[fb.off_hb + n] = v1[k];
            }
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
  stage_mlp(p.theta, p.enc0W, p.enc0b, p.enc1W, p.enc1b, dx, p.H, 2 * dz, p.DIE, DOE, s_mlp);
  mlp_backward(s_mlp, p.H, p.DIE, DOE, s_x, 8, dx, 2 * dz, s_gpe, part, p.enc0W, p.enc0b, p.enc1W, p.enc1b, s_raw, nullptr,
               rows);

  if (failed && tid == 0 && p.err) atomicExch(p.err, 1);
  tc_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tm), "r"(512u) : "memory");
}

// grad[i] = sum over the tiles' partials in a fixed order (+ Keras Adam), and the three loss scalars
struct TcfAdam {
  float *theta, *m, *v;
  float lr_t, one_minus_b1, one_minus_b2, eps;
};
__global__ void __launch_bounds__(256) tcf_finish_kernel(const float* __restrict__ gpart, int n_part, int P,
                                                         float* __restrict__ grad, const float* __restrict__ spart,
                                                         int64_t B, float klw, float* __restrict__ scalars, const TcfAdam ad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int c = 0;
    for (; c + 4 <= n_part; c += 4) {
      s0 += gpart[(size_t)c * P + i];
      s1 += gpart[(size_t)(c + 1) * P + i];
      s2 += gpart[(size_t)(c + 2) * P + i];
      s3 += gpart[(size_t)(c + 3) * P + i];
    }
    for (; c < n_part; ++c) s0 += gpart[(size_t)c * P + i];
    const float t = (s0 + s1) + (s2 + s3);
    grad[i] = t;
    if (ad.theta) {
      const float mi = ad.m[i] + (t - ad.m[i]) * ad.one_minus_b1;
      const float vi = ad.v[i] + (t * t - ad.v[i]) * ad.one_minus_b2;
      ad.m[i] = mi;
      ad.v[i] = vi;
      ad.theta[i] = ad.theta[i] - ad.lr_t * mi / (sqrtf(vi) + ad.eps);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && scalars) {
    float sa = 0.f, sc = 0.f;
    for (int c = 0; c < n_part; ++c) { sa += spart[2 * c]; sc += spart[2 * c + 1]; }
    const float kl = sa / (float)B, nll = sc / (float)B;
    scalars[0] = nll + klw * kl;
    scalars[1] = nll;
    scalars[2] = kl;
  }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

struct TcfCfg {
  TParams p;
  size_t smem;
  int max_tiles;
  float *gpart, *spart;
  unsigned short* wpk;
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
  p.DIE = round_up(d.dx + 1, 4); p.DOE = round_up(2 * d.dz, 4);
  p.DID = round_up(d.dz + 1, 4); p.DOD = round_up(2 * d.dx, 4);
  for (int i = 0; i < d.num_blocks; ++i) {
    const FlowBlock& b = pl->blocks[i];
    TBlk& t = p.blk[i];
    t.cs0 = b.cs0; t.nc = b.cs1 - b.cs0; t.ts0 = b.ts0; t.cin = b.cin;
    t.off_d1W = (int)b.off_d1W; t.off_d1b = (int)b.off_d1b; t.off_hW = (int)b.off_hW; t.off_hb = (int)b.off_hb;
  }
  const int nchH = p.Hp / 8, nchR = p.RP / 8;
  int off = 0;
  auto take = [&](int bytes) { int o0 = off; off += round_up(bytes, 16); return o0; };
  p.o_hid = take(3 * nchH * (int)CSB);
  const int mlp_floats = d.hidden * ((p.DIE + p.DOE) > (p.DID + p.DOD) ? (p.DIE + p.DOE) : (p.DID + p.DOD)) + 16;
  const int graw_bytes = 3 * nchR * (int)CSB;
  p.o_graw = take(graw_bytes > 4 * mlp_floats ? graw_bytes : 4 * mlp_floats);
  p.o_w = take(3 * nchH * p.RP * 16);
  // s_raw doubles as scratch of the MLP phases: 8 row groups x FM x 16 outputs, or hidden x 24 accumulators
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
  p.o_w1 = take(d.num_blocks * 4 * p.Hp * 4); p.o_b1 = take(d.num_blocks * p.Hp * 4);
  f->smem = (size_t)off;
  if (f->smem + 1024 > (size_t)max_smem_optin()) { delete f; return VMS_OK; }
  f->max_tiles = sm_count();
  void *g = nullptr, *s = nullptr, *w = nullptr;
  const size_t wpk_bytes = (size_t)d.num_blocks * 3 * nchH * p.RP * 8 * sizeof(unsigned short);
  if (cudaMalloc(&g, (size_t)f->max_tiles * p.P * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&s, (size_t)f->max_tiles * 2 * sizeof(float)) != cudaSuccess || cudaMalloc(&w, wpk_bytes) != cudaSuccess) {
    cudaGetLastError();
    if (g) cudaFree(g);
    if (s) cudaFree(s);
    delete f;
    return VMS_OK;  // experimental path: simply unavailable
  }
  f->gpart = (float*)g; f->spart = (float*)s; f->wpk = (unsigned short*)w;
  pl->tcf = f;
  return VMS_OK;
}

void tcf_destroy(vms_elbo_plan_s* pl) {
  if (!pl->tcf) return;
  cudaFree(pl->tcf->gpart);
  cudaFree(pl->tcf->spart);
  cudaFree(pl->tcf->wpk);
  delete pl->tcf;
  pl->tcf = nullptr;
}

// valid rows per tile: 64 (the measured configuration), or 32 with VMS_TCF_ROWS=32 (UNTESTED on the device at the end
// of round 1: twice the CTAs, SIMT phases halved, the M = 64 products padded with zero rows)
static int tcf_rows() {
  const char* e = getenv("VMS_TCF_ROWS");
  return (e && atoi(e) == 32) ? 32 : FM;
}

bool tcf_available(const vms_elbo_plan_s* pl, int64_t B) {
  const int rows = tcf_rows();
  return pl->tcf && (B + rows - 1) / rows <= pl->tcf->max_tiles;
}

// forward + backward (+ Adam when `adam`): 3 launches (prepack, the tile kernel, finish)
vms_status tcf_run(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B, float* grad,
                   float* scalars, cudaStream_t st, const FusedAdam* adam) {
  TcfCfg* f = pl->tcf;
  VMS_REQUIRE(f && tcf_available(pl, B), VMS_ERR_UNSUPPORTED, "elbo (mode 3): shape or batch not supported");
  TParams p = f->p;
  p.B = B; p.theta = theta; p.x = x; p.eps = eps;
  p.wpk = f->wpk; p.gpart = f->gpart; p.spart = f->spart; p.err = pl->tc_err;
  p.rows = tcf_rows();
  const int n_tiles = (int)((B + p.rows - 1) / p.rows);
  tcf_prepack_kernel<<<dim3((p.Hp * p.RP + 255) / 256, p.nb), 256, 0, st>>>(p, f->wpk);
  VMS_LAUNCH_CHECK("tcf_prepack_kernel");
  if (p.RP == 64) {
    VMS_CUDA(cudaFuncSetAttribute(tcf_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));
    tcf_kernel<64><<<n_tiles, FT, f->smem, st>>>(p);
  } else {
    VMS_CUDA(cudaFuncSetAttribute(tcf_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem));
    tcf_kernel<96><<<n_tiles, FT, f->smem, st>>>(p);
  }
  VMS_LAUNCH_CHECK("tcf_kernel");
  TcfAdam ad = {};
  if (adam) {
    ad.theta = adam->theta; ad.m = adam->m; ad.v = adam->v;
    ad.lr_t = adam->lr_t; ad.one_minus_b1 = adam->one_minus_b1; ad.one_minus_b2 = adam->one_minus_b2; ad.eps = adam->eps;
  }
  tcf_finish_kernel<<<(p.P + 255) / 256, 256, 0, st>>>(f->gpart, n_tiles, p.P, grad, f->spart, B, p.klw,
                                                      scalars ? scalars : pl->scalars, ad);
  VMS_LAUNCH_CHECK("tcf_finish_kernel");
  return VMS_OK;
}

}  // namespace vms
