// elbo_fused.cu -- the whole VAE ELBO step (forward, or forward + backward) as ONE persistent sm_100a kernel.
//
// Replaces, for the model family of the reference's tests (tests/test_models.py:161-228; C1 / C2 of SURVEY 8d):
//   models.py:289-322  VAE.call  (encoder -> sample -> prior -> regulariser -> decoder)
//   mappings.py:125-155 FCDeepNN.call (Dense relu -> Dense), tfp.layers.IndependentNormal
//   dists.py:414-439   FlowedDistribution.call + tfp TransformedDistribution.log_prob
//   flows.py:154-207, :281-355  SplineBijector / RQSSplineRealNVP (density direction = chain inverse)
//   losses.py:58, :253 LogProbLoss (batch mean) and KLDivergenceEstimate
// and TF autodiff through all of it.  Same arithmetic as the unfused plan in elbo.cu (which stays as the generic
// path and as an on-device cross-check); this file changes WHERE the intermediates live.
//
// Design (B200).  A configuration row needs 44 KB of weights (L2-resident, shared by all rows) and ~5 KB of
// activations that the unfused path round-trips through HBM/L2 between ~55 launches.  Here one CTA owns a tile of
// 32 rows and keeps every activation of the tile in shared memory (~210 KB of the 227 KB) from the first encoder
// layer to the last weight gradient:
//   * grid = min(#tiles, #SMs) persistent CTAs of 512 threads (batch 4096 -> 128 CTAs, one tile each);
//   * wide layers (the 100 x 95 spline-parameter heads, the 200-wide MLP layers) are FP32 FFMA "outer-product" GEMMs:
//     a lane owns output columns (coalesced / conflict-free operand), a warp owns a slab of 4..16 output rows whose
//     operand is read by 128-bit shared-memory broadcast, accumulators stay in registers.  Forward, input-gradient
//     and weight-gradient GEMMs are the same routine with different operand strides;
//   * thin layers (2..12 outputs) split the contraction over the 16 warps and combine partial sums in shared memory;
//   * the spline uses the octet routines of rqs_device.cuh on shared-memory logits (32 octets = 32 rows);
//   * weight gradients leave the CTA once per tile, into a per-CTA partial gradient (plain stores, deterministic);
//     a second small kernel sums the partials in fixed order and finishes the three loss scalars.
// FP32 FFMA rather than tcgen05: parity is 1e-5 relative in float32 (TF32 inputs give 1e-3), contraction lengths
// are 1..200, and a 32-row tile is below the 128-row UMMA tile -- see DESIGN.md "Why not tensor cores here".
#include "elbo_plan.cuh"
#include "rqs_device.cuh"
#include "tile_gemm.cuh"
#include <string.h>

namespace vms {

constexpr int kMaxBlocks = 8;

struct FBlk {
  int cs0, nc, ts0, dt, cin, ldr;
  int off_d1W, off_d1b, off_hW, off_hb;
};

struct FusedParams {
  int dx, dz, hidden, nb, K, fh;
  float bin_min, scale, klw;
  int64_t B;
  int n_tiles, P, n_mlp;  // n_mlp: number of leading floats of theta (encoder + decoder) kept resident in shared memory
  int enc0W, enc0b, enc1W, enc1b, dec0W, dec0b, dec1W, dec1b;
  FBlk blk[kMaxBlocks];
  const float *theta, *x, *eps;
  const float* pack;                // re-packed flow-block weights (prepack_kernel)
  int pack_stride, pack_wt, pack_small;
  float *z, *logq, *logpz, *logpx;  // optional per-row outputs
  float *gpart, *spart;             // [grid][P] partial gradients, [grid][2] partial loss sums
  // shared-memory pitches and offsets (floats)
  int ldx, ldz, ldh, ldf, ldfp, ldpe, ldpd, ldc, ldrm, ldwt, sb;
  int o_xs, o_xT, o_eps, o_zR, o_zT, o_he, o_hd, o_pe, o_pd, o_u, o_lp, o_hid, o_cond, o_raw, o_W, o_Wp, o_B, o_gz,
      o_gua, o_gub, o_scr, o_tsc, o_bar;
};

struct FusedCfg {
  FusedParams p;
  size_t smem_bytes;
  int max_grid;
  float *gpart, *spart, *pack;
  // optional per-launch timing of the main kernel (bench.py's roofline leg): event pairs recorded around it
  bool timing = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t ev_used = 0;
};

// ------------------------------------------------------------------------------------------------ bulk staging
// Flow-block weights reach shared memory as ONE bulk asynchronous copy per block (cp.async.bulk, the 1-D TMA path,
// completion signalled on an mbarrier): a single thread issues it, nobody spends instructions on it.  The first
// version staged with 4-byte cp.async (rows of 3K-1 floats are not 16-byte aligned in theta) -- 9,500 LDGSTS per block
// and 12 % of the kernel's samples stalled in that loop.  Bulk copies need 16-byte aligned, 16-byte-multiple blocks,
// so a tiny kernel first re-packs theta's flow blocks into the plan's `pack` buffer, per block:
//   [ hW as [fh][ldrm] | hW^T as [ldr][ldwt] | d1W d1b hb (sb floats) ]     (prepack_kernel, once per step)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// Issue the copy of block `blk`'s packed weights (thread 0 only; callers make sure every earlier reader of the staging
// buffers has passed a __syncthreads()).  W <- natural or transposed heads weight, B[blk & 1] <- d1W | d1b | hb.
__device__ __forceinline__ void stage_block(const FusedParams& p, int blk, bool transposed) {
  if (threadIdx.x != 0) return;
  extern __shared__ __align__(16) float sm[];
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + p.o_bar);
  const float* src = p.pack + (size_t)blk * p.pack_stride;
  const unsigned wn = (unsigned)(p.fh * p.ldrm) * 4u, wt = (unsigned)(p.blk[blk].ldr * p.ldwt) * 4u, sb = (unsigned)p.sb * 4u;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy reads of the buffers are done
  mbar_expect_tx(bar, (transposed ? wt : wn) + sb);
  bulk_g2s(sm + p.o_W, transposed ? src + p.pack_wt : src, transposed ? wt : wn, bar);
  bulk_g2s(sm + p.o_B + (blk & 1) * p.sb, src + p.pack_small, sb, bar);
}

// One thread per SOURCE element of a block's [hW | hb] and [d1W | d1b]: coalesced reads of theta, two scattered
// writes (natural + transposed layouts).  The padding columns of `pack` are zeroed once, at plan creation.
__global__ void __launch_bounds__(256) prepack_kernel(const FusedParams p, float* __restrict__ pack) {
  const int blk = blockIdx.y;
  const FBlk& fb = p.blk[blk];
  float* dst = pack + (size_t)blk * p.pack_stride;
  const float* th = p.theta;
  const int n_hw = p.fh * fb.ldr, n_small = fb.cin * p.fh + p.fh;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_hw) {
    const float v = __ldg(th + fb.off_hW + e);
    const int k = e / fb.ldr, n = e - k * fb.ldr;
    dst[k * p.ldrm + n] = v;
    dst[p.pack_wt + n * p.ldwt + k] = v;
  } else if (e < n_hw + n_small) {
    dst[p.pack_small + (e - n_hw)] = __ldg(th + fb.off_d1W + (e - n_hw));
  } else if (e < n_hw + n_small + fb.ldr) {
    dst[p.pack_small + (e - n_hw)] = __ldg(th + fb.off_hb + (e - n_hw - n_small));
  }
}

// (A 6-instruction tanh built on ex2.approx / rcp.approx was tried here: |error| 3e-7 on hid moved log p(z) by up to
// 7e-5 on the ill-conditioned test flows -- outside the 1e-5 parity budget -- so the accurate tanhf stays.)
// Conditioner hidden layer hid = tanh(cond d1W + d1b) of block `blk`; an empty conditioner input is ones((B,1))
// (flows.py:184-185).  row_major = false: hidT [fh][FR] (operand of the forward GEMM);  true: hidR [FR][ldf] and the
// row-major conditioner input condR (operands of the weight gradients).  Executed by warps [w0, w0 + nw) only, so the
// backward pass can run it beside the spline's reverse mode (which occupies 8 of the 16 warps).
__device__ __noinline__ void hidden_layer(const FusedParams& p, int blk, bool row_major, int w0, int nw) {
  extern __shared__ __align__(16) float sm[];
  const FBlk& fb = p.blk[blk];
  const int fh = p.fh, dz = p.dz;
  const float* uin = sm + p.o_u + (blk + 1) * FR * dz;
  const float* Bb = sm + p.o_B + (blk & 1) * p.sb;
  const float* d1b = Bb + fb.cin * fh;
  float* hid = sm + p.o_hid;
  const int warp = (int)(threadIdx.x >> 5) - w0, lane = threadIdx.x & 31;
  if (warp < 0 || warp >= nw) return;
  if (!row_major) {
    // a warp takes features j = warp, warp + nw, ...; lanes take the 32 rows
    const int r = lane;
#pragma unroll 2
    for (int j = warp; j < fh; j += nw) {
      float a = d1b[j];
      for (int c = 0; c < fb.cin; ++c) a = fmaf(fb.nc > 0 ? uin[r * dz + fb.cs0 + c] : 1.f, Bb[c * fh + j], a);
      hid[j * FR + r] = tanhf(a);
    }
  } else {
    // a warp takes rows r = warp, warp + nw, ...; lanes take consecutive j
#pragma unroll 1
    for (int r = warp; r < FR; r += nw) {
#pragma unroll 4
      for (int j = lane; j < fh; j += 32) {
        float a = d1b[j];
        for (int c = 0; c < fb.cin; ++c) a = fmaf(fb.nc > 0 ? uin[r * dz + fb.cs0 + c] : 1.f, Bb[c * fh + j], a);
        hid[r * p.ldf + j] = tanhf(a);
      }
      if (lane <= fb.cin)
        sm[p.o_cond + r * p.ldc + lane] = (lane == fb.cin || fb.nc == 0) ? 1.f : uin[r * dz + fb.cs0 + lane];
    }
  }
}

// Spline of block `blk` in the density direction (inverse), one octet per row: u[blk] <- inverse(u[blk+1]), lpz += ildj.
__device__ __noinline__ void spline_forward(const FusedParams& p, int blk) {
  extern __shared__ __align__(16) float sm[];
  const FBlk& fb = p.blk[blk];
  const int dz = p.dz, K = p.K;
  const float* uin = sm + p.o_u + (blk + 1) * FR * dz;
  float* uout = sm + p.o_u + blk * FR * dz;
  const int r = threadIdx.x >> 3, j = threadIdx.x & 7;
  if (r >= FR) return;  // whole warps: 4 rows per warp, FR rows in all
  const float* rr = sm + p.o_raw + (blk * FR + r) * p.ldrm;
  float ldj_acc = 0.f;
  for (int d = 0; d < fb.dt; ++d) {
    const float v = uin[r * dz + fb.ts0 + d];
    float out, ldj, ldj_all;
    bool writer;
    rqsdev::octet_apply<4, true, false>(rr + d * K, rr + fb.dt * K + d * K, rr + 2 * fb.dt * K + d * (K - 1), v, j, K,
                                        true, p.bin_min, p.scale, out, ldj, ldj_all, writer);
    if (writer) uout[r * dz + fb.ts0 + d] = out;
    ldj_acc += ldj_all;
  }
  if (j == 0) sm[p.o_lp + FR + r] += ldj_acc;
  for (int c = j; c < fb.nc; c += 8) uout[r * dz + fb.cs0 + c] = uin[r * dz + fb.cs0 + c];
}

// Reverse mode of spline_forward: gnxt[ts] <- g_in, gnxt[cs] <- gcur[cs], raw-logit gradients in both layouts
// (grawR [FR][ldrm] for the weight gradient, grawT [ldrm][FR] for the input gradient).
__device__ __noinline__ void spline_backward(const FusedParams& p, int blk, int gcur, int gnxt, int nr, float g_logpz) {
  extern __shared__ __align__(16) float sm[];
  const FBlk& fb = p.blk[blk];
  const int dz = p.dz, K = p.K, ldrm = p.ldrm;
  const float* uin = sm + p.o_u + (blk + 1) * FR * dz;
  const int r = threadIdx.x >> 3, j = threadIdx.x & 7;
  if (r >= FR) return;  // whole warps
  const float* rr = sm + p.o_raw + (blk * FR + r) * ldrm;
  float* gr = sm + p.o_scr + r * ldrm;
  float* grawT = sm + p.o_scr + FR * ldrm;
  const float g_ldj = r < nr ? g_logpz : 0.f;
  for (int d = 0; d < fb.dt; ++d) {
    const float v = uin[r * dz + fb.ts0 + d];
    const float g_out = sm[gcur + r * dz + fb.ts0 + d];
    float g_in, gw[4], gh[4], gs[4];
    bool writer;
    const int ow = d * K, oh = fb.dt * K + d * K, os = 2 * fb.dt * K + d * (K - 1);
    rqsdev::octet_backward<4, true, false>(rr + ow, rr + oh, rr + os, v, g_out, g_ldj, j, K, true, p.bin_min, p.scale,
                                           g_in, writer, gw, gh, gs);
    if (writer) sm[gnxt + r * dz + fb.ts0 + d] = g_in;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 4 * j + q;
      if (k < K) {
        gr[ow + k] = gw[q];
        gr[oh + k] = gh[q];
        grawT[(ow + k) * FR + r] = gw[q];
        grawT[(oh + k) * FR + r] = gh[q];
      }
      if (k < K - 1) {
        gr[os + k] = gs[q];
        grawT[(os + k) * FR + r] = gs[q];
      }
    }
  }
  for (int c = j; c < fb.nc; c += 8) sm[gnxt + r * dz + fb.cs0 + c] = sm[gcur + r * dz + fb.cs0 + c];
}

// ------------------------------------------------------------------------------------------------ the kernel
template <bool BWD>
__global__ void __launch_bounds__(FT, 1) elbo_fused_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the finish kernel may be scheduled early (it waits)
  const int dx = p.dx, dz = p.dz, H = p.hidden, nb = p.nb, fh = p.fh;
  const int ldh = p.ldh, ldrm = p.ldrm;
  // shared-memory map (offsets in floats; see fused_create):
  //   xs [FR][ldx] x row-major, column dx = 1   xT [dx][FR]   eps [FR][dz]   zR [FR][ldz] column dz = 1   zT [dz][FR]
  //   he, hd [FR][ldh] column H = 1 (bias rows of the enc.1 / dec.1 weight gradients)   pe [FR][ldpe]   pd [FR][ldpd]
  //   u [nb+1][FR][dz] chain-inverse states (u[nb] = z, u[0] = base sample)   lp [3][FR] = log q, log p(z), log p(x|z)
  //   hid: forward hidT [fh][FR]; backward (same storage) hidR [FR][ldf], column fh = 1      cond [FR][ldc]
  //   raw [nb][FR][ldrm] raw spline parameters (kept for the backward pass)
  //   W: staged heads weight [fh][ldrm] (forward) or its transpose [ldr][ldwt] (backward)
  //   Wp: encoder + decoder weights, resident for the CTA's lifetime     B [2][sb]: d1W | d1b | hb by block parity
  //   gz, gua, gub [FR][dz]      scr: backward scratch, re-carved per phase
  float* xs = sm + p.o_xs;
  float* epsS = sm + p.o_eps;
  float* zR = sm + p.o_zR;
  float* pe = sm + p.o_pe;
  float* pd = sm + p.o_pd;
  float* u = sm + p.o_u;
  float* lq = sm + p.o_lp;
  float* lpz = lq + FR;
  float* gz = sm + p.o_gz;
  float* gp = p.gpart + (size_t)blockIdx.x * p.P;

  // CTA-lifetime state: resident MLP weights, constant columns
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + p.o_bar);
  unsigned phase = 0;   // parity of the staging mbarrier: one copy batch in flight at a time, every thread waits for each
  bool pending = false;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  for (int i = tid; i < p.n_mlp; i += FT) cp_async4(sm + p.o_Wp + i, p.theta + i);
  if (nb > 0) {
    stage_block(p, nb - 1, false);
    pending = true;
  }
  for (int i = tid; i < FR * p.ldz; i += FT) zR[i] = (i % p.ldz) == dz ? 1.f : 0.f;
  for (int i = tid; i < FR * p.ldc; i += FT) sm[p.o_cond + i] = 0.f;
  float cta_kl = 0.f, cta_nll = 0.f;
  const float invB = 1.0f / (float)p.B;
  bool first = true;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * FR;
    const int nr = (int)min((int64_t)FR, p.B - row0);
    const bool more = tile + (int)gridDim.x < p.n_tiles;
    // ---------------------------------------------------------------- F0: stage the tile's inputs
    for (int i = tid; i < FR * p.ldx; i += FT) {
      const int r = i / p.ldx, c = i - r * p.ldx;
      float v = c == dx ? 1.f : 0.f;
      if (c < dx && r < nr) v = __ldg(p.x + (row0 + r) * dx + c);
      xs[i] = v;
      if (c < dx) sm[p.o_xT + c * FR + r] = v;
    }
    for (int i = tid; i < FR * dz; i += FT) epsS[i] = i < nr * dz ? __ldg(p.eps + row0 * dz + i) : 0.f;
    for (int i = tid; i < FR; i += FT) {  // bias rows of the enc.1 / dec.1 weight gradients 
      sm[p.o_he + i * ldh + H] = 1.f;
      sm[p.o_hd + i * ldh + H] = 1.f;
    }
    cp_async_commit_wait_all();
    if (pending) {
      mbar_wait(bar, phase);
      phase ^= 1;
      pending = false;
    }
    __syncthreads();
    // ---------------------------------------------------------------- F1: he = relu(x W + b)   (mappings.py:151-153)
    outer_gemm<8, 2, false>(p.o_xT, FR, p.o_Wp + p.enc0W, H, 1, FR, H, dx, 0, 0,
                             epi_store(p.o_he, ldh, p.o_Wp + p.enc0b, 1));
    __syncthreads();
    // ---------------------------------------------------------------- F2: encoder head parameters
    thin_gemm(p.o_he, ldh, p.o_Wp + p.enc1W, 2 * dz, 1, H, 2 * dz, p.o_pe, p.ldpe, p.o_Wp + p.enc1b, 0, p.o_scr);
    __syncthreads();
    // ---------------------------------------------------------------- F3: z = eps * softplus(raw) + loc, log q(z|x)
    if (tid < FR) {
      const int r = tid;
      float s = 0.f;
      for (int d = 0; d < dz; ++d) {
        const float loc = pe[r * p.ldpe + d], sc = softplus_tf(pe[r * p.ldpe + dz + d]);
        const float zz = __fadd_rn(__fmul_rn(epsS[r * dz + d], sc), loc);  // separate TF mul and add ops
        s += normal_lp(zz, loc, sc);
        u[(nb * FR + r) * dz + d] = zz;
        zR[r * p.ldz + d] = zz;
        sm[p.o_zT + d * FR + r] = zz;
      }
      lq[r] = s;
      lpz[r] = 0.f;
    }
    __syncthreads();
    // ---------------------------------------------------------------- F4: prior log p(z), chain inverse (flows.py:323)
#pragma unroll 1
    for (int i = nb - 1; i >= 0; --i) {
      const int hb = p.o_B + (i & 1) * p.sb + p.blk[i].cin * fh + fh;
      const int raw_i = p.o_raw + i * FR * ldrm;
      hidden_layer(p, i, false, 0, FW);
      __syncthreads();
      // raw = hid hW + hb: the three Dense heads of flows.py:140-152 as one GEMM
      outer_gemm<4, 3, true>(p.o_hid, FR, p.o_W, ldrm, 1, FR, p.blk[i].ldr, fh, raw_i, ldrm, epi_store(raw_i, ldrm, hb, 0));
      __syncthreads();
      // the staging buffer is free: fetch the next block's weights (or block 0's transposed, for the backward pass)
      // while the spline runs
      if (i > 0 || BWD) {
        stage_block(p, i > 0 ? i - 1 : 0, i == 0);
        pending = true;
      }
      spline_forward(p, i);
      if (pending) {
        mbar_wait(bar, phase);
        phase ^= 1;
        pending = false;
      }
      __syncthreads();
    }
    if (tid < FR) {
      float s = lpz[tid];
      for (int d = 0; d < dz; ++d) s += normal_lp(u[tid * dz + d], 0.f, 1.f);
      lpz[tid] = s;
    }
    // ---------------------------------------------------------------- F6-F8: decoder and log p(x|z)
    outer_gemm<8, 2, false>(p.o_zT, FR, p.o_Wp + p.dec0W, H, 1, FR, H, dz, 0, 0,
                             epi_store(p.o_hd, ldh, p.o_Wp + p.dec0b, 1));
    __syncthreads();
    thin_gemm(p.o_hd, ldh, p.o_Wp + p.dec1W, 2 * dx, 1, H, 2 * dx, p.o_pd, p.ldpd, p.o_Wp + p.dec1b, 0, p.o_scr);
    __syncthreads();
    if (tid < FR) {
      const int r = tid;
      float s = 0.f;
      for (int d = 0; d < dx; ++d)
        s += normal_lp(xs[r * p.ldx + d], pd[r * p.ldpd + d], softplus_tf(pd[r * p.ldpd + dx + d]));
      // tile sums in row order on warp 0 (deterministic), optional per-row outputs
      const bool ok = r < nr;
      const float a = warp_sum(ok ? lq[r] - lpz[r] : 0.f), c = warp_sum(ok ? -s : 0.f);
      cta_kl += a;
      cta_nll += c;
      if (ok) {
        if (p.logq) p.logq[row0 + r] = lq[r];
        if (p.logpz) p.logpz[row0 + r] = lpz[r];
        if (p.logpx) p.logpx[row0 + r] = s;
        if (p.z)
          for (int d = 0; d < dz; ++d) p.z[(row0 + r) * dz + d] = zR[r * p.ldz + d];
      }
    }
    if (!BWD) {
      __syncthreads();
      if (nb > 0 && more) {
        stage_block(p, nb - 1, false);
        pending = true;
      }
      continue;
    }
    // ================================================================ backward
    const float g_logpx = -invB, g_logq = p.klw * invB, g_logpz = -p.klw * invB;
    {
      // B1: decoder head, d/d params of g * sum_d log N(x_d; loc_d, softplus(raw_d))
      const int gpd = p.o_scr;                  // [FR][ldpd]
      const int gpdT = gpd + FR * p.ldpd;       // [2 dx][FR]
      const int ghd = gpdT + 2 * dx * FR;       // [FR][ldh]
      for (int e = tid; e < FR * dx; e += FT) {
        const int r = e / dx, d = e - r * dx;
        const float g = r < nr ? g_logpx : 0.f;
        const float loc = pd[r * p.ldpd + d], rawp = pd[r * p.ldpd + dx + d];
        const float sc = softplus_tf(rawp);
        const float uu = xs[r * p.ldx + d] / sc - loc / sc;
        const float g1 = g * (uu / sc), g2 = g * ((uu * uu - 1.f) / sc) * sigmoidf_(rawp);
        sm[gpd + r * p.ldpd + d] = g1;
        sm[gpd + r * p.ldpd + dx + d] = g2;
        sm[gpdT + d * FR + r] = g1;
        sm[gpdT + (dx + d) * FR + r] = g2;
      }
      __syncthreads();
      // B2: [W; b] gradient of dec.1:  g[k][n] = sum_r hd[r][k] gpd[r][n]   (row k = H is the bias: hd[r][H] = 1)
      outer_gemm<8, 1, false>(gpd, p.ldpd, p.o_hd, ldh, 1, 2 * dx, H + 1, FR, 0, 0, epi_grad(gp + p.dec1W, 1, 2 * dx, first));
      // B3: g_hd = (gpd W^T) * relu'
      outer_gemm<8, 2, false>(gpdT, FR, p.o_Wp + p.dec1W, 1, 2 * dx, FR, H, 2 * dx, 0, 0,
                               epi_mask(1, ghd, ldh, p.o_hd, ldh));
      __syncthreads();
      // B4: [W; b] gradient of dec.0;  B5: g_z = ghd W^T
      outer_gemm<4, 1, false>(p.o_zR, p.ldz, ghd, ldh, 1, dz + 1, H, FR, 0, 0, epi_grad(gp + p.dec0W, H, 1, first));
      thin_gemm(ghd, ldh, p.o_Wp + p.dec0W, 1, H, H, dz, p.o_gz, dz, -1, 0, p.o_tsc);
      __syncthreads();
    }
    // ---------------------------------------------------------------- B6: prior
    int gcur = p.o_gua, gnxt = p.o_gub;
    if (nb == 0) {
      for (int e = tid; e < FR * dz; e += FT) gz[e] += -((e / dz) < nr ? g_logpz : 0.f) * u[e];
      __syncthreads();
    } else {
      for (int e = tid; e < FR * dz; e += FT) sm[gcur + e] = -((e / dz) < nr ? g_logpz : 0.f) * u[e];
      const int grawR = p.o_scr;               // [FR][ldrm]
      const int grawT = grawR + FR * ldrm;     // [ldrm][FR]
      const int gpre = grawT + ldrm * FR;      // [FR][ldfp]
      for (int e = tid; e < FR; e += FT) sm[p.o_hid + e * p.ldf + fh] = 1.f;  // bias row of the heads weight gradient
      __syncthreads();
#pragma unroll 1
      for (int i = 0; i < nb; ++i) {
        const FBlk& fb = p.blk[i];
        // the spline's reverse mode occupies warps 0-7 (one octet per row); warps 8-15 recompute the hidden layer
        if ((tid >> 5) < FW / 2)
          spline_backward(p, i, gcur, gnxt, nr, g_logpz);
        else
          hidden_layer(p, i, true, FW / 2, FW / 2);
        __syncthreads();
        // [hW; hb] gradient: g[k][n] = sum_r hid[r][k] graw[r][n]   (row k = fh is the bias)
        outer_gemm<8, 3, false>(p.o_hid, p.ldf, grawR, ldrm, 1, fh + 1, fb.ldr, FR, 0, 0,
                                 epi_grad(gp + fb.off_hW, fb.ldr, 1, first));
        // g_pre = (graw hW^T) * tanh'
        outer_gemm<4, 4, true>(grawT, FR, p.o_W, p.ldwt, 1, FR, fh, fb.ldr, gpre, p.ldfp,
                               epi_mask(2, gpre, p.ldfp, p.o_hid, p.ldf));
        __syncthreads();
        // the staging buffer is free again: next block's transposed weights, or the next tile's first forward block
        if (i + 1 < nb || more) {
          stage_block(p, i + 1 < nb ? i + 1 : nb - 1, i + 1 < nb);
          pending = true;
        }
        // [d1W; d1b] gradient and the conditioner-input gradient
        outer_gemm<4, 1, false>(p.o_cond, p.ldc, gpre, p.ldfp, 1, fb.cin + 1, fh, FR, 0, 0,
                                epi_grad(gp + fb.off_d1W, fh, 1, first));
        if (fb.nc > 0) thin_gemm(gpre, p.ldfp, p.o_B + (i & 1) * p.sb, 1, fh, fh, fb.nc, gnxt + fb.cs0, dz, -1, 1, p.o_tsc);
        if (pending && i + 1 < nb) {  // (the next tile's first block is waited for at the top of the tile loop)
          mbar_wait(bar, phase);
          phase ^= 1;
          pending = false;
        }
        __syncthreads();
        const int t = gcur;
        gcur = gnxt;
        gnxt = t;
      }
    }
    {
      // B8: encoder head (explicit z path + parameter path + reparameterisation z = eps s + loc)
      const int gpe = p.o_scr;                  // [FR][ldpe]
      const int gpeT = gpe + FR * p.ldpe;       // [2 dz][FR]
      const int ghe = gpeT + 2 * dz * FR;       // [FR][ldh]
      for (int e = tid; e < FR * dz; e += FT) {
        const int r = e / dz, d = e - r * dz;
        const float gq = r < nr ? g_logq : 0.f;
        const float loc = pe[r * p.ldpe + d], rawp = pe[r * p.ldpe + dz + d];
        const float sc = softplus_tf(rawp);
        const float zz = zR[r * p.ldz + d];
        const float uu = zz / sc - loc / sc;
        const float gzz = gz[e] + (nb > 0 ? sm[gcur + e] : 0.f) + gq * (-uu / sc);
        const float g1 = gq * (uu / sc) + gzz;
        const float g2 = (gq * ((uu * uu - 1.f) / sc) + gzz * epsS[e]) * sigmoidf_(rawp);
        sm[gpe + r * p.ldpe + d] = g1;
        sm[gpe + r * p.ldpe + dz + d] = g2;
        sm[gpeT + d * FR + r] = g1;
        sm[gpeT + (dz + d) * FR + r] = g2;
      }
      __syncthreads();
      outer_gemm<4, 1, false>(gpe, p.ldpe, p.o_he, ldh, 1, 2 * dz, H + 1, FR, 0, 0, epi_grad(gp + p.enc1W, 1, 2 * dz, first));
      outer_gemm<8, 2, false>(gpeT, FR, p.o_Wp + p.enc1W, 1, 2 * dz, FR, H, 2 * dz, 0, 0,
                               epi_mask(1, ghe, ldh, p.o_he, ldh));
      __syncthreads();
      outer_gemm<8, 1, false>(p.o_xs, p.ldx, ghe, ldh, 1, dx + 1, H, FR, 0, 0, epi_grad(gp + p.enc0W, H, 1, first));
      __syncthreads();
    }
    first = false;
  }
  if (tid == 0) {
    p.spart[2 * blockIdx.x] = cta_kl;
    p.spart[2 * blockIdx.x + 1] = cta_nll;
  }
}

// grad[i] = sum_c gpart[c][i] in a fixed order (16 partial groups per block, four independent running sums each,
// combined in a fixed order: deterministic for a given grid), and the three loss scalars.
constexpr int kFinGroups = 16;
struct AdamArgs {
  float *theta, *m, *v;  // theta == NULL: no optimiser step
  float lr_t, one_minus_b1, one_minus_b2, eps;
};
__global__ void __launch_bounds__(32 * kFinGroups) fused_finish_kernel(const float* __restrict__ gpart, int n_part, int P,
                                                                       float* __restrict__ grad,
                                                                       const float* __restrict__ spart, int64_t B,
                                                                       float klw, float* __restrict__ scalars,
                                                                       const AdamArgs ad) {
  __shared__ float sh[kFinGroups][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  // programmatic dependent launch (elbo_tcf.cu has the same pair): this kernel may be scheduled while the tile kernel is
  // still running and reads its partials behind this; returns at once after an ordinary launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (grad) {
    const int i = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (i < P) {
      const int per = (n_part + kFinGroups - 1) / kFinGroups;
      const int c0 = grp * per, c1 = min(n_part, c0 + per);
      const float* g = gpart + i;
      int c = c0;
      for (; c + 4 <= c1; c += 4) {
        s0 += g[(size_t)c * P];
        s1 += g[(size_t)(c + 1) * P];
        s2 += g[(size_t)(c + 2) * P];
        s3 += g[(size_t)(c + 3) * P];
      }
      for (; c < c1; ++c) s0 += g[(size_t)c * P];
    }
    sh[grp][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (grp == 0 && i < P) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < kFinGroups; ++g) t += sh[g][lane];
      grad[i] = t;
      if (ad.theta) {  // Keras Adam on the flat buffer, same arithmetic as adam_kernel (adam.cu)
        const float mi = ad.m[i] + (t - ad.m[i]) * ad.one_minus_b1;
        const float vi = ad.v[i] + (t * t - ad.v[i]) * ad.one_minus_b2;
        ad.m[i] = mi;
        ad.v[i] = vi;
        ad.theta[i] = ad.theta[i] - ad.lr_t * mi / (sqrtf(vi) + ad.eps);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && scalars) {
    float sa = 0.f, sc = 0.f;
    for (int c = 0; c < n_part; ++c) {
      sa += spart[2 * c];
      sc += spart[2 * c + 1];
    }
    const float kl = sa / (float)B, nll = sc / (float)B;
    scalars[0] = nll + klw * kl;
    scalars[1] = nll;
    scalars[2] = kl;
  }
}

// ------------------------------------------------------------------------------------------------ host side
static inline int r4(int v) { return (v + 3) & ~3; }

vms_status fused_create(vms_elbo_plan_s* pl) {
  const vms_elbo_desc& d = pl->d;
  pl->fused = nullptr;
  // shapes the fused kernel is written for; anything else runs the unfused plan
  if (d.dx > kMaxThin / 2 || d.dz > kMaxThin / 2 || d.num_blocks > kMaxBlocks) return VMS_OK;
  if (d.num_blocks > 0 && (d.num_bins > 32 || d.num_bins % 4 != 0)) return VMS_OK;
  FusedCfg* f = new FusedCfg();
  FusedParams& p = f->p;
  memset(&p, 0, sizeof(p));
  p.dx = d.dx; p.dz = d.dz; p.hidden = d.hidden; p.nb = d.num_blocks; p.K = d.num_bins; p.fh = d.num_blocks ? d.flow_hidden : 4;
  p.bin_min = d.bin_min;
  p.scale = (float)((double)d.bin_max - (double)d.bin_min - (double)d.num_bins * 1e-2);  // flows.py:92
  p.klw = d.kl_weight;
  const Offsets& o = pl->off;
  p.P = (int)o.total;
  p.enc0W = (int)o.enc0W; p.enc0b = (int)o.enc0b; p.enc1W = (int)o.enc1W; p.enc1b = (int)o.enc1b;
  p.dec0W = (int)o.dec0W; p.dec0b = (int)o.dec0b; p.dec1W = (int)o.dec1W; p.dec1b = (int)o.dec1b;
  int max_ldr = 4, max_cin = 1;
  for (int i = 0; i < d.num_blocks; ++i) {
    const FlowBlock& b = pl->blocks[i];
    FBlk& fb = p.blk[i];
    fb.cs0 = b.cs0; fb.nc = b.cs1 - b.cs0; fb.ts0 = b.ts0; fb.dt = b.dt; fb.cin = b.cin; fb.ldr = b.ldr;
    fb.off_d1W = (int)b.off_d1W; fb.off_d1b = (int)b.off_d1b; fb.off_hW = (int)b.off_hW; fb.off_hb = (int)b.off_hb;
    max_ldr = b.ldr > max_ldr ? b.ldr : max_ldr;
    max_cin = b.cin > max_cin ? b.cin : max_cin;
  }
  if (max_cin > kMaxThin) { delete f; return VMS_OK; }
  p.ldx = r4(d.dx + 1); p.ldz = r4(d.dz + 1); p.ldh = (d.hidden + 1) | 1; p.ldf = r4(p.fh + 1); p.ldfp = p.fh | 1;
  p.ldpe = r4(2 * d.dz); p.ldpd = r4(2 * d.dx); p.ldc = r4(max_cin + 1); p.ldrm = r4(max_ldr);
  p.ldwt = r4(p.fh);
  p.sb = r4(max_cin * p.fh + p.fh + p.ldrm);
  p.n_mlp = (int)o.dec1b + 2 * d.dx;
  int off = 0;
  auto take = [&](int n) { int o0 = off; off += r4(n); return o0; };
  p.o_xs = take(FR * p.ldx); p.o_xT = take(d.dx * FR); p.o_eps = take(FR * d.dz);
  p.o_zR = take(FR * p.ldz); p.o_zT = take(d.dz * FR);
  p.o_he = take(FR * p.ldh); p.o_hd = take(FR * p.ldh);
  p.o_pe = take(FR * p.ldpe); p.o_pd = take(FR * p.ldpd);
  p.o_u = take((d.num_blocks + 1) * FR * d.dz); p.o_lp = take(3 * FR);
  p.o_hid = take(p.fh * FR > FR * p.ldf ? p.fh * FR : FR * p.ldf); p.o_cond = take(FR * p.ldc);
  p.o_raw = take(d.num_blocks * FR * p.ldrm);
  const int w_fwd = p.fh * p.ldrm, w_bwd = max_ldr * p.ldwt;
  p.o_W = take(d.num_blocks ? (w_fwd > w_bwd ? w_fwd : w_bwd) : 4);
  p.o_Wp = take(p.n_mlp); p.o_B = take(2 * p.sb);
  p.o_gz = take(FR * d.dz); p.o_gua = take(FR * d.dz); p.o_gub = take(FR * d.dz);
  const int scr_dec = FR * p.ldpd + 2 * d.dx * FR + FR * p.ldh;
  const int scr_enc = FR * p.ldpe + 2 * d.dz * FR + FR * p.ldh;
  const int scr_flow = d.num_blocks ? 2 * FR * p.ldrm + FR * p.ldfp : 0;
  int scr = scr_dec > scr_enc ? scr_dec : scr_enc;
  scr = scr_flow > scr ? scr_flow : scr;
  // thin_gemm scratch: the forward thin layers (2 dx / 2 dz outputs) borrow scr, the backward ones (dz / nc outputs)
  // have a small region of their own
  const int thin_fwd = FW * r4(2 * (d.dx > d.dz ? d.dx : d.dz)) * FR;
  scr = thin_fwd > scr ? thin_fwd : scr;
  p.o_tsc = take(FW * r4(d.dz > max_cin ? d.dz : max_cin) * FR);
  p.o_scr = take(scr);
  p.o_bar = take(4);
  off += 64;  // slab reads may run a few floats past the last row of an operand
  f->smem_bytes = (size_t)off * sizeof(float);
  if (f->smem_bytes > (size_t)max_smem_optin()) { delete f; return VMS_OK; }
  f->max_grid = sm_count();
  cudaError_t e = cudaFuncSetAttribute(elbo_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_bytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(elbo_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem_bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete f;
    return VMS_OK;
  }
  void* g = nullptr;
  void* s = nullptr;
  if (cudaMalloc(&g, (size_t)f->max_grid * p.P * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&s, (size_t)f->max_grid * 2 * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    if (g) cudaFree(g);
    delete f;
    set_error("elbo_plan_create: cudaMalloc of the fused partial-gradient buffers failed");
    return VMS_ERR_CUDA;
  }
  f->gpart = (float*)g;
  f->spart = (float*)s;
  // packed flow-block weights: [hW natural | hW transposed | small], each part a multiple of 16 bytes
  p.pack_wt = p.fh * p.ldrm;
  p.pack_small = p.pack_wt + r4(max_ldr * p.ldwt);
  p.pack_stride = p.pack_small + p.sb;
  f->pack = nullptr;
  if (d.num_blocks > 0) {
    void* pk = nullptr;
    if (cudaMalloc(&pk, (size_t)d.num_blocks * p.pack_stride * sizeof(float)) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(g);
      cudaFree(s);
      delete f;
      set_error("elbo_plan_create: cudaMalloc of the packed-weights buffer failed");
      return VMS_ERR_CUDA;
    }
    f->pack = (float*)pk;
    cudaMemset(pk, 0, (size_t)d.num_blocks * p.pack_stride * sizeof(float));
  }
  pl->fused = f;
  return VMS_OK;
}

vms_status fused_set_timing(vms_elbo_plan_s* pl, int max_launches) {
  FusedCfg* f = pl->fused;
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

vms_status fused_kernel_ms(vms_elbo_plan_s* pl, double* total_ms, int* launches) {
  FusedCfg* f = pl->fused;
  *total_ms = 0.0;
  *launches = 0;
  if (!f) return VMS_OK;
  for (size_t i = 0; i < f->ev_used; ++i) {
    float ms = 0.f;
    VMS_CUDA(cudaEventSynchronize(f->ev[i].second));
    VMS_CUDA(cudaEventElapsedTime(&ms, f->ev[i].first, f->ev[i].second));
    *total_ms += ms;
  }
  *launches = (int)f->ev_used;
  f->ev_used = 0;
  return VMS_OK;
}

void fused_destroy(vms_elbo_plan_s* pl) {
  if (!pl->fused) return;
  for (auto& e : pl->fused->ev) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  cudaFree(pl->fused->gpart);
  cudaFree(pl->fused->pack);
  cudaFree(pl->fused->spart);
  delete pl->fused;
  pl->fused = nullptr;
}

vms_status fused_run(vms_elbo_plan_s* pl, const float* theta, const float* x, const float* eps, int64_t B, bool backward,
                     float* z, float* logq, float* logpz, float* logpx, float* grad, float* scalars, cudaStream_t st,
                     const FusedAdam* adam) {
  FusedCfg* f = pl->fused;
  FusedParams p = f->p;
  p.B = B;
  p.n_tiles = (int)((B + FR - 1) / FR);
  p.theta = theta; p.x = x; p.eps = eps;
  p.z = z; p.logq = logq; p.logpz = logpz; p.logpx = logpx;
  p.gpart = f->gpart; p.spart = f->spart; p.pack = f->pack;
  if (p.nb > 0) {
    const int per_blk = p.fh * p.ldrm + p.sb;  // >= source elements of any block
    prepack_kernel<<<dim3((per_blk + 255) / 256, p.nb), 256, 0, st>>>(p, f->pack);
    VMS_LAUNCH_CHECK("prepack_kernel");
  }
  const int grid = p.n_tiles < f->max_grid ? p.n_tiles : f->max_grid;
  const bool timed = f->timing && f->ev_used < f->ev.size() && (pl->timing_calls++ % (unsigned)pl->timing_every) == 0u;
  if (timed) VMS_CUDA(cudaEventRecord(f->ev[f->ev_used].first, st));
  if (backward)
    elbo_fused_kernel<true><<<grid, FT, f->smem_bytes, st>>>(p);
  else
    elbo_fused_kernel<false><<<grid, FT, f->smem_bytes, st>>>(p);
  VMS_LAUNCH_CHECK("elbo_fused_kernel");
  if (timed) VMS_CUDA(cudaEventRecord(f->ev[f->ev_used++].second, st));
  float* sc = scalars ? scalars : pl->scalars;
  const int nblk = backward ? (p.P + 31) / 32 : 1;
  AdamArgs ad = {};
  if (adam && backward) {
    ad.theta = adam->theta; ad.m = adam->m; ad.v = adam->v;
    ad.lr_t = adam->lr_t; ad.one_minus_b1 = adam->one_minus_b1; ad.one_minus_b2 = adam->one_minus_b2; ad.eps = adam->eps;
  }
  {
    static cudaLaunchAttribute attr[1];
    static int pdl = -1;
    if (pdl < 0) {
      const char* e = getenv("VMS_TCF_PDL");  // 0: ordinary launches (also for the tensor-core plan's finish kernels)
      pdl = (e && e[0] == '0') ? 0 : 1;
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nblk);
    cfg.blockDim = dim3(32 * kFinGroups);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    VMS_CUDA(cudaLaunchKernelEx(&cfg, fused_finish_kernel, (const float*)f->gpart, grid, p.P, backward ? grad : (float*)nullptr,
                                (const float*)f->spart, B, p.klw, sc, ad));
  }
  VMS_LAUNCH_CHECK("fused_finish_kernel");
  return VMS_OK;
}

}  // namespace vms
