// rqs.cu -- K1: fused spline-bin search + rational-quadratic-spline evaluation with per-element log-det (sm_100a).
//
// Replaces flows.py:86-101 / :394-409 (softmax*scale+1e-2, softplus+1e-2 "activations") and the
// tfp.bijectors.RationalQuadraticSpline object built at flows.py:204-207 / :512-515 (forward, inverse, fldj,
// ildj = -fldj(inverse)), plus TF autodiff through both (backward kernel).
//
// Design (B200).  The op is HBM-bound: (3K-1) raw logits in, 2 scalars out per element (392 B @ K=32), so the kernel
// is built around coalesced 16-byte loads and a small instruction count per element:
//   * 8 lanes ("octet") cooperate on one element; each lane owns K/8 consecutive bins (4 bins @ K<=32: ONE float4
//     load per logit array per lane, 128 contiguous bytes per element, 4 elements per warp instruction);
//   * softmax max / sum and the cumulative sum are 3-step shuffles inside the octet (not 5-step warp scans), the
//     per-lane part is a 4-iteration sequential loop;
//   * no shared memory and ~64 registers => full occupancy, memory-level parallelism comes from resident warps;
//   * exactly one lane finds the bin; it evaluates the spline and writes y / log-det; it alone reads the TWO knot
//     slopes the element needs (the other K-3 raw slopes never leave HBM);
//   * prefix sums and knot positions are accumulated in float64 (B200 runs FP64 at half the FP32 rate): float32 knots
//     (what TFP computes) carry ~ulp(range) error that narrow bins amplify to 1e-4 in the log-det (parity).
#include "rqs_device.cuh"

namespace {

using namespace vms::rqsdev;
constexpr int kThreads = 256;

struct RqsParams {
  int64_t n_rows;
  int n_dims;  // Dt: transformed dims per row, processed sequentially by the row's octet
  int K;
  float bin_min, bin_max, scale;  // scale = bin_max - bin_min - K*1e-2 (flows.py:92)
  const float* v_in;  int64_t ld_in;
  const float* raw_w; int64_t ld_w;
  const float* raw_h; int64_t ld_h;
  const float* raw_s; int64_t ld_s;
  float* v_out; int64_t ld_out;
  float* ldj;      // [n_rows, Dt] or null
  float* ldj_sum;  // [n_rows] or null
  int accumulate;
  int inverse_dir;
  // backward only
  const float* g_out; int64_t ld_g_out;
  const float* g_ldj_sum;
  float* g_in; int64_t ld_g_in;
  float* g_w; int64_t ld_gw;
  float* g_h; int64_t ld_gh;
  float* g_s; int64_t ld_gs;
};

template <int BPL, bool VEC>
__global__ void __launch_bounds__(kThreads) rqs_apply_kernel(const RqsParams p) {
  const int j = threadIdx.x & (kOct - 1);
  const int64_t oct0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kOct;
  const int64_t n_oct = (int64_t)gridDim.x * (kThreads / kOct);
  const int K = p.K;
  const bool inv = p.inverse_dir != 0;
  const int64_t n_iter = (p.n_rows + n_oct - 1) / n_oct;  // uniform trip count: shuffles need all lanes present
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row = oct0 + it * n_oct;
    const bool active = row < p.n_rows;
    const int64_t r = active ? row : p.n_rows - 1;  // inactive octets redo the last row, writes predicated off
    float ldj_acc = 0.f;
    for (int d = 0; d < p.n_dims; ++d) {
      const float v = __ldg(p.v_in + r * p.ld_in + d);
      float out, ldj, ldj_all;
      bool writer;
      octet_apply<BPL, VEC, true>(p.raw_w + r * p.ld_w + (int64_t)d * K, p.raw_h + r * p.ld_h + (int64_t)d * K,
                                  p.raw_s + r * p.ld_s + (int64_t)d * (K - 1), v, j, K, inv, p.bin_min, p.scale, out,
                                  ldj, ldj_all, writer);
      if (writer && active) {
        p.v_out[r * p.ld_out + d] = out;
        if (p.ldj) p.ldj[r * p.n_dims + d] = ldj;
      }
      ldj_acc += ldj_all;
    }
    if (p.ldj_sum && j == 0 && active) {
      float* dst = p.ldj_sum + r;
      *dst = p.accumulate ? *dst + ldj_acc : ldj_acc;
    }
  }
}

template <int BPL, bool VEC>
__global__ void __launch_bounds__(kThreads) rqs_backward_kernel(const RqsParams p) {
  const int j = threadIdx.x & (kOct - 1);
  const int64_t oct0 = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kOct;
  const int64_t n_oct = (int64_t)gridDim.x * (kThreads / kOct);
  const int K = p.K;
  const bool inv = p.inverse_dir != 0;
  const int64_t n_iter = (p.n_rows + n_oct - 1) / n_oct;
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row = oct0 + it * n_oct;
    const bool active = row < p.n_rows;
    const int64_t r = active ? row : p.n_rows - 1;
    const float g_ldj = p.g_ldj_sum ? __ldg(p.g_ldj_sum + r) : 0.f;
    for (int d = 0; d < p.n_dims; ++d) {
      const float v = __ldg(p.v_in + r * p.ld_in + d);
      const float g_out = __ldg(p.g_out + r * p.ld_g_out + d);
      float g_in, gw[BPL], gh[BPL], gs[BPL];
      bool writer;
      octet_backward<BPL, VEC, true>(p.raw_w + r * p.ld_w + (int64_t)d * K, p.raw_h + r * p.ld_h + (int64_t)d * K,
                                     p.raw_s + r * p.ld_s + (int64_t)d * (K - 1), v, g_out, g_ldj, j, K, inv, p.bin_min,
                                     p.scale, g_in, writer, gw, gh, gs);
      if (active) {
        if (writer) p.g_in[r * p.ld_g_in + d] = g_in;
        store_bins<BPL, VEC>(p.g_w + r * p.ld_gw + (int64_t)d * K, j, K, gw);
        store_bins<BPL, VEC>(p.g_h + r * p.ld_gh + (int64_t)d * K, j, K, gh);
        store_bins<BPL, false>(p.g_s + r * p.ld_gs + (int64_t)d * (K - 1), j, K - 1, gs);
      }
    }
  }
}

inline bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

vms_status configure(RqsParams& p, int& grid) {
  VMS_REQUIRE(p.K >= 2 && p.K <= 64, VMS_ERR_INVALID_ARG, "num_bins must be in [2, 64], got %d", p.K);
  VMS_REQUIRE(p.n_dims >= 1 && p.n_dims <= 4096, VMS_ERR_INVALID_ARG, "n_dims must be in [1, 4096], got %d", p.n_dims);
  VMS_REQUIRE(p.n_rows >= 0, VMS_ERR_INVALID_ARG, "n_rows < 0");
  VMS_REQUIRE(p.bin_max > p.bin_min, VMS_ERR_INVALID_ARG, "bin_range must be increasing");
  // flows.py:92: Python float arithmetic, then one cast to float32
  p.scale = (float)((double)p.bin_max - (double)p.bin_min - (double)p.K * 1e-2);
  VMS_REQUIRE(p.scale > 0.f, VMS_ERR_INVALID_ARG, "bin_range too narrow for %d bins", p.K);
  const int64_t n_blocks = (p.n_rows + (kThreads / kOct) - 1) / (kThreads / kOct);
  const int64_t cap = 16LL * vms::sm_count();
  grid = (int)(n_blocks < cap ? n_blocks : cap);
  return VMS_OK;
}

bool vector_ok(const RqsParams& p, bool backward) {
  if (p.K % 4) return false;
  bool ok = aligned16(p.raw_w) && aligned16(p.raw_h) && p.ld_w % 4 == 0 && p.ld_h % 4 == 0;
  if (backward) ok = ok && aligned16(p.g_w) && aligned16(p.g_h) && p.ld_gw % 4 == 0 && p.ld_gh % 4 == 0;
  return ok;
}

#define VMS_RQS_DISPATCH(KERNEL, name)                                                    \
  do {                                                                                    \
    if (p.n_rows == 0) return VMS_OK;                                                     \
    if (p.K <= 32) {                                                                      \
      if (vec) KERNEL<4, true><<<grid, kThreads, 0, st>>>(p);                             \
      else KERNEL<4, false><<<grid, kThreads, 0, st>>>(p);                                \
    } else {                                                                              \
      if (vec) KERNEL<8, true><<<grid, kThreads, 0, st>>>(p);                             \
      else KERNEL<8, false><<<grid, kThreads, 0, st>>>(p);                                \
    }                                                                                     \
    VMS_LAUNCH_CHECK(name);                                                               \
    return VMS_OK;                                                                        \
  } while (0)

vms_status run_apply(RqsParams p, cudaStream_t st) {
  int grid;
  vms_status s = configure(p, grid);
  if (s) return s;
  const bool vec = vector_ok(p, false);
  VMS_RQS_DISPATCH(rqs_apply_kernel, "rqs_apply_kernel");
}
vms_status run_backward(RqsParams p, cudaStream_t st) {
  int grid;
  vms_status s = configure(p, grid);
  if (s) return s;
  const bool vec = vector_ok(p, true);
  VMS_RQS_DISPATCH(rqs_backward_kernel, "rqs_backward_kernel");
}

RqsParams from_args(const vms_rqs_args& a) {
  RqsParams p = {};
  p.n_rows = a.n_rows; p.n_dims = a.n_dims; p.K = a.num_bins;
  p.bin_min = a.bin_min; p.bin_max = a.bin_max;
  p.v_in = a.v_in; p.ld_in = a.ld_in;
  p.raw_w = a.raw_w; p.ld_w = a.ld_w;
  p.raw_h = a.raw_h; p.ld_h = a.ld_h;
  p.raw_s = a.raw_s; p.ld_s = a.ld_s;
  p.v_out = a.v_out; p.ld_out = a.ld_out;
  p.ldj = a.ldj; p.ldj_sum = a.ldj_sum;
  p.accumulate = a.accumulate; p.inverse_dir = a.inverse_dir;
  return p;
}

}  // namespace

namespace vms {
vms_status rqs_prepare() { return VMS_OK; }  // no opt-in shared memory any more; kept for elbo.cu
// rqs_stream.cu: thread-per-element streaming kernels for the contiguous, aligned form with K in {20, 32}
bool rqs_stream_try(const float* v, const float* rw, const float* rh, const float* rs, int64_t n, int K, float bin_min,
                    float bin_max, int inverse_dir, float* out, float* ldj, const float* g_out, const float* g_ldj,
                    float* g_in, float* g_w, float* g_h, float* g_s, bool backward, cudaStream_t st, vms_status* status);
}  // namespace vms

extern "C" {

vms_status vms_rqs_apply(const vms_rqs_args* a, vms_stream stream) {
  VMS_REQUIRE(a, VMS_ERR_INVALID_ARG, "args is NULL");
  VMS_REQUIRE(a->n_rows == 0 || (a->v_in && a->raw_w && a->raw_h && a->raw_s && a->v_out), VMS_ERR_INVALID_ARG,
              "vms_rqs_apply: NULL tensor pointer");
  return run_apply(from_args(*a), vms::as_stream(stream));
}

vms_status vms_rqs_apply_backward(const vms_rqs_bwd_args* a, vms_stream stream) {
  VMS_REQUIRE(a, VMS_ERR_INVALID_ARG, "args is NULL");
  RqsParams p = from_args(a->fwd);
  VMS_REQUIRE(p.n_rows == 0 || (p.v_in && p.raw_w && p.raw_h && p.raw_s && a->g_out && a->g_in && a->g_raw_w &&
                                a->g_raw_h && a->g_raw_s),
              VMS_ERR_INVALID_ARG, "vms_rqs_apply_backward: NULL tensor pointer");
  p.g_out = a->g_out; p.ld_g_out = a->ld_g_out;
  p.g_ldj_sum = a->g_ldj_sum;
  p.g_in = a->g_in; p.ld_g_in = a->ld_g_in;
  p.g_w = a->g_raw_w; p.ld_gw = a->ld_gw;
  p.g_h = a->g_raw_h; p.ld_gh = a->ld_gh;
  p.g_s = a->g_raw_s; p.ld_gs = a->ld_gs;
  return run_backward(p, vms::as_stream(stream));
}

static vms_status contiguous(const float* v, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                             float lo, float hi, float* out, float* ldj, int inverse_dir, vms_stream stream) {
  VMS_REQUIRE(n >= 0, VMS_ERR_INVALID_ARG, "n_elem < 0");
  VMS_REQUIRE(n == 0 || (v && rw && rh && rs && out), VMS_ERR_INVALID_ARG, "vms_rqs: NULL tensor pointer");
  VMS_REQUIRE(K >= 2 && K <= 64, VMS_ERR_INVALID_ARG, "num_bins must be in [2, 64], got %d", K);
  VMS_REQUIRE(hi > lo, VMS_ERR_INVALID_ARG, "bin_range must be increasing");
  {
    vms_status s = VMS_OK;
    if (vms::rqs_stream_try(v, rw, rh, rs, n, K, lo, hi, inverse_dir, out, ldj, nullptr, nullptr, nullptr, nullptr, nullptr,
                            nullptr, false, vms::as_stream(stream), &s))
      return s;
  }
  vms_rqs_args a = {};
  a.n_rows = n; a.n_dims = 1; a.num_bins = K; a.bin_min = lo; a.bin_max = hi;
  a.v_in = v; a.ld_in = 1;
  a.raw_w = rw; a.ld_w = K; a.raw_h = rh; a.ld_h = K; a.raw_s = rs; a.ld_s = K - 1;
  a.v_out = out; a.ld_out = 1;
  a.ldj = ldj; a.ldj_sum = nullptr; a.accumulate = 0; a.inverse_dir = inverse_dir;
  return run_apply(from_args(a), vms::as_stream(stream));
}

vms_status vms_rqs_forward(const float* x, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                           float lo, float hi, float* y, float* ldj, vms_stream stream) {
  return contiguous(x, rw, rh, rs, n, K, lo, hi, y, ldj, 0, stream);
}
vms_status vms_rqs_inverse(const float* y, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                           float lo, float hi, float* x, float* ldj, vms_stream stream) {
  return contiguous(y, rw, rh, rs, n, K, lo, hi, x, ldj, 1, stream);
}

vms_status vms_rqs_backward(const float* v_in, const float* rw, const float* rh, const float* rs, int64_t n, int K,
                            float lo, float hi, int inverse_dir, const float* g_out, const float* g_ldj, float* g_in,
                            float* g_rw, float* g_rh, float* g_rs, vms_stream stream) {
  VMS_REQUIRE(n >= 0, VMS_ERR_INVALID_ARG, "n_elem < 0");
  VMS_REQUIRE(n == 0 || (v_in && rw && rh && rs && g_out && g_in && g_rw && g_rh && g_rs), VMS_ERR_INVALID_ARG,
              "vms_rqs_backward: NULL tensor pointer");
  VMS_REQUIRE(K >= 2 && K <= 64, VMS_ERR_INVALID_ARG, "num_bins must be in [2, 64], got %d", K);
  VMS_REQUIRE(hi > lo, VMS_ERR_INVALID_ARG, "bin_range must be increasing");
  {
    vms_status s = VMS_OK;
    if (vms::rqs_stream_try(v_in, rw, rh, rs, n, K, lo, hi, inverse_dir, nullptr, nullptr, g_out, g_ldj, g_in, g_rw, g_rh,
                            g_rs, true, vms::as_stream(stream), &s))
      return s;
  }
  RqsParams p = {};
  p.n_rows = n; p.n_dims = 1; p.K = K; p.bin_min = lo; p.bin_max = hi;
  p.v_in = v_in; p.ld_in = 1;
  p.raw_w = rw; p.ld_w = K; p.raw_h = rh; p.ld_h = K; p.raw_s = rs; p.ld_s = K - 1;
  p.inverse_dir = inverse_dir;
  p.g_out = g_out; p.ld_g_out = 1;
  p.g_ldj_sum = g_ldj;
  p.g_in = g_in; p.ld_g_in = 1;
  p.g_w = g_rw; p.ld_gw = K; p.g_h = g_rh; p.ld_gh = K; p.g_s = g_rs; p.ld_gs = K - 1;
  return run_backward(p, vms::as_stream(stream));
}

}  // extern "C"
