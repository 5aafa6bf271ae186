/*
 * vms_b200.h -- C ABI of libvms_b200.so: sm_100a kernels for the vaemolsim hot path.
 *
 * The reference (Monroe-Molecular-Simulation-Group/vae-mol-sim) is pure Python over TensorFlow /
 * TensorFlow-Probability and has no FFI of its own.  The seam this library plugs into is the set of
 * object protocols its layers return and its losses / models / MCMC driver consume (Bijector,
 * Distribution, Tensor-with-.numpy()).  Each entry point below names the reference call site whose
 * arithmetic it replaces (paths are relative to the reference root, `vaemolsim/...`).
 *
 * Conventions
 *   - every pointer marked "device" is CUDA device memory owned by the caller; the library never frees or
 *     retains a caller pointer past the call (plans own only what they allocate themselves);
 *   - row-major, float32 unless stated; `ld_*` are leading dimensions (row strides) in elements;
 *   - all compute entry points are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream); launch errors are returned synchronously, execution errors at the next sync;
 *   - return value: 0 = OK, >0 = vms_status code; `vms_last_error()` gives a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute call returns VMS_ERR_CUDA.
 */
#ifndef VMS_B200_H
#define VMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int vms_status;
enum {
  VMS_OK = 0,
  VMS_ERR_INVALID_ARG = 1, /* Python host maps to ValueError */
  VMS_ERR_SHAPE = 2,       /* ValueError */
  VMS_ERR_CUDA = 3,        /* RuntimeError */
  VMS_ERR_NCCL = 4,        /* RuntimeError */
  VMS_ERR_UNSUPPORTED = 5  /* NotImplementedError */
};

typedef void* vms_stream; /* cudaStream_t */
typedef void* vms_event;  /* cudaEvent_t  */

/* ------------------------------------------------------------------------------------------------ runtime */
const char* vms_last_error(void);
int vms_abi_version(void);
vms_status vms_device_count(int* count);
vms_status vms_set_device(int device);
/* out[0]=SM count, out[1]=cc major, out[2]=cc minor, out[3]=max opt-in smem/block (bytes), out[4]=L2 bytes */
vms_status vms_device_info(int device, int64_t out[5]);
vms_status vms_malloc(void** device_ptr, size_t bytes);
vms_status vms_free(void* device_ptr);
vms_status vms_malloc_host(void** pinned_ptr, size_t bytes);
vms_status vms_free_host(void* pinned_ptr);
vms_status vms_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes, vms_stream stream);
vms_status vms_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes, vms_stream stream);
vms_status vms_memcpy_d2d(void* dst_device, const void* src_device, size_t bytes, vms_stream stream);
/* strided 2-D copy (rows x width_bytes), device to device: column slices of [B, D] tensors */
vms_status vms_memcpy2d_d2d(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                            size_t rows, vms_stream stream);
vms_status vms_memset(void* device_ptr, int value, size_t bytes, vms_stream stream);
vms_status vms_stream_create(vms_stream* stream);
vms_status vms_stream_destroy(vms_stream stream);
vms_status vms_stream_synchronize(vms_stream stream);
vms_status vms_device_synchronize(void);
vms_status vms_event_create(vms_event* ev);
vms_status vms_event_destroy(vms_event ev);
vms_status vms_event_record(vms_event ev, vms_stream stream);
vms_status vms_stream_wait_event(vms_stream stream, vms_event ev); /* later work on `stream` waits for `ev` (device side) */
vms_status vms_event_synchronize(vms_event ev);
vms_status vms_event_elapsed_ms(vms_event start, vms_event stop, float* ms);
/* CUDA graphs for launch-bound host loops (the tape-based training step issues ~100 small kernels from the host): begin puts
 * the stream into capture (relaxed mode), end instantiates what was captured and reports the number of this library's kernel
 * launches in it (they did not run; vms_launch_count advances per replay), abort discards a capture, launch replays. */
vms_status vms_graph_begin_capture(vms_stream stream);
vms_status vms_graph_end_capture(vms_stream stream, void** graph_exec, int* n_kernels);
vms_status vms_graph_abort_capture(vms_stream stream);
vms_status vms_graph_launch(void* graph_exec, int n_kernels, vms_stream stream);
vms_status vms_graph_destroy(void* graph_exec);
/* counts kernels launched by this library in this process (bench.py's `gpu_launches`) */
unsigned long long vms_launch_count(void);

/* ------------------------------------------------------------------------------------- K1: RQS bijector
 * Replaces  flows.py:86-101 (`SplineBijector._bin_positions/_slopes` activations), flows.py:394-409 (same for
 * `MaskedSplineBijector`) and the `tfp.bijectors.RationalQuadraticSpline(bin_widths, bin_heights, knot_slopes,
 * range_min)` object built at flows.py:204-207 / :512-515: its `_forward`, `_inverse` and
 * `_forward_log_det_jacobian` (ildj = -fldj(inverse)).  Activations are fused: inputs are the RAW Dense / MADE
 * outputs.  One "element" = one transformed scalar with its K + K + (K-1) raw logits.
 *
 * Contiguous form (the op boundary of SURVEY 8b):  x [n_elem], raw_w/raw_h [n_elem, K], raw_s [n_elem, K-1],
 * y [n_elem], ldj [n_elem] (nullable).  2 <= K <= 64.
 */
vms_status vms_rqs_forward(const float* x, const float* raw_w, const float* raw_h, const float* raw_s,
                           int64_t n_elem, int K, float bin_min, float bin_max, float* y, float* ldj,
                           vms_stream stream);
vms_status vms_rqs_inverse(const float* y, const float* raw_w, const float* raw_h, const float* raw_s,
                           int64_t n_elem, int K, float bin_min, float bin_max, float* x, float* ldj,
                           vms_stream stream);
/* Reverse mode of the two ops above.  `inverse_dir` selects which op is differentiated.
 * v_in is that op's input (x for forward, y for inverse); g_out / g_ldj are upstream gradients of its two
 * outputs (g_ldj nullable = 0).  Outputs: g_in [n_elem], g_raw_w/h [n_elem,K], g_raw_s [n_elem,K-1].
 * tf.where semantics out of range: g_in = g_out, parameter gradients 0. */
vms_status vms_rqs_backward(const float* v_in, const float* raw_w, const float* raw_h, const float* raw_s,
                            int64_t n_elem, int K, float bin_min, float bin_max, int inverse_dir,
                            const float* g_out, const float* g_ldj, float* g_in, float* g_raw_w, float* g_raw_h,
                            float* g_raw_s, vms_stream stream);

/* Strided form used by the coupling layers (flows.py:312 RealNVP, :628-637 MAF): rows of a [B, D] tensor,
 * `n_dims` transformed columns starting at v_in / v_out, raw logits as [B, n_dims*K] row blocks inside wider
 * buffers (e.g. the fused [w|h|s] output of one conditioner GEMM).  ldj_sum [B] (nullable) receives the
 * event-summed log-det (`event_ndims=1`), added to the existing value when accumulate != 0. */
typedef struct {
  int64_t n_rows;
  int32_t n_dims;
  int32_t num_bins;
  float bin_min, bin_max;
  const float* v_in;   int64_t ld_in;
  const float* raw_w;  int64_t ld_w;
  const float* raw_h;  int64_t ld_h;
  const float* raw_s;  int64_t ld_s;
  float* v_out;        int64_t ld_out;
  float* ldj;          /* [n_rows, n_dims] contiguous, nullable */
  float* ldj_sum;      /* [n_rows], nullable */
  int32_t accumulate;  /* ldj_sum += instead of = */
  int32_t inverse_dir; /* 0: forward + fldj, 1: inverse + ildj */
} vms_rqs_args;
vms_status vms_rqs_apply(const vms_rqs_args* args, vms_stream stream);

typedef struct {
  vms_rqs_args fwd;       /* same geometry / inputs as the forward call being differentiated (outputs unused) */
  const float* g_out;  int64_t ld_g_out;  /* upstream grad of v_out */
  const float* g_ldj_sum;                 /* [n_rows] upstream grad of the event-summed ldj, nullable */
  float* g_in;         int64_t ld_g_in;
  float* g_raw_w;      int64_t ld_gw;
  float* g_raw_h;      int64_t ld_gh;
  float* g_raw_s;      int64_t ld_gs;
} vms_rqs_bwd_args;
vms_status vms_rqs_apply_backward(const vms_rqs_bwd_args* args, vms_stream stream);

/* --------------------------------------------------------------------------------- K2/K3: dense layers
 * Replaces Keras Dense at mappings.py:107-121 (`FCDeepNN.build`), flows.py:136-152 (`SplineBijector` nets) and
 * the pre-masked Dense layers of `tfp.bijectors.AutoregressiveNetwork` (flows.py:454-487, dists.py:301-305;
 * masks are baked into W by the host exactly as TFP's masked initializer + constraint do).
 *   out[B,N] = act( x[B,K] @ W[K,N] + b[N] (+ cond[B,C] @ Wc[C,N]) )      act: 0 none, 1 relu, 2 tanh
 * `ones_input` != 0 replaces x by ones([B,1]) (flows.py:184-185: empty RealNVP conditioner input).          */
enum { VMS_ACT_NONE = 0, VMS_ACT_RELU = 1, VMS_ACT_TANH = 2 };
vms_status vms_dense_forward(const float* x, int64_t ld_x, const float* W, const float* b, int64_t B, int K,
                             int N, int act, const float* cond, int64_t ld_c, const float* Wc, int C,
                             float* out, int64_t ld_out, vms_stream stream);
/* Reverse mode.  g_out is the gradient w.r.t. the layer OUTPUT (post-activation); `out` is the saved output
 * (needed for relu / tanh; may be NULL for act none).  Any of g_x, g_W, g_b, g_cond, g_Wc may be NULL.
 * g_W / g_b / g_Wc are ACCUMULATED (+=) when accumulate != 0, else overwritten; g_x likewise via accumulate_x.
 * workspace: device scratch of at least vms_dense_backward_workspace(B, K, N, C) bytes.                      */
size_t vms_dense_backward_workspace(int64_t B, int K, int N, int C);
vms_status vms_dense_backward(const float* x, int64_t ld_x, const float* W, int64_t B, int K, int N, int act,
                              const float* out, int64_t ld_out, const float* g_out, int64_t ld_g,
                              const float* cond, int64_t ld_c, const float* Wc, int C, float* g_x, int64_t ld_gx,
                              int accumulate_x, float* g_W, float* g_b, float* g_cond, int64_t ld_gc, float* g_Wc,
                              int accumulate, void* workspace, vms_stream stream);
/* mappings.py:144-149: out = concat([x[:, ~periodic], cos(x[:, periodic]), sin(x[:, periodic])]).
 * periodic: device uint8 [D].  out is [B, D + n_periodic]. */
vms_status vms_periodic_featurise(const float* x, int64_t B, int D, const uint8_t* periodic, float* out,
                                  vms_stream stream);

/* --------------------------------------------------------------------------- K4: distributions (log_prob)
 * kinds per degree of freedom (dists.py:164-173): */
enum { VMS_DIST_NORMAL = 0, VMS_DIST_VONMISES = 1 };
/* scale / concentration transform applied to the raw parameter */
enum {
  VMS_SCALE_IDENTITY = 0,     /* parameter already constrained */
  VMS_SCALE_SOFTPLUS = 1,     /* tfp.layers.IndependentNormal (tests/test_models.py:167-170); dists.py:607 */
  VMS_SCALE_SOFTPLUS_EPS = 2  /* parameter_properties bijector Softplus(low=eps32): dists.py:56-78 */
};
/* Generic blockwise log_prob -- replaces `IndependentBlockwise.call` dists.py:210-217 + `tfp.distributions.Blockwise.
 * log_prob`, `IndependentVonMises.new` dists.py:602-610, tfp.layers.IndependentNormal, and the per-dof distributions
 * built inside `AutoregressiveBlockwise.call` dists.py:326-336.
 *   x [B, D] (ld_x); params [B, ld_p]; for dof i: kind[i], column offsets loc_off[i] (Normal loc, or von Mises
 *   sine), loc2_off[i] (von Mises cosine; ignored for Normal), scale_off[i]; loc = p[loc_off] or atan2(p[loc_off],
 *   p[loc2_off]).  lp[B] = sum_i log_prob_i  (added to the existing value if accumulate).
 * The four int arrays are HOST pointers (D <= 64), copied into kernel arguments.                               */
vms_status vms_blockwise_log_prob(const float* x, int64_t ld_x, const float* params, int64_t ld_p, int64_t B, int D,
                                  const int32_t* kind, const int32_t* loc_off, const int32_t* loc2_off,
                                  const int32_t* scale_off, int scale_mode, float* lp, int accumulate,
                                  vms_stream stream);
/* The constrained parameters themselves (`make_param_transform`, dists.py:28-87): loc [B, D], scale [B, D].
 * A von Mises dof with loc2_off[i] < 0 (or loc2_off NULL) takes loc = p[loc_off[i]] directly (no atan2).         */
vms_status vms_blockwise_params(const float* params, int64_t ld_p, int64_t B, int D, const int32_t* kind,
                                const int32_t* loc_off, const int32_t* loc2_off, const int32_t* scale_off,
                                int scale_mode, float* loc, float* scale, vms_stream stream);
/* One sample per (row, dof) -- replaces `tfp.distributions.Blockwise.sample` over the per-dof distributions of
 * dists.py:210-217 / :326-336 / :602-610 (used by mcmc.py:100-108 and models.py:139, :564).  Normal dofs: eps * scale +
 * loc with eps [B, D] (ld_eps) given by the caller (parity mode) or, when eps is NULL, Philox4x32-10 + Box-Muller keyed by
 * (seed, row, dof).  von Mises dofs: tfp's Best-Fisher rejection sampler on the same Philox stream, result wrapped to
 * [-pi, pi).  TF's RNG streams are not reproducible outside TF: sampling parity is distributional, not bitwise.      */
vms_status vms_blockwise_sample(const float* params, int64_t ld_p, int64_t B, int D, const int32_t* kind,
                                const int32_t* loc_off, const int32_t* loc2_off, const int32_t* scale_off, int scale_mode,
                                const float* eps, int64_t ld_eps, unsigned long long seed, float* out, int64_t ld_out,
                                vms_stream stream);
/* `IndependentDeterministic` log_prob (dists.py:688-704, tfp Deterministic): lp[B] = 0 where x == loc in every coordinate,
 * -inf elsewhere. */
vms_status vms_deterministic_log_prob(const float* x, int64_t ld_x, const float* loc, int64_t ld_loc, int64_t B, int D,
                                      float* lp, vms_stream stream);
/* out[0..n) ~ N(0, 1): Philox4x32-10 + Box-Muller, element i drawn from counter (offset + i) / 4 under key `seed`, so a
 * stream continues across calls by advancing `offset` (the reparameterisation noise of models.py:310 when the caller does
 * not supply eps; TF's own streams are not reproducible outside TF). */
vms_status vms_standard_normal(unsigned long long seed, unsigned long long offset, int64_t n, float* out, vms_stream stream);
/* Standard-normal base density used by every flowed prior in the reference's tests / notebooks
 * (tests/test_models.py:172-175): lp[B] (+)= sum_d -0.5 x^2 - 0.5 log 2pi.                                     */
vms_status vms_std_normal_log_prob(const float* x, int64_t ld_x, int64_t B, int D, float* lp, int accumulate,
                                   vms_stream stream);
/* Reparameterised Normal sample + log_prob in one pass (models.py:310 `encode_dist.sample()`, mcmc.py:100
 * `experimental_sample_and_log_prob`): params [B, ld_p] with loc at column loc_off + d, raw scale at scale_off + d;
 * z = eps * scale + loc (TFP Normal._sample_n);  lp = log N(z; loc, scale) summed over D.  z, lp nullable.       */
vms_status vms_normal_sample_log_prob(const float* params, int64_t ld_p, int loc_off, int scale_off, int scale_mode,
                                      const float* eps, int64_t B, int D, float* z, int64_t ld_z, float* lp,
                                      vms_stream stream);
/* Reverse mode of sum_d log N(x_d; loc_d, scale(raw_d)) for Normal dofs with per-row upstream g_lp [B]:
 *   g_x (nullable, += if accumulate_x), g_params [B, ld_gp] written at loc_off/scale_off columns (overwritten).
 * Replaces TF autodiff through `Normal._log_prob`.                                                              */
vms_status vms_normal_log_prob_backward(const float* x, int64_t ld_x, const float* params, int64_t ld_p, int loc_off,
                                        int scale_off, int scale_mode, const float* g_lp, int64_t B, int D,
                                        float* g_x, int64_t ld_gx, int accumulate_x, float* g_params, int64_t ld_gp,
                                        vms_stream stream);

/* ------------------------------------------------------------------------------- K5: ELBO / KL reductions
 * losses.py:253 `KLDivergenceEstimate.call` (weight applied as in `InfoRegularizer.__call__` :174-194):
 *   out[0] = weight * mean_B(lq - lp).   lp_b NULL => out = weight * mean(lq)  (LogProbRegularizer :296 uses
 *   sign = -1).  Deterministic (fixed-order two-level reduction).                                               */
vms_status vms_kl_mean(const float* lq, const float* lp, int64_t B, float weight, float* out, vms_stream stream);
/* losses.py:58 `LogProbLoss.call` + Keras mean reduction: out[0] = scale * mean_B(v) (scale = -1 for the NLL).  */
vms_status vms_scaled_mean(const float* v, int64_t B, float scale, float* out, vms_stream stream);
/* out = a*x + b*y elementwise (y nullable); the `+` of mcmc.py:103,109 on device tensors.                      */
vms_status vms_axpby(const float* x, const float* y, float a, float b, int64_t n, float* out, vms_stream stream);

/* Column-wise affine map of tfp.bijectors.Shift / Scale (flows.py:53-58, `make_domain_transform`):
 *   shift_first == 0: out[b,d] = x[b,d] * scale[d] + shift[d];   shift_first != 0: out = (x + shift) * scale.
 * scale / shift are device float32 [D]; either may be NULL (1 / 0).                                             */
vms_status vms_affine_cols(const float* x, int64_t ld_x, int64_t B, int D, const float* scale, const float* shift,
                           int shift_first, float* out, int64_t ld_out, vms_stream stream);

/* ------------------------------------------------------------------------------- batch normalisation (default off)
 * Replaces tf.keras.layers.BatchNormalization inside FCDeepNN (mappings.py:113-114) and tfp.bijectors.BatchNormalization
 * between flow blocks (flows.py:308-309, :623-624); both variants are off by default in the reference.
 *   vms_batch_moments   mean[D], var[D] = tf.nn.moments(x, axis=0): two passes, biased variance, fixed-order sums;
 *                       workspace: vms_batch_moments_workspace(B, D) bytes of device memory
 *   vms_batchnorm_coeffs  per-column scale / shift for vms_affine_cols:
 *                       denormalize == 0: x * inv + (beta - mean * inv), inv = rsqrt(var + eps) * gamma;
 *                       denormalize != 0: x * r + (mean - beta * r), r = sqrt(var + eps) / gamma (the bijector's forward);
 *                       ldj[1] (nullable) = +-(sum_d log gamma_d - 0.5 log(var_d + eps)): the map's log-det per row.
 *                       gamma / beta may be NULL (1 / 0).
 *   vms_broadcast_scalar  out[0..n) = *scalar (the per-row log-det vector of a Chain)                                 */
size_t vms_batch_moments_workspace(int64_t B, int D);
vms_status vms_batch_moments(const float* x, int64_t ld_x, int64_t B, int D, float* mean, float* var, void* workspace,
                             vms_stream stream);
vms_status vms_batchnorm_coeffs(const float* mean, const float* var, const float* gamma, const float* beta, int D, float eps,
                                int denormalize, float* scale, float* shift, float* ldj, vms_stream stream);
vms_status vms_broadcast_scalar(const float* scalar, int64_t n, float* out, vms_stream stream);
/* Reverse mode of the NORMALISING direction y = (x - mean) rsqrt(var + eps) gamma + beta (TF autodiff through
 * tf.nn.batch_normalization + tf.nn.moments; mappings.py:113-114 in training, flows.py:308-309 / :623-624 inverse):
 *   batch_stats != 0: mean / var are the batch moments of x (gradients flow through them);  == 0: constants (moving statistics)
 *   g_ldj_total (nullable, device scalar): sum over the rows of the upstream gradient of the bijector's per-row log-det
 *   sum_d log gamma_d - 0.5 log(var_d + eps).   g_x += ..., g_gamma += ..., g_beta += ... (g_gamma / g_beta nullable).
 * workspace: vms_batchnorm_backward_workspace(D) bytes of device memory.  Fixed-order column sums (deterministic). */
/* Cross-replica batch statistics for data-parallel training: each rank computes the moments of its shard
 * (vms_batch_moments), packs n mean_r (stage 1) or n (var_r + (mean_r - mean)^2) (stage 2, mean_g = the global mean) plus n
 * into buf [D + 1], the caller sums buf over the ranks (one small allreduce per stage) and vms_bn_sync_unpack divides by the
 * summed count: the global mean / biased variance, i.e. tf.nn.moments of the whole batch.  Reverse mode: _backward_sums
 * writes this rank's [s1 | s2 | G | B] as floats (2 D + 2), the caller sums a copy over the ranks, and _backward_apply uses
 * the GLOBAL sums for g_x and the LOCAL ones for g_gamma / g_beta (the trainer sums parameter gradients over the ranks). */
vms_status vms_bn_sync_pack(const float* mean_r, const float* var_r, const float* mean_g, int64_t n_rows, int D, float* buf,
                            vms_stream stream);
vms_status vms_bn_sync_unpack(const float* buf, int D, float* out, vms_stream stream);
vms_status vms_batchnorm_backward_sums(const float* x, int64_t ld_x, int64_t B, int D, const float* mean, const float* var,
                                       float eps, const float* g_out, int64_t ld_g, const float* g_ldj_total, float* sums,
                                       void* workspace, vms_stream stream);
vms_status vms_batchnorm_backward_apply(const float* x, int64_t ld_x, int64_t B, int D, const float* mean, const float* var,
                                        const float* gamma, float eps, const float* g_out, int64_t ld_g, const float* local_sums,
                                        const float* global_sums, float* g_x, int64_t ld_gx, float* g_gamma, float* g_beta,
                                        vms_stream stream);
size_t vms_batchnorm_backward_workspace(int D);
vms_status vms_batchnorm_backward(const float* x, int64_t ld_x, int64_t B, int D, const float* mean, const float* var,
                                  const float* gamma, float eps, int batch_stats, const float* g_out, int64_t ld_g,
                                  const float* g_ldj_total, float* g_x, int64_t ld_gx, float* g_gamma, float* g_beta,
                                  void* workspace, vms_stream stream);

/* ------------------------------------------------------------------------------- K6: DistanceSelection
 * Replaces `DistanceSelection.call` mappings.py:362-455 (TF sub / div / round / mul / reduce_sum / top_k / gather):
 *   local = coords - ref;  if box: local -= box * rint(local / box);  d2 = (lx^2 + ly^2) + lz^2 (no FMA);
 *   the k = max_included smallest d2 (ties -> lower index), sorted ascending;  rows with d2 > cutoff^2 zeroed.
 * coords: dense [B, N, 3] (row_splits NULL) or ragged values [sum N_i, 3] with row_splits int64 [B+1] (device).
 * ref [B, 3]; box: NULL, [3] (box_per_row = 0) or [B, 3] (box_per_row = 1); info [B, N, P] / ragged [sum N_i, P]
 * (nullable).  Outputs out_xyz [B, k, 3], out_info [B, k, P] (nullable), out_idx int32 [B, k] (nullable; exact
 * top-k indices including beyond-cutoff and padding slots, as tf.math.top_k would return them).
 * Rows shorter than k behave as if padded with float32.max coordinates (mappings.py:417-426).                  */
vms_status vms_dist_select(const float* coords, const int64_t* row_splits, int64_t B, int64_t N, const float* ref,
                           const float* box, int box_per_row, float cutoff_sq, int k, const float* info, int P,
                           float* out_xyz, float* out_info, int32_t* out_idx, vms_stream stream);
/* The same selection for B sites around ONE frame: coords [N, 3] and info [N, P] are shared by every row, i.e. the result
 * of vms_dist_select on the frame tiled B times (what mappings.py:362-455 makes a caller materialise for a simulation box:
 * `batch_size = tf.shape(coords)[0]`, :399) with the frame uploaded and read from HBM once.  No reference counterpart.   */
vms_status vms_dist_select_frame(const float* coords, int64_t N, const float* ref, int64_t B, const float* box,
                                 int box_per_row, float cutoff_sq, int k, const float* info, int P, float* out_xyz,
                                 float* out_info, int32_t* out_idx, vms_stream stream);

/* ------------------------------------------------------------------------------- K7: MC acceptance
 * Replaces mcmc.py:116-128:  log_acc = E_new + rev - E_old - fwd (float64; log-probs float32 promoted);
 * acc = log_acc >= log_u;  rejected rows restore x_old / E_old.  x_new_inout [B, D] is overwritten in place,
 * E_out [B] receives the selected energies, acc [B] uint8 (nullable), n_acc += number accepted (device u64).   */
vms_status vms_mc_accept(const double* E_new, const double* E_old, const float* fwd, const float* rev,
                         const double* log_u, int64_t B, int D, const float* x_old, float* x_new_inout,
                         double* E_out, uint8_t* acc, unsigned long long* n_acc, vms_stream stream);
/* The same rule for an energy callback that returns float32 (a tfp log_prob, MC_Moves_with_VAEs.ipynb cell 38): NumPy then
 * evaluates mcmc.py:116 in float32, ((E_new + rev) - E_old) - fwd, and only the comparison with log_u is in float64.     */
vms_status vms_mc_accept_f32(const float* E_new, const float* E_old, const float* fwd, const float* rev,
                             const double* log_u, int64_t B, int D, const float* x_old, float* x_new_inout, float* E_out,
                             uint8_t* acc, unsigned long long* n_acc, vms_stream stream);
/* Device-resident energies so the MC loop is not PCIe-bound (SURVEY 7 "hard parts"):
 *   kind 0: tests/test_mcmc.py:28-32  E = sum_d (x_d - means_d)^2, evaluated in float64 like the NumPy callback. */
vms_status vms_energy_quadratic(const float* x, int64_t B, int D, const double* means, double* E, vms_stream stream);
/*   Gaussian-mixture log-density, the "energy" of examples/MC_Moves_with_VAEs.ipynb cell 5 / 38 (tfp Mixture of Independent
 *   Normals, `data_dist.log_prob(configs)`): log_w [n_comp] = log cat probabilities, loc / scale [n_comp, D]; float32
 *   arithmetic as tfp (per-component sums, max-shifted logsumexp), float32 out like `log_prob(...).numpy()`; pair it
 *   with vms_mc_accept_f32.  n_comp <= 16.                                                                            */
vms_status vms_energy_gmm(const float* x, int64_t B, int D, int n_comp, const float* log_w, const float* loc,
                          const float* scale, float* E, vms_stream stream);

/* ------------------------------------------------------------------------------- K8: optimiser
 * Keras Adam (tests/test_models.py:181: lr 1e-3, beta 0.9 / 0.999, eps 1e-7), on a flat parameter buffer:
 *   lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  m += (g - m)(1 - b1);  v += (g^2 - v)(1 - b2);
 *   theta -= lr_t m / (sqrt(v) + eps).   `grad_scale` multiplies g first (1/world_size after an allreduce-sum).
 * g may be a stack of `n_partials` partial gradients [n_partials, n] that are summed in fixed order first.      */
vms_status vms_adam_step(float* theta, const float* g, int n_partials, float grad_scale, float* m, float* v,
                         int64_t n, int64_t t, double lr, double beta1, double beta2, double eps, vms_stream stream);
vms_status vms_sum_partials(const float* g, int n_partials, int64_t n, float scale, float* out, vms_stream stream);
/* The same update for every weight tensor of a model in one launch per VMS_ADAM_MULTI_MAX tensors (Keras applies Adam per
 * variable, models.py:85-139 `fit`; the op-by-op training path is bound by its launch count).  mask (nullable): 0 / 1
 * multiplier of the gradient, the constraint of tfp's AutoregressiveNetwork kernels (flows.py:450-487).                */
#define VMS_ADAM_MULTI_MAX 64
typedef struct {
  float* theta;
  const float* grad;
  const float* mask;
  float* m;
  float* v;
  int64_t n;
} vms_adam_tensor;
vms_status vms_adam_step_multi(const vms_adam_tensor* tensors, int n_tensors, float grad_scale, int64_t t, double lr,
                               double beta1, double beta2, double eps, vms_stream stream);
/* The same with the step count t and the bias-corrected learning rate in DEVICE memory (t_dev: int64, advanced by the call;
 * lr_t_dev: float scratch), so that a training step captured in a CUDA graph (vms_graph_*) stays valid from step to step. */
vms_status vms_adam_step_multi_dev(const vms_adam_tensor* tensors, int n_tensors, float grad_scale, long long* t_dev,
                                   float* lr_t_dev, double lr, double beta1, double beta2, double eps, vms_stream stream);

/* ------------------------------------------------------------------------------- fused ELBO step (C1 / C2)
 * One handle = one VAE of the family used by the reference's tests (tests/test_models.py:161-228):
 *   encoder  FCDeepNN dx -> hidden(relu) -> 2 dz  + tfp.layers.IndependentNormal(dz)
 *   prior    N(0, I)  (num_blocks = 0)  or  FlowedDistribution(RQSSplineRealNVP(num_blocks, K, flow_hidden), N(0,I))
 *   decoder  FCDeepNN dz -> hidden(relu) -> 2 dx  + tfp.layers.IndependentNormal(dx)
 *   loss     LogProbLoss (mean) + weight * KLDivergenceEstimate          (models.py:289-322, losses.py:58,:253)
 * Parameters live in ONE flat float32 device buffer owned by the caller, in the order
 *   enc.0.W enc.0.b enc.1.W enc.1.b dec.0.W dec.0.b dec.1.W dec.1.b  then per flow block: d1.W d1.b heads.W heads.b
 * (all W row-major [in, out], the Keras layout).  heads.W [flow_hidden, Dt*(3K-1)] is the column-wise concatenation
 * [bin_widths.W | bin_heights.W | knot_slopes.W] of the three Dense heads of flows.py:140-152 (heads.b likewise), so
 * one GEMM produces a block's raw spline parameters; the Python host converts from / to per-layer Keras arrays.   */
typedef struct {
  int32_t dx, dz, hidden;
  int32_t num_blocks, num_bins, flow_hidden; /* num_blocks = 0 => N(0, I) prior */
  float bin_min, bin_max;
  float kl_weight;
  int64_t max_batch;
} vms_elbo_desc;
typedef struct vms_elbo_plan_s* vms_elbo_plan;
vms_status vms_elbo_plan_create(const vms_elbo_desc* desc, vms_elbo_plan* plan);
vms_status vms_elbo_plan_destroy(vms_elbo_plan plan);
int64_t vms_elbo_param_count(const vms_elbo_desc* desc);
/* Two implementations sit behind the plan: a FUSED one (a single persistent kernel, 32-row tiles resident in shared
 * memory; chosen automatically when the shape fits: dx, dz <= 8, num_bins <= 32 and a multiple of 4, shared memory
 * <= 227 KB) and an UNFUSED one (per-layer kernels replayed as a CUDA graph; any shape).  mode 0 = auto, 1 = force
 * the unfused float32-FFMA path (used by the tests to cross-check the paths on the device), 2 = the unfused plan with
 * every RealNVP coupling block (flows.py:184-207 conditioner + spline, forward and reverse mode) as ONE tcgen05
 * kernel per block (flow_tc.cu: 3 x BF16 split, accumulators in TMEM, hidden layer and raw spline parameters never leave
 * the SM) -- the large-batch configuration; VMS_ERR_UNSUPPORTED when a block's shape does not fit (one transformed
 * dimension, <= 4 conditioner columns, 8 <= flow hidden <= 111, num_bins <= 32 and a multiple of 4, encoder / decoder
 * widths dx, dz <= 7, 2 dx, 2 dz <= 16, hidden <= 240).
 * mode 3 = whole-step tensor-core kernel (elbo_tcf.cu: all coupling blocks on tcgen05 + encoder / decoder of a 32-row tile
 * in one kernel; forward+backward / train_step only, B <= 3 x 32 x #SMs: up to three waves of one-tile CTAs); AUTO mode takes
 * it for those calls whenever the shape fits (VMS_TCF_AUTO=0 disables; an explicit tc_auto_batch threshold wins).  Its weight images (pre-split heads matrices, transposed MLP weights) are
 * written by the Adam update of the previous vms_elbo_train_step; call vms_elbo_plan_invalidate after changing theta by any
 * other means between two train steps (forward_backward always re-packs).
 * mode 4 = force the single FFMA fused kernel (elbo_fused.cu) for every call it supports (float32 FFMA cross-check).
 * vms_elbo_plan_tc_status: synchronises the device and reports whether any tensor-core completion wait ran into its
 * bound since the last call (err = 1: results of that interval are invalid); clears the flag.                      */
vms_status vms_elbo_plan_set_mode(vms_elbo_plan plan, int mode);
int vms_elbo_plan_is_fused(vms_elbo_plan plan);
vms_status vms_elbo_plan_tc_status(vms_elbo_plan plan, int* err);
/* Which implementation a forward + backward call with batch B takes: 0 = single FFMA fused kernel, 1 = per-layer FFMA
 * plan, 2 = tensor-core plan (one kernel per coupling block), 3 = whole-step tensor-core kernel. */
int vms_elbo_plan_path(vms_elbo_plan plan, int64_t B);
/* The plan's packed weight images no longer describe theta (host-side assignment between two train steps). */
vms_status vms_elbo_plan_invalidate(vms_elbo_plan plan);
/* Batch from which mode 0 prefers the mode-2 plan over the fused kernels (default: above one wave of 32-row tiles for the
 * FFMA kernel, above three waves for the whole-step tensor-core kernel; an explicit value applies to both). */
vms_status vms_elbo_plan_set_tc_auto_batch(vms_elbo_plan plan, int64_t batch);
/* Measurement aid (bench.py's roofline leg): with max_launches > 0 the fused path brackets its main kernel with CUDA
 * events on the launching stream for the next max_launches calls; vms_elbo_plan_kernel_ms synchronises, returns the
 * summed device time of that kernel and the number of launches measured, and resets the counter. */
vms_status vms_elbo_plan_set_timing(vms_elbo_plan plan, int max_launches);
/* The same, measuring every `every`-th call only (1 = every call): the two event records around the kernel cost the step
 * ~5 us (0.111 -> 0.1055 ms per C2 step without them), so a bench that also reports the step time samples the launches. */
vms_status vms_elbo_plan_set_timing_every(vms_elbo_plan plan, int max_launches, int every);
vms_status vms_elbo_plan_kernel_ms(vms_elbo_plan plan, double* total_ms, int* launches);
/* Forward only.  x [B, dx], eps [B, dz] (the reparameterisation noise is an INPUT in parity mode).  Outputs (any
 * nullable): z [B, dz], logq [B], logpz [B], logpx [B], scalars[3] = {loss, nll, kl} (kl unweighted mean).      */
vms_status vms_elbo_forward(vms_elbo_plan plan, const float* theta, const float* x, const float* eps, int64_t B,
                            float* z, float* logq, float* logpz, float* logpx, float* scalars, vms_stream stream);
/* Forward + backward: additionally writes the flat gradient of `loss` w.r.t. theta into grad [param_count].     */
vms_status vms_elbo_forward_backward(vms_elbo_plan plan, const float* theta, const float* x, const float* eps,
                                     int64_t B, float* grad, float* scalars, vms_stream stream);
/* One single-GPU training step: forward + backward + Adam (same update as vms_adam_step with grad_scale = 1).  On the
 * fused path the update rides in the kernel that sums the per-CTA partial gradients: 2 launches per step.  grad
 * [param_count] still receives the gradient.  (Data-parallel training keeps the three calls separate: the gradient
 * allreduce sits between vms_elbo_forward_backward and vms_adam_step.)                                            */
vms_status vms_elbo_train_step(vms_elbo_plan plan, float* theta, const float* x, const float* eps, int64_t B, float* grad,
                               float* scalars, float* m, float* v, int64_t t, double lr, double beta1, double beta2,
                               double eps_adam, vms_stream stream);
/* One DATA-PARALLEL training step (one process per GPU, peer buffers as for vms_peer_allreduce_adam below): forward + backward
 * on this rank's shard, gradient sum over the ranks through NVLink peer memory, Adam with 1 / world.  When the whole-step
 * tensor-core kernel serves the batch its finish kernel performs the exchange itself -- it sums the tile partials into this
 * rank's slot, the last block raises the rank's flag, and after the (bounded) wait every thread pulls its parameter from all
 * ranks, updates theta / m / v and writes the next step's weight images: two launches per step, as on one GPU.  Every other
 * plan runs vms_elbo_forward_backward into the slot followed by vms_peer_allreduce_adam.  `step` = 1, 2, 3, ... (the slot is
 * step & 1), identical on all ranks. */
vms_status vms_elbo_train_step_peer(vms_elbo_plan plan, float* theta, const float* x, const float* eps, int64_t B, float* scalars,
                                    float* m, float* v, int64_t t, double lr, double beta1, double beta2, double eps_adam,
                                    int world, int rank, void* const* peer_bases, unsigned long long step, vms_stream stream);

/* ------------------------------------------------------------------------------- fused MC run (C4a)
 * n_steps complete VAE-proposal MC steps (mcmc.py:68-130, the loop of mcmc.py:133-159) of B independent chains in ONE
 * launch, for the Gaussian-VAE family of tests/test_mcmc.py:14-26: FCDeepNN dx -> hidden(relu) -> 2 dz encoder and
 * dz -> hidden(relu) -> 2 dx decoder with tfp.layers.IndependentNormal heads, N(0, I) prior, and the quadratic energy
 * of tests/test_mcmc.py:28-32 evaluated on the device in float64.  theta: the flat parameter buffer of the ELBO plan
 * with num_blocks = 0 (enc.0.W enc.0.b enc.1.W enc.1.b dec.0.W dec.0.b dec.1.W dec.1.b).
 *   x [B, dx] float32 and E [B] float64 are the chain state, updated in place (E is computed from x first when
 *   energies_valid == 0);  log_u [n_steps, B] float64 = log of the host's PCG64 uniforms (mcmc.py:119), so decisions
 *   are bit-identical to the reference under the same seed;  means [dx] float64.
 *   noise: NULL => Philox4x32-10 + Box-Muller on the device keyed by (seed, global chain index, step0 + step): results
 *   do not depend on the grid or on how chains are sharded over GPUs;  else float32 [n_steps, B, 2 dz + dx] =
 *   eps(z1) | eps(z2) | eps(x2) per chain-step (parity mode: the reference draws in this order, mcmc.py:100-102).
 *   n_acc (device u64) += number of accepted proposals.  Optional traces [n_steps, B]: acc (uint8), forward_log_p,
 *   reverse_log_p (float32), proposal energies (float64).  All pointers are device pointers.
 *   The launcher picks the kernel by B: below 128 chains per SM a warp owns four chains and the hidden layers are summed as
 *   32 unit streams, above it 1 / 2 / 4 lanes own a chain and sum four streams -- noise and uniforms are the same, the
 *   log-probabilities agree to float32 rounding.  VMS_MC_TPC = 1 | 2 | 4 (four streams) or 8 (32 streams) pins one order
 *   when runs with different chain counts per GPU must reproduce each other bit for bit.                                   */
typedef struct {
  int32_t dx, dz, hidden;
} vms_mc_desc;
typedef struct vms_mc_plan_s* vms_mc_plan;
int64_t vms_mc_param_count(const vms_mc_desc* desc);
vms_status vms_mc_plan_create(const vms_mc_desc* desc, vms_mc_plan* plan);
vms_status vms_mc_plan_destroy(vms_mc_plan plan);
vms_status vms_mc_run(vms_mc_plan plan, const float* theta, float* x, double* E, int energies_valid, const float* noise,
                      unsigned long long seed, unsigned long long step0, const double* log_u, const double* means,
                      int64_t B, int n_steps, unsigned long long* n_acc, uint8_t* acc_trace, float* fwd_trace,
                      float* rev_trace, double* e_new_trace, vms_stream stream);

/* The accept uniforms drawn ON THE DEVICE from NumPy's PCG64 stream (mcmc.py:119 `self._rng.random(size=B)` followed by
 * np.log): chain c at step k owns draw k * n_global + chain0 + c of the stream that starts at `state`; the 128-bit LCG is
 * jumped ahead per chain (exact integer arithmetic), u = (xsl_rr(state) >> 11) * 2^-53 is bit-identical to NumPy's, log u
 * is CUDA's double log.  Decisions can differ from np.log's only where |log_acc - log u| <= 1e-13 max(1, |log u|): those
 * chain-steps are counted in n_uncertain (device u64, +=) and the caller re-runs the call on the host stream (vms_mc_run)
 * when it is non-zero.  The caller advances its generator by n_steps * n_global afterwards.
 *   state / inc: `np.random.Generator.bit_generator.state['state']` (128-bit each, split hi / lo);
 *   stride_mul / stride_add: the affine map of n_global LCG steps (state' = stride_mul * state + stride_add mod 2^128).
 * vms_mc_plan_has_device_rng: 1 when the plan's kernel carries the stream (the C4a shape dx = 6, dz = 2), else 0 and
 * vms_mc_run_pcg64 returns VMS_ERR_UNSUPPORTED.  Optional trace log_u_trace [n_steps, B] float64 = the log u used.       */
typedef struct {
  uint64_t state_hi, state_lo, inc_hi, inc_lo;
  uint64_t stride_mul_hi, stride_mul_lo, stride_add_hi, stride_add_lo;
  int64_t chain0;
} vms_pcg64_stream;
int vms_mc_plan_has_device_rng(vms_mc_plan plan);
/* Global index of chain 0 of the following vms_mc_run calls (default 0): the device noise stream is keyed by the GLOBAL chain
 * index, so a job sharded over GPUs draws the noise of the single-GPU run (vms_mc_run_pcg64 takes it from rng->chain0). */
vms_status vms_mc_plan_set_chain_offset(vms_mc_plan plan, int64_t chain0);
vms_status vms_mc_run_pcg64(vms_mc_plan plan, const float* theta, float* x, double* E, int energies_valid, const float* noise,
                            unsigned long long seed, unsigned long long step0, const vms_pcg64_stream* rng,
                            const double* means, int64_t B, int n_steps, unsigned long long* n_acc,
                            unsigned long long* n_uncertain, uint8_t* acc_trace, float* fwd_trace, float* rev_trace,
                            double* e_new_trace, double* log_u_trace, vms_stream stream);

/* ------------------------------------------------------------------------------- K9: fused MC, MC-notebook family (C4b)
 * Whole MC steps (mcmc.py:68-130, the loop of MCMC.run :133-159) for the model of examples/MC_Moves_with_VAEs.ipynb in ONE
 * launch, chain state in registers:
 *   encoder  FCDeepNN(2 -> enc_hidden -> 2, relu) into tfp.layers.IndependentNormal(1)                  (cell 11)
 *   prior    FlowedDistribution(RQSSplineMAF over a 1-D latent, Independent N(0, 1))                     (cell 14)
 *   decoder  FCDeepNN(1 -> dec_hidden -> (2, 2), relu) into AutoregressiveBlockwise(2, [Normal] * 2, conditional = z,
 *            MADE hidden_units made_hidden[0..2], activation made_act)  (cell 17; dists.py:246-340)
 *   energy   mixture-of-independent-Normals log-density (cell 5 / 38), float32 -- see vms_energy_gmm.
 * All pointers are device pointers to the host mirror's live weight tensors (row-major [in, out] kernels, as Keras stores
 * them; MADE kernels pre-masked; made_Wc = the bias-free conditional kernels [1, out] of every MADE layer).
 * `tables`: n_blocks knot tables in the chain's SAMPLING order (block 0 is applied first to the base noise), each
 * vms_rqs_knot_table_doubles(n_bins) doubles, built by vms_rqs_knot_table from the raw conditioner outputs.  A masked
 * autoregressive flow over ONE dimension has input-independent parameters (flows.py:450-515 with event size 1: the MADE
 * mask is empty), so the host evaluates the three conditioner networks of a block once per call, on one row.             */
typedef struct {
  int dx, dz;                     /* must be 2 and 1 */
  int enc_hidden, dec_hidden;
  const float *enc_W0, *enc_b0, *enc_W1, *enc_b1;
  const float *dec_W0, *dec_b0, *dec_W1, *dec_b1;
  int made_hidden[3];             /* [<= 16, <= 512, <= 16] */
  int made_act;                   /* VMS_ACT_*: tfp AutoregressiveNetwork `activation` (default none) */
  int made_first_dof;             /* 0 / 1: the MADE masks put this dof first (input_order left-to-right / right-to-left): its
                                     parameters depend on the conditional input only and the other dof is hidden from every
                                     unit, so tfp's D + 1 sampling passes + the log_prob pass give bit-identical values to ONE
                                     pass; -1: no such structure, run every pass */
  const float* made_W[4];
  const float* made_b[4];
  const float* made_Wc[4];
  int n_blocks, n_bins;
  float range_min, range_max;
  const double* tables;
  int n_comp;                     /* <= 16 */
  const float *gmm_log_w, *gmm_loc, *gmm_scale;   /* [n_comp], [n_comp, 2], [n_comp, 2] */
} vms_mc_nb_model;
/* Knot table of n_splines splines from raw parameters raw_w / raw_h [n_splines, n_bins], raw_s [n_splines, n_bins - 1]:
 * flows.py:86-101 "activations" (softmax * (range - n_bins 1e-2) + 1e-2; softplus + 1e-2) and the cumulative knot positions
 * of tfp RationalQuadraticSpline, in the arithmetic of the RQS kernels (float64 knots, float32 derivatives).
 * Layout per spline: double kx[n_bins + 1], double ky[n_bins + 1], float d[n_bins + 1] (+ pad to 8 bytes).               */
int64_t vms_rqs_knot_table_doubles(int n_bins);
vms_status vms_rqs_knot_table(const float* raw_w, const float* raw_h, const float* raw_s, int n_splines, int n_bins,
                              float range_min, float range_max, double* tables, vms_stream stream);
/* 1 when the shape is inside the kernel's family (the host mirror falls back to the op-by-op path otherwise). */
int vms_mc_nb_supported(const vms_mc_nb_model* model);
/* n_steps MC steps of B chains.  x [B, 2] and E [B] float32 are updated in place (E is computed first when
 * energies_valid = 0).  noise: NULL (device Philox stream keyed by (seed, chain0 + chain, step0 + step)) or
 * [n_steps, B, 4] = eps(z1) | eps(z2) | eps(x2) [2] per chain, the order mcmc.py:100-102 draws it.  Accept uniforms: exactly
 * one of log_u [n_steps, B] (float64 logs of the host stream, mcmc.py:119) and rng (the stream regenerated on the device,
 * see vms_pcg64_stream; uncertain decisions counted in n_uncertain).  Acceptance: NumPy's float32 evaluation of mcmc.py:116
 * (vms_mc_accept_f32).  n_acc += accepted moves.  Optional traces [n_steps, B]: acc u8, fwd / rev / e_new f32, log_u f64. */
vms_status vms_mc_nb_run(const vms_mc_nb_model* model, float* x, float* E, int energies_valid, const float* noise,
                         unsigned long long seed, unsigned long long step0, const double* log_u, const vms_pcg64_stream* rng,
                         int64_t chain0, int64_t B, int n_steps, unsigned long long* n_acc, unsigned long long* n_uncertain,
                         uint8_t* acc_trace, float* fwd_trace, float* rev_trace, float* e_new_trace, double* log_u_trace,
                         vms_stream stream);

/* ------------------------------------------------------------------------------- data-parallel exchange step
 * The single collective of data-parallel training (north_star: one gradient allreduce per step) fused with the Adam
 * update, as ONE kernel over NVLink peer memory.  Every rank owns a buffer of vms_peer_buffer_bytes(P) bytes
 * ([2][P] float gradient slots, double-buffered by step parity, + flags), zero-initialised, shared with the other
 * ranks of the node through CUDA IPC (handles are exchanged by the host: torch.distributed in this repo).
 *   step s (1, 2, ...): the rank's gradient is written into slot (s & 1) of ITS buffer (pass base + (s & 1) * P as the
 *   `grad` of vms_elbo_forward_backward), then every rank calls vms_peer_allreduce_adam(step = s): ranks signal and
 *   wait through flags in peer memory, each thread pulls one parameter's gradient from all world buffers over NVLink,
 *   sums in rank order (identical on every rank), scales by grad_scale (1 / world: the loss is a batch MEAN,
 *   losses.py:253) and applies Keras Adam (same update as vms_adam_step) to the local theta / m / v.
 *   peer_bases: HOST array of `world` device pointers (this rank's own buffer at index `rank`).  world <= 8.      */
size_t vms_peer_buffer_bytes(int64_t n_params);
vms_status vms_ipc_get_handle(void* device_ptr, unsigned char handle[64]);
vms_status vms_ipc_open_handle(const unsigned char handle[64], void** device_ptr);
vms_status vms_ipc_close_handle(void* device_ptr);
vms_status vms_peer_allreduce_adam(int world, int rank, void* const* peer_bases, int64_t n_params, unsigned long long step,
                                   float grad_scale, float* theta, float* m, float* v, int64_t t, double lr, double beta1,
                                   double beta2, double eps, float* grad_out, vms_stream stream);

/* ------------------------------------------------------------------------------- reverse mode of the op-by-op path
 * What TF autodiff does for compositions outside the fused ELBO family (tests/test_models.py:189-262: von Mises encoder,
 * periodic FCDeepNN, blockwise / autoregressive / MAF-flowed decoders; models.py:85-139 FlowModel).  The host keeps a tape of
 * the kernels it launched (vaemolsim_b200/_autodiff.py) and replays it backwards through these entry points, the existing
 * vms_dense_backward / vms_rqs_apply_backward / vms_normal_log_prob_backward, and vms_adam_step.  Every output ACCUMULATES.
 *   vms_blockwise_log_prob_backward   d/dx and d/dparams of vms_blockwise_log_prob (Normal and von Mises dofs, atan2 /
 *                                     softplus parameter transforms of dists.py:56-78 included); g_x / g_params nullable
 *   vms_std_normal_log_prob_backward  g_x += -x g_lp
 *   vms_blockwise_sample_backward     d z / d params of a reparameterised sample z (vms_blockwise_sample): Normal = pathwise;
 *                                     von Mises = loc pathwise + tfp's IMPLICIT reparameterisation in the concentration,
 *                                     -dF/dk / p(z) with F = von_mises_cdf (Hill's 20-term series below k = 10.5, corrected
 *                                     Normal approximation above), as tfp von_mises.py `_von_mises_sample_bwd`
 *   vms_periodic_featurise_backward   mappings.py:144-149: g_x += g_out[~periodic part], -sin x g_cos + cos x g_sin
 *   vms_add_cols / vms_add_scalar / vms_mul_inplace / vms_sum_all   strided +=, dst += alpha * scalar[0] (scalar NULL: alpha),
 *                                     dst *= src (MADE masks on kernel gradients), out[0] += alpha * sum(src) (fixed order)   */
vms_status vms_blockwise_log_prob_backward(const float* x, int64_t ld_x, const float* params, int64_t ld_p, int64_t B, int D,
                                           const int32_t* kind, const int32_t* loc_off, const int32_t* loc2_off,
                                           const int32_t* scale_off, int scale_mode, const float* g_lp, float* g_x,
                                           int64_t ld_gx, float* g_params, int64_t ld_gp, vms_stream stream);
vms_status vms_std_normal_log_prob_backward(const float* x, int64_t ld_x, int64_t B, int D, const float* g_lp, float* g_x,
                                            int64_t ld_gx, vms_stream stream);
vms_status vms_blockwise_sample_backward(const float* params, int64_t ld_p, int64_t B, int D, const int32_t* kind,
                                         const int32_t* loc_off, const int32_t* loc2_off, const int32_t* scale_off,
                                         int scale_mode, const float* z, int64_t ld_z, const float* g_z, int64_t ld_gz,
                                         float* g_params, int64_t ld_gp, vms_stream stream);
vms_status vms_periodic_featurise_backward(const float* x, int64_t B, int D, const uint8_t* periodic, int n_periodic,
                                           const float* g_out, float* g_x, vms_stream stream);
vms_status vms_add_cols(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t B, int D, float alpha,
                        vms_stream stream);
vms_status vms_add_scalar(float* dst, int64_t n, const float* scalar, float alpha, vms_stream stream);
vms_status vms_mul_inplace(float* dst, const float* src, int64_t n, vms_stream stream);
vms_status vms_sum_all(const float* src, int64_t n, float alpha, float* out, vms_stream stream);

/* ------------------------------------------------------------ geometric-algebra attention over a selected point cloud
 * mappings.py:480-561 (AttentionBlock), :564-688 (ParticleEmbedding): the arithmetic of
 * geometric_algebra_attention.keras.VectorAttention(score_net, value_net, reduce, merge_fun='concat', join_fun='concat',
 * rank=2) (mappings.py:518-525, :633-647; third-party, unpinned, restated in oracle/gaa.py) and of
 * tf.keras.layers.LayerNormalization / Masking.  Pair tensors are [B, n(i), n(j), .]: entry (i, j) belongs to the geometric
 * product r_j * r_i; reduce = 0 sums over j for every i ([B, n, D]), reduce = 1 over all pairs of a cloud ([B, D]).
 *   vms_gaa_zero_mask         Masking(mask_value=0).compute_mask: mask[p] = any(coords[p, :] != 0)
 *   vms_gaa_pair_invariants   out[b, i, j, :] = (r_j . r_i, |r_j ^ r_i|)     (vecvec + vecvec_invariants, TF op order)
 *   vms_layernorm_forward     y = act(LayerNormalization(x)) over the last axis of [R, H] (biased variance, epsilon inside the
 *                             square root); stats [R, 2] = (mean, 1 / sqrt(var + eps)) for the reverse mode (nullable)
 *   vms_layernorm_backward    g_x += ..., g_gamma += ..., g_beta += ... (column sums over per-CTA partials, fixed order; H <= 128)
 *   vms_gaa_pair_merge        out[b, i, j, :] = u[b, j, :] + w[b, i, :]  (u = v merge_kernel_0, w = v merge_kernel_1)
 *   vms_gaa_pair_merge_backward   g_u[b, p, :] += sum_i g[b, i, p, :],  g_w[b, p, :] += sum_j g[b, p, j, :]
 *   vms_gaa_attend            scores [B, n, n] (pairs with a masked member: -1e9) -> softmax (per row / per cloud) ->
 *                             out = sum attention x values [B, n, n, D]; attention [B, n, n] is kept for the reverse mode
 *   vms_gaa_attend_backward   g_scores += att (g_out . values - sum att g_out . values) (0 at masked pairs), g_values += att g_out
 *   vms_gaa_attention_forward ONE kernel for a whole VectorAttention layer (inference / sampling path): a thread owns a pair
 *                             from its invariants to its score, weights are shared-memory broadcasts, online softmax; no pair
 *                             tensor reaches HBM.  D <= 32, H <= 64 (vms_gaa_attention_forward_supported), else
 *                             VMS_ERR_UNSUPPORTED and the caller composes the kernels above with vms_dense_forward.      */
typedef struct vms_gaa_weights {
  const float* merge0; const float* merge1;           /* merge_kernel_0 / _1 [D, D] */
  const float* join1; const float* join2;             /* join_kernel_1 (invariant values) / _2 (merged values) [D, D] */
  const float* score_w1; const float* score_b1;       /* score_net: Dense(H, act) [D, H], [H] */
  const float* score_w2; const float* score_b2;       /*            Dense(1)      [H, 1], [1] */
  const float* value_w1; const float* value_b1;       /* value_net: Dense(H) [2, H], [H] */
  const float* value_gamma; const float* value_beta;  /*            LayerNormalization [H], [H]; then Activation */
  const float* value_w2; const float* value_b2;       /*            Dense(D) [H, D], [D] */
} vms_gaa_weights;
vms_status vms_gaa_zero_mask(const float* coords, int64_t n_particles, uint8_t* mask, vms_stream stream);
vms_status vms_gaa_pair_invariants(const float* coords, int64_t B, int n, float* out, vms_stream stream);
vms_status vms_layernorm_forward(const float* x, int64_t ld_x, int64_t R, int H, const float* gamma, const float* beta,
                                 float eps, int act, float* y, int64_t ld_y, float* stats, vms_stream stream);
size_t vms_layernorm_backward_workspace(int64_t R, int H);
vms_status vms_layernorm_backward(const float* x, int64_t ld_x, int64_t R, int H, const float* gamma, const float* stats,
                                  int act, const float* y, int64_t ld_y, const float* g_y, int64_t ld_gy, float* g_x,
                                  int64_t ld_gx, float* g_gamma, float* g_beta, void* workspace, vms_stream stream);
vms_status vms_gaa_pair_merge(const float* u, int64_t ld_u, const float* w, int64_t ld_w, int64_t B, int n, int D, float* out,
                              vms_stream stream);
vms_status vms_gaa_pair_merge_backward(const float* g, int64_t B, int n, int D, float* g_u, int64_t ld_gu, float* g_w,
                                       int64_t ld_gw, vms_stream stream);
vms_status vms_gaa_attend(const float* scores, const float* values, const uint8_t* mask, int64_t B, int n, int D, int reduce,
                          float* out, float* attention, vms_stream stream);
vms_status vms_gaa_attend_backward(const float* attention, const float* values, const uint8_t* mask, int64_t B, int n, int D,
                                   int reduce, const float* g_out, float* g_scores, float* g_values, vms_stream stream);
int vms_gaa_attention_forward_supported(int n, int D, int H);
vms_status vms_gaa_attention_forward(const float* coords, const float* values, int64_t ld_v, const uint8_t* mask, int64_t B,
                                     int n, int D, int H, const vms_gaa_weights* w, int reduce, int act, float ln_eps,
                                     float* out, vms_stream stream);

/* ------------------------------------------------------------------------------- machine-peak probes (measurement aid)
 * The denominators of this repo's compute-bound roofline fractions, measured on the device they are quoted for
 * (BASELINE.md section 2 asks for them; `bench.py` runs them live and `scripts/measure_peaks.py` writes profiles/*.json):
 *   vms_probe_ffma  FP32 FFMA issue peak: 2048 resident threads per SM, 8 independent FMA chains per thread out of
 *                   registers, `iters` x 128 FMAs per thread; tflops = 2 x FMAs / best-of-`reps` event time.
 *   vms_probe_mma   tcgen05.mma cta_group::1 issue peak, kind 0 = kind::f16 with bfloat16 inputs (K = 16 per MMA), kind 1 =
 *                   kind::tf32 (K = 8): one CTA per SM issues n_mma back-to-back M x N x K MMAs on shared-memory-resident
 *                   operands into two alternating TMEM accumulators; nothing is loaded or stored in the timed region.     */
vms_status vms_probe_ffma(int iters, int reps, double* tflops, double* ms, vms_stream stream);
/* The same with the packed instruction fma.rn.f32x2 (SASS FFMA2: two float32 FMAs per instruction and register pair). */
vms_status vms_probe_ffma2(int iters, int reps, double* tflops, double* ms, vms_stream stream);
vms_status vms_probe_mma(int kind, int M, int N, int n_mma, int reps, double* tflops, double* ms, vms_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* VMS_B200_H */
