"""ctypes binding of libvms_b200.so (include/vms_b200.h) and a minimal device `Tensor`.

This is the only module that touches the shared library.  There is deliberately NO fallback: if the library is
missing or no CUDA device is present, the first compute call raises (the reference's TF ops are replaced by
sm_100a kernels, not by NumPy).  NumPy is used for host-side staging only.
"""
import ctypes as C
import math
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvms_b200.so')

c_f32p = C.c_void_p  # device pointers travel as integers
c_i64 = C.c_int64
c_int = C.c_int
c_f32 = C.c_float
c_f64 = C.c_double
c_size = C.c_size_t
c_vp = C.c_void_p


class RqsArgs(C.Structure):
    _fields_ = [('n_rows', c_i64), ('n_dims', C.c_int32), ('num_bins', C.c_int32), ('bin_min', c_f32),
                ('bin_max', c_f32), ('v_in', c_vp), ('ld_in', c_i64), ('raw_w', c_vp), ('ld_w', c_i64),
                ('raw_h', c_vp), ('ld_h', c_i64), ('raw_s', c_vp), ('ld_s', c_i64), ('v_out', c_vp),
                ('ld_out', c_i64), ('ldj', c_vp), ('ldj_sum', c_vp), ('accumulate', C.c_int32),
                ('inverse_dir', C.c_int32)]


class RqsBwdArgs(C.Structure):
    _fields_ = [('fwd', RqsArgs), ('g_out', c_vp), ('ld_g_out', c_i64), ('g_ldj_sum', c_vp), ('g_in', c_vp),
                ('ld_g_in', c_i64), ('g_raw_w', c_vp), ('ld_gw', c_i64), ('g_raw_h', c_vp), ('ld_gh', c_i64),
                ('g_raw_s', c_vp), ('ld_gs', c_i64)]


class McDesc(C.Structure):
    _fields_ = [('dx', C.c_int32), ('dz', C.c_int32), ('hidden', C.c_int32)]


class Pcg64Stream(C.Structure):
    _fields_ = [('state_hi', C.c_uint64), ('state_lo', C.c_uint64), ('inc_hi', C.c_uint64), ('inc_lo', C.c_uint64),
                ('stride_mul_hi', C.c_uint64), ('stride_mul_lo', C.c_uint64), ('stride_add_hi', C.c_uint64),
                ('stride_add_lo', C.c_uint64), ('chain0', c_i64)]


class McNbModel(C.Structure):
    """vms_mc_nb_model (include/vms_b200.h): device pointers to the live weights of the MC notebook's model family."""
    _fields_ = [('dx', C.c_int), ('dz', C.c_int), ('enc_hidden', C.c_int), ('dec_hidden', C.c_int),
                ('enc_W0', c_vp), ('enc_b0', c_vp), ('enc_W1', c_vp), ('enc_b1', c_vp),
                ('dec_W0', c_vp), ('dec_b0', c_vp), ('dec_W1', c_vp), ('dec_b1', c_vp),
                ('made_hidden', C.c_int * 3), ('made_act', C.c_int), ('made_first_dof', C.c_int),
                ('made_W', c_vp * 4), ('made_b', c_vp * 4), ('made_Wc', c_vp * 4),
                ('n_blocks', C.c_int), ('n_bins', C.c_int), ('range_min', c_f32), ('range_max', c_f32),
                ('tables', c_vp), ('n_comp', C.c_int), ('gmm_log_w', c_vp), ('gmm_loc', c_vp), ('gmm_scale', c_vp)]


class AdamTensor(C.Structure):
    """vms_adam_tensor (include/vms_b200.h)."""
    _fields_ = [('theta', c_vp), ('grad', c_vp), ('mask', c_vp), ('m', c_vp), ('v', c_vp), ('n', c_i64)]


class GaaWeights(C.Structure):
    """vms_gaa_weights (include/vms_b200.h): device pointers of one VectorAttention layer."""
    _fields_ = [(k, c_vp) for k in ('merge0', 'merge1', 'join1', 'join2', 'score_w1', 'score_b1', 'score_w2', 'score_b2',
                                    'value_w1', 'value_b1', 'value_gamma', 'value_beta', 'value_w2', 'value_b2')]


class ElboDesc(C.Structure):
    _fields_ = [('dx', C.c_int32), ('dz', C.c_int32), ('hidden', C.c_int32), ('num_blocks', C.c_int32),
                ('num_bins', C.c_int32), ('flow_hidden', C.c_int32), ('bin_min', c_f32), ('bin_max', c_f32),
                ('kl_weight', c_f32), ('max_batch', c_i64)]


# name -> (restype, argtypes); restype None means vms_status (checked)
_SIGS = {
    'vms_last_error': (C.c_char_p, []),
    'vms_abi_version': (c_int, []),
    'vms_launch_count': (C.c_ulonglong, []),
    'vms_device_count': (None, [C.POINTER(c_int)]),
    'vms_set_device': (None, [c_int]),
    'vms_device_info': (None, [c_int, C.POINTER(c_i64)]),
    'vms_malloc': (None, [C.POINTER(c_vp), c_size]),
    'vms_free': (None, [c_vp]),
    'vms_malloc_host': (None, [C.POINTER(c_vp), c_size]),
    'vms_free_host': (None, [c_vp]),
    'vms_memcpy_h2d': (None, [c_vp, c_vp, c_size, c_vp]),
    'vms_memcpy_d2h': (None, [c_vp, c_vp, c_size, c_vp]),
    'vms_memcpy_d2d': (None, [c_vp, c_vp, c_size, c_vp]),
    'vms_memcpy2d_d2d': (None, [c_vp, c_size, c_vp, c_size, c_size, c_size, c_vp]),
    'vms_memset': (None, [c_vp, c_int, c_size, c_vp]),
    'vms_stream_create': (None, [C.POINTER(c_vp)]),
    'vms_stream_destroy': (None, [c_vp]),
    'vms_stream_synchronize': (None, [c_vp]),
    'vms_device_synchronize': (None, []),
    'vms_event_create': (None, [C.POINTER(c_vp)]),
    'vms_event_destroy': (None, [c_vp]),
    'vms_event_record': (None, [c_vp, c_vp]),
    'vms_stream_wait_event': (None, [c_vp, c_vp]),
    'vms_event_synchronize': (None, [c_vp]),
    'vms_event_elapsed_ms': (None, [c_vp, c_vp, C.POINTER(c_f32)]),
    'vms_graph_begin_capture': (None, [c_vp]),
    'vms_graph_end_capture': (None, [c_vp, C.POINTER(c_vp), C.POINTER(c_int)]),
    'vms_graph_abort_capture': (None, [c_vp]),
    'vms_graph_launch': (None, [c_vp, c_int, c_vp]),
    'vms_graph_destroy': (None, [c_vp]),
    'vms_rqs_forward': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_f32, c_vp, c_vp, c_vp]),
    'vms_rqs_inverse': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_f32, c_vp, c_vp, c_vp]),
    'vms_rqs_backward': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_f32, c_int, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_vp, c_vp]),
    'vms_rqs_apply': (None, [C.POINTER(RqsArgs), c_vp]),
    'vms_rqs_apply_backward': (None, [C.POINTER(RqsBwdArgs), c_vp]),
    'vms_dense_forward': (None, [c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_int, c_vp,
                                 c_i64, c_vp]),
    'vms_dense_backward_workspace': (c_size, [c_i64, c_int, c_int, c_int]),
    'vms_dense_backward': (None, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_vp,
                                  c_i64, c_vp, c_int, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_i64, c_vp, c_int, c_vp,
                                  c_vp]),
    'vms_periodic_featurise': (None, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    'vms_blockwise_log_prob': (None, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_int, c_vp,
                                      c_int, c_vp]),
    'vms_deterministic_log_prob': (None, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    'vms_batch_moments_workspace': (c_size, [c_i64, c_int]),
    'vms_batch_moments': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vms_batchnorm_coeffs': (None, [c_vp, c_vp, c_vp, c_vp, c_int, c_f32, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vms_broadcast_scalar': (None, [c_vp, c_i64, c_vp, c_vp]),
    'vms_bn_sync_pack': (None, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    'vms_bn_sync_unpack': (None, [c_vp, c_int, c_vp, c_vp]),
    'vms_batchnorm_backward_sums': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_f32, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    'vms_batchnorm_backward_apply': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_f32, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64,
                                            c_vp, c_vp, c_vp]),
    'vms_batchnorm_backward_workspace': (c_size, [c_int]),
    'vms_batchnorm_backward': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_f32, c_int, c_vp, c_i64, c_vp, c_vp, c_i64,
                                      c_vp, c_vp, c_vp, c_vp]),
    'vms_standard_normal': (None, [C.c_ulonglong, C.c_ulonglong, c_i64, c_vp, c_vp]),
    'vms_blockwise_sample': (None, [c_vp, c_i64, c_i64, c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_int, c_vp, c_i64, C.c_ulonglong, c_vp,
                                    c_i64, c_vp]),
    'vms_blockwise_params': (None, [c_vp, c_i64, c_i64, c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_int, c_vp, c_vp, c_vp]),
    'vms_std_normal_log_prob': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_int, c_vp]),
    'vms_normal_sample_log_prob': (None, [c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp, c_i64, c_vp,
                                          c_vp]),
    'vms_normal_log_prob_backward': (None, [c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_i64, c_int, c_vp,
                                            c_i64, c_int, c_vp, c_i64, c_vp]),
    'vms_kl_mean': (None, [c_vp, c_vp, c_i64, c_f32, c_vp, c_vp]),
    'vms_scaled_mean': (None, [c_vp, c_i64, c_f32, c_vp, c_vp]),
    'vms_axpby': (None, [c_vp, c_vp, c_f32, c_f32, c_i64, c_vp, c_vp]),
    'vms_affine_cols': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    'vms_dist_select': (None, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_int, c_f32, c_int, c_vp, c_int, c_vp, c_vp,
                               c_vp, c_vp]),
    'vms_dist_select_frame': (None, [c_vp, c_i64, c_vp, c_i64, c_vp, c_int, c_f32, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vms_mc_accept': (None, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_mc_accept_f32': (None, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_energy_quadratic': (None, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    'vms_energy_gmm': (None, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_adam_step_multi': (None, [C.POINTER(AdamTensor), c_int, c_f32, c_i64, c_f64, c_f64, c_f64, c_f64, c_vp]),
    'vms_adam_step_multi_dev': (None, [C.POINTER(AdamTensor), c_int, c_f32, c_vp, c_vp, c_f64, c_f64, c_f64, c_f64, c_vp]),
    'vms_adam_step': (None, [c_vp, c_vp, c_int, c_f32, c_vp, c_vp, c_i64, c_i64, c_f64, c_f64, c_f64, c_f64, c_vp]),
    'vms_sum_partials': (None, [c_vp, c_int, c_i64, c_f32, c_vp, c_vp]),
    'vms_elbo_plan_create': (None, [C.POINTER(ElboDesc), C.POINTER(c_vp)]),
    'vms_elbo_plan_destroy': (None, [c_vp]),
    'vms_elbo_param_count': (c_i64, [C.POINTER(ElboDesc)]),
    'vms_elbo_plan_set_mode': (None, [c_vp, c_int]),
    'vms_elbo_plan_is_fused': (c_int, [c_vp]),
    'vms_elbo_plan_tc_status': (None, [c_vp, C.POINTER(c_int)]),
    'vms_elbo_plan_path': (c_int, [c_vp, c_i64]),
    'vms_elbo_plan_invalidate': (None, [c_vp]),
    'vms_elbo_plan_set_tc_auto_batch': (None, [c_vp, c_i64]),
    'vms_elbo_plan_set_timing': (None, [c_vp, c_int]),
    'vms_elbo_plan_set_timing_every': (None, [c_vp, c_int, c_int]),
    'vms_elbo_plan_kernel_ms': (None, [c_vp, C.POINTER(c_f64), C.POINTER(c_int)]),
    'vms_elbo_train_step': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_f64, c_f64, c_f64, c_f64,
                                   c_vp]),
    'vms_peer_buffer_bytes': (c_size, [c_i64]),
    'vms_ipc_get_handle': (None, [c_vp, c_vp]),
    'vms_ipc_open_handle': (None, [c_vp, C.POINTER(c_vp)]),
    'vms_ipc_close_handle': (None, [c_vp]),
    'vms_elbo_train_step_peer': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_f64, c_f64, c_f64, c_f64, c_int,
                                        c_int, C.POINTER(c_vp), C.c_ulonglong, c_vp]),
    'vms_peer_allreduce_adam': (None, [c_int, c_int, C.POINTER(c_vp), c_i64, C.c_ulonglong, c_f32, c_vp, c_vp, c_vp, c_i64,
                                       c_f64, c_f64, c_f64, c_f64, c_vp, c_vp]),
    'vms_mc_param_count': (c_i64, [C.POINTER(McDesc)]),
    'vms_mc_plan_create': (None, [C.POINTER(McDesc), C.POINTER(c_vp)]),
    'vms_mc_plan_destroy': (None, [c_vp]),
    'vms_mc_run': (None, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, C.c_ulonglong, C.c_ulonglong, c_vp, c_vp, c_i64, c_int, c_vp,
                          c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_mc_plan_has_device_rng': (c_int, [c_vp]),
    'vms_mc_plan_set_chain_offset': (None, [c_vp, c_i64]),
    'vms_mc_run_pcg64': (None, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, C.c_ulonglong, C.c_ulonglong, C.POINTER(Pcg64Stream),
                                c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_rqs_knot_table_doubles': (c_i64, [c_int]),
    'vms_rqs_knot_table': (None, [c_vp, c_vp, c_vp, c_int, c_int, c_f32, c_f32, c_vp, c_vp]),
    'vms_mc_nb_supported': (c_int, [C.POINTER(McNbModel)]),
    'vms_mc_nb_run': (None, [C.POINTER(McNbModel), c_vp, c_vp, c_int, c_vp, C.c_ulonglong, C.c_ulonglong, c_vp,
                             C.POINTER(Pcg64Stream), c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_blockwise_log_prob_backward': (None, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_int, c_vp, c_vp, c_i64,
                                               c_vp, c_i64, c_vp]),
    'vms_std_normal_log_prob_backward': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_i64, c_vp]),
    'vms_blockwise_sample_backward': (None, [c_vp, c_i64, c_i64, c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), c_int, c_vp, c_i64, c_vp, c_i64, c_vp,
                                             c_i64, c_vp]),
    'vms_periodic_featurise_backward': (None, [c_vp, c_i64, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    'vms_add_cols': (None, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_f32, c_vp]),
    'vms_add_scalar': (None, [c_vp, c_i64, c_vp, c_f32, c_vp]),
    'vms_mul_inplace': (None, [c_vp, c_vp, c_i64, c_vp]),
    'vms_sum_all': (None, [c_vp, c_i64, c_f32, c_vp, c_vp]),
    'vms_gaa_zero_mask': (None, [c_vp, c_i64, c_vp, c_vp]),
    'vms_gaa_pair_invariants': (None, [c_vp, c_i64, c_int, c_vp, c_vp]),
    'vms_layernorm_forward': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_f32, c_int, c_vp, c_i64, c_vp, c_vp]),
    'vms_layernorm_backward_workspace': (c_size, [c_i64, c_int]),
    'vms_layernorm_backward': (None, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp,
                                      c_vp, c_vp, c_vp]),
    'vms_gaa_pair_merge': (None, [c_vp, c_i64, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp]),
    'vms_gaa_pair_merge_backward': (None, [c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_vp, c_i64, c_vp]),
    'vms_gaa_attend': (None, [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    'vms_gaa_attend_backward': (None, [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    'vms_gaa_attention_forward_supported': (c_int, [c_int, c_int, c_int]),
    'vms_gaa_attention_forward': (None, [c_vp, c_vp, c_i64, c_vp, c_i64, c_int, c_int, c_int, C.POINTER(GaaWeights), c_int,
                                         c_int, c_f32, c_vp, c_vp]),
    'vms_probe_ffma': (None, [c_int, c_int, C.POINTER(c_f64), C.POINTER(c_f64), c_vp]),
    'vms_probe_ffma2': (None, [c_int, c_int, C.POINTER(c_f64), C.POINTER(c_f64), c_vp]),
    'vms_probe_mma': (None, [c_int, c_int, c_int, c_int, c_int, C.POINTER(c_f64), C.POINTER(c_f64), c_vp]),
    'vms_elbo_forward': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_elbo_forward_backward': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
}
EXPORTS = tuple(sorted(_SIGS))

_ERRORS = {1: ValueError, 2: ValueError, 3: RuntimeError, 4: RuntimeError, 5: NotImplementedError}
_lib = None
_lock = threading.Lock()


class CaptureUnsupported(RuntimeError):
    """An operation that cannot be part of a CUDA graph (host <-> device copy, synchronisation) was issued while a training
    step was being captured: the trainer discards the capture and runs the step eagerly."""


# [0]: a CUDA-graph capture of the library's stream is in progress (set by _autodiff.Trainer)
_capturing = [False]
_NOT_CAPTURABLE = ('vms_memcpy_h2d', 'vms_memcpy_d2h', 'vms_stream_synchronize', 'vms_device_synchronize', 'vms_event_synchronize',
                   'vms_malloc_host')


class _Lib(object):
    """Attribute access returns checked callables: lib.vms_xxx(...) raises on a non-zero status."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name, (res, args) in _SIGS.items():
            fn = getattr(cdll, name)
            fn.argtypes = args
            fn.restype = c_int if res is None else res
            call = self._checked(fn, name) if res is None else fn
            if name in _NOT_CAPTURABLE:
                call = self._guarded(call, name)
            setattr(self, name, call)

    @staticmethod
    def _guarded(call, name):
        def guarded(*a):
            if _capturing[0]:
                raise CaptureUnsupported(name)
            return call(*a)

        guarded.__name__ = name
        return guarded

    def _checked(self, fn, name):
        cdll = self._cdll

        def call(*a):
            st = fn(*a)
            if st != 0:
                msg = cdll.vms_last_error().decode('utf-8', 'replace')
                raise _ERRORS.get(st, RuntimeError)('%s: %s' % (name, msg))

        call.__name__ = name
        return call


def load():
    """Load libvms_b200.so (no CUDA call is made).  Raises RuntimeError if the library has not been built."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError('vaemolsim_b200: %s is missing -- build it with `python -m vaemolsim_b200.build` '
                                   '(there is no CPU fallback)' % LIB_PATH)
            _lib = _Lib(C.CDLL(LIB_PATH))
    return _lib


# ------------------------------------------------------------------------------------------------- device context
_ctx = None


class _Context(object):

    def __init__(self):
        self.lib = load()
        n = c_int(0)
        try:
            self.lib.vms_device_count(C.byref(n))
        except RuntimeError as e:
            raise RuntimeError('vaemolsim_b200 needs a CUDA device (sm_100a); none usable: %s' % e)
        if n.value < 1:
            raise RuntimeError('vaemolsim_b200 needs a CUDA device (sm_100a); none found and there is no CPU fallback')
        dev = int(os.environ.get('LOCAL_RANK', '0')) % n.value
        self.device = dev
        self.lib.vms_set_device(dev)
        s = c_vp()
        self.lib.vms_stream_create(C.byref(s))
        self.stream = s.value
        self.pool = {}
        info = (c_i64 * 5)()
        self.lib.vms_device_info(dev, info)
        self.sm_count, self.cc = int(info[0]), (int(info[1]), int(info[2]))
        self.l2_bytes = int(info[4])

    def alloc(self, nbytes):
        size = 512
        while size < nbytes:
            size <<= 1
        free = self.pool.get(size)
        if free:
            return free.pop(), size
        p = c_vp()
        self.lib.vms_malloc(C.byref(p), size)
        return p.value, size

    def release(self, ptr, size):
        self.pool.setdefault(size, []).append(ptr)

    def synchronize(self):
        self.lib.vms_stream_synchronize(self.stream)


# Bumped by every host call that changes model parameters on the device (Dense.set_weights / assign / rebind, MADE
# set_weights, Model.set_weights, a Trainer step, the fused plans' training steps): consumers that cache something DERIVED
# from the weights (the MC notebook kernel's knot tables) compare it with the value they cached at.
_param_epoch = [0]


def bump_param_epoch():
    _param_epoch[0] += 1


def param_epoch():
    return _param_epoch[0]


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = _Context()
    return _ctx


def synchronize():
    ctx().synchronize()


# ------------------------------------------------------------------------------------------------- Tensor
class Tensor(object):
    """A contiguous row-major device array.  Mirrors the little of tf.Tensor that the reference's consumers use:
    `.numpy()`, `.shape`, `+`, `-`, unary minus, `float * t` (mcmc.py:103,109,112; losses.py:58,253)."""

    __array_priority__ = 100

    def __init__(self, shape, dtype=np.float32, _ptr=None, _base=None, _ld=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = math.prod(self.shape) * self.dtype.itemsize  # (math.prod: np.prod costs ~3 us per small tuple)
        self._base = _base
        # leading dimension (row stride in elements) of a 2-D tensor; a column view has ld > shape[1]
        self.ld = int(_ld) if _ld is not None else (self.shape[-1] if len(self.shape) >= 1 else 1)
        if _ptr is None:
            self.ptr, self._size = ctx().alloc(max(self.nbytes, 1))
        else:
            self.ptr, self._size = _ptr, None

    def __del__(self):
        if getattr(self, '_size', None) is not None and _ctx is not None:
            try:
                _ctx.release(self.ptr, self._size)
            except Exception:
                pass

    # -- construction
    @staticmethod
    def from_numpy(a, dtype=None):
        a = np.ascontiguousarray(a) if dtype is None else np.ascontiguousarray(a, dtype=dtype)
        t = Tensor(a.shape, a.dtype)
        if t.nbytes:
            c = ctx()
            c.lib.vms_memcpy_h2d(t.ptr, a.ctypes.data, t.nbytes, c.stream)
            c.synchronize()  # `a` may be a temporary
        return t

    @staticmethod
    def zeros(shape, dtype=np.float32):
        t = Tensor(shape, dtype)
        if t.nbytes:
            c = ctx()
            c.lib.vms_memset(t.ptr, 0, t.nbytes, c.stream)
        return t

    def fill_zero(self):
        if self.nbytes:
            c = ctx()
            c.lib.vms_memset(self.ptr, 0, self.nbytes, c.stream)
        return self

    # -- DLPack (the zero-copy bridge a TF / JAX / CuPy / torch caller uses: `tf.experimental.dlpack.to_dlpack(t)` ->
    #    `Tensor.from_dlpack`, `tf.experimental.dlpack.from_dlpack(t.__dlpack__())`); see _dlpack below
    def __dlpack__(self, stream=None, **kwargs):
        return _dlpack_export(self, stream)

    def __dlpack_device__(self):
        return (_kDLCUDA, ctx().device)

    @staticmethod
    def from_dlpack(obj, stream=None):
        return _dlpack_import(obj, stream)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return math.prod(self.shape)

    @property
    def contiguous(self):
        return len(self.shape) < 2 or self.ld == self.shape[-1]

    def cols(self, start, stop):
        """Column view [:, start:stop] of a 2-D tensor (no copy; shares memory, ld = parent's)."""
        if len(self.shape) != 2:
            raise ValueError('cols() needs a 2-D tensor')
        return Tensor((self.shape[0], stop - start), self.dtype, _ptr=self.ptr + start * self.dtype.itemsize,
                      _base=self, _ld=self.ld)

    def contig(self):
        """Contiguous copy of a column view (or self when already contiguous)."""
        if self.contiguous:
            return self
        t = Tensor(self.shape, self.dtype)
        c = ctx()
        it = self.dtype.itemsize
        c.lib.vms_memcpy2d_d2d(t.ptr, self.shape[1] * it, self.ptr, self.ld * it, self.shape[1] * it, self.shape[0],
                               c.stream)
        _record_copy(self, t)
        return t

    def assign_cols(self, start, src):
        """self[:, start:start+src.shape[1]] = src  (device-side strided copy)."""
        c = ctx()
        it = self.dtype.itemsize
        c.lib.vms_memcpy2d_d2d(self.ptr + start * it, self.ld * it, src.ptr, src.ld * it, src.shape[1] * it,
                               src.shape[0], c.stream)

    def numpy(self):
        if not self.contiguous:
            return self.contig().numpy()
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            c = ctx()
            c.lib.vms_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, c.stream)
            c.synchronize()
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        n = self.size
        shape = list(shape)
        if -1 in shape:
            i = shape.index(-1)
            rest = math.prod([s for s in shape if s != -1])
            shape[i] = n // rest if rest else 0
        if math.prod(shape) != n:
            raise ValueError('cannot reshape tensor of size %d into shape %s' % (n, tuple(shape)))
        return Tensor(shape, self.dtype, _ptr=self.ptr, _base=self)

    def copy(self):
        t = Tensor(self.shape, self.dtype)
        if self.nbytes:
            c = ctx()
            c.lib.vms_memcpy_d2d(t.ptr, self.ptr, self.nbytes, c.stream)
        _record_copy(self, t)
        return t

    # -- arithmetic (float32 only; anything else goes through NumPy on the host)
    def _axpby(self, other, a, b):
        c = ctx()
        if not self.contiguous:
            return self.contig()._axpby(other, a, b)
        if isinstance(other, Tensor) and not other.contiguous:
            other = other.contig()
        if isinstance(other, Tensor):
            if other.shape != self.shape or self.dtype != np.float32 or other.dtype != np.float32:
                return Tensor.from_numpy(a * self.numpy() + b * other.numpy())
            out = Tensor(self.shape)
            c.lib.vms_axpby(self.ptr, other.ptr, a, b, self.size, out.ptr, c.stream)
            _record_axpby(self, other, a, b, out)
            return out
        return Tensor.from_numpy((a * self.numpy() + b * np.asarray(other)).astype(self.dtype))

    def __add__(self, other):
        return self._axpby(other, 1.0, 1.0)

    __radd__ = __add__

    def __sub__(self, other):
        return self._axpby(other, 1.0, -1.0)

    def __rsub__(self, other):
        return self._axpby(other, -1.0, 1.0)

    def __neg__(self):
        return self.__mul__(-1.0)

    def __mul__(self, k):
        if not self.contiguous:
            return self.contig().__mul__(k)
        if isinstance(k, Tensor) or np.ndim(k) != 0 or self.dtype != np.float32:
            return Tensor.from_numpy((self.numpy() * np.asarray(k)).astype(self.dtype))
        c = ctx()
        out = Tensor(self.shape)
        c.lib.vms_axpby(self.ptr, None, float(k), 0.0, self.size, out.ptr, c.stream)
        _record_axpby(self, None, float(k), 0.0, out)
        return out

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self.__mul__(1.0 / k)

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __repr__(self):
        return 'Tensor(shape=%s, dtype=%s, device=cuda:%d)' % (self.shape, self.dtype.name, ctx().device)


# ------------------------------------------------------------------------------------------------- DLPack
# north_star: "a ctypes/DLPack-bridged TF custom op".  The capsules are built by hand with ctypes (no torch, no cupy): a
# `DLManagedTensor` (dlpack.h, v0.x ABI -- what tf.experimental.dlpack / torch.utils.dlpack / cupy exchange) whose deleter
# drops the reference that keeps the producing Tensor alive.  Import wraps the foreign device pointer without copying and
# calls the producer's deleter when the wrapping Tensor dies.
_kDLCUDA = 2
_DL_CODES = {'i': 0, 'u': 1, 'f': 2}
_DL_KINDS = {0: 'i', 1: 'u', 2: 'f'}


class _DLDevice(C.Structure):
    _fields_ = [('device_type', C.c_int32), ('device_id', C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [('code', C.c_uint8), ('bits', C.c_uint8), ('lanes', C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [('data', C.c_void_p), ('device', _DLDevice), ('ndim', C.c_int32), ('dtype', _DLDataType),
                ('shape', C.POINTER(C.c_int64)), ('strides', C.POINTER(C.c_int64)), ('byte_offset', C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DLDeleter = C.CFUNCTYPE(None, C.POINTER(_DLManagedTensor))
_DLManagedTensor._fields_ = [('dl_tensor', _DLTensor), ('manager_ctx', C.c_void_p), ('deleter', _DLDeleter)]

_dl_live = {}  # id -> (managed struct, shape array, strides array, Tensor): kept alive until the consumer's deleter runs
_pyapi = C.pythonapi
_pyapi.PyCapsule_New.restype = C.py_object
_pyapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_pyapi.PyCapsule_IsValid.restype = C.c_int
_pyapi.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_pyapi.PyCapsule_GetPointer.restype = C.c_void_p
_pyapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_pyapi.PyCapsule_SetName.restype = C.c_int
_pyapi.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]


@_DLDeleter
def _dl_deleter(mt):
    _dl_live.pop(C.addressof(mt.contents), None)


_PyCapsuleDestructor = C.CFUNCTYPE(None, C.c_void_p)


@_PyCapsuleDestructor
def _dl_capsule_destructor(capsule):
    # a capsule nobody consumed still carries the name "dltensor": run the deleter ourselves (DLPack protocol)
    cap = C.cast(capsule, C.py_object)
    if _pyapi.PyCapsule_IsValid(cap, b'dltensor'):
        ptr = _pyapi.PyCapsule_GetPointer(cap, b'dltensor')
        _dl_live.pop(ptr, None)


def _dlpack_export(t, stream=None):
    """PyCapsule("dltensor") over `t`'s device memory.  Work queued on the library stream is completed first (the
    consumer's stream is unknown to a ctypes producer), so the consumer may read immediately."""
    if t.dtype.kind not in _DL_CODES:
        raise TypeError('__dlpack__: unsupported dtype %s' % t.dtype)
    ctx().synchronize()
    nd = len(t.shape)
    shape = (C.c_int64 * max(nd, 1))(*t.shape)
    strides = None
    if nd == 2 and not t.contiguous:
        strides = (C.c_int64 * 2)(t.ld, 1)
    mt = _DLManagedTensor()
    mt.dl_tensor.data = t.ptr
    mt.dl_tensor.device = _DLDevice(_kDLCUDA, ctx().device)
    mt.dl_tensor.ndim = nd
    mt.dl_tensor.dtype = _DLDataType(_DL_CODES[t.dtype.kind], t.dtype.itemsize * 8, 1)
    mt.dl_tensor.shape = C.cast(shape, C.POINTER(C.c_int64))
    mt.dl_tensor.strides = C.cast(strides, C.POINTER(C.c_int64)) if strides is not None else None
    mt.dl_tensor.byte_offset = 0
    mt.manager_ctx = None
    mt.deleter = _dl_deleter
    _dl_live[C.addressof(mt)] = (mt, shape, strides, t)
    return _pyapi.PyCapsule_New(C.addressof(mt), b'dltensor', C.cast(_dl_capsule_destructor, C.c_void_p))


class _Foreign(object):
    """Owner of an imported DLManagedTensor: calls the producer's deleter when the last view dies."""

    def __init__(self, mt_ptr):
        self.mt_ptr = mt_ptr

    def __del__(self):
        try:
            mt = C.cast(self.mt_ptr, C.POINTER(_DLManagedTensor))
            if mt.contents.deleter:
                mt.contents.deleter(mt)
        except Exception:
            pass


def _dlpack_import(obj, stream=None):
    """Zero-copy Tensor over a DLPack producer (an object with `__dlpack__`, or a "dltensor" capsule): CUDA device memory on
    this context's device, compact row-major (or a 2-D row-strided view)."""
    cap = obj.__dlpack__() if hasattr(obj, '__dlpack__') else obj
    if not _pyapi.PyCapsule_IsValid(cap, b'dltensor'):
        raise ValueError('from_dlpack: not a "dltensor" capsule (already consumed?)')
    ptr = _pyapi.PyCapsule_GetPointer(cap, b'dltensor')
    mt = C.cast(ptr, C.POINTER(_DLManagedTensor)).contents
    dl = mt.dl_tensor
    if dl.device.device_type != _kDLCUDA:
        raise ValueError('from_dlpack: only CUDA device memory can be wrapped (device_type %d); there is no host path'
                         % dl.device.device_type)
    if dl.device.device_id != ctx().device:
        raise ValueError('from_dlpack: tensor lives on cuda:%d, this context is cuda:%d' % (dl.device.device_id, ctx().device))
    if dl.dtype.lanes != 1 or dl.dtype.code not in _DL_KINDS:
        raise TypeError('from_dlpack: unsupported dtype (code %d, lanes %d)' % (dl.dtype.code, dl.dtype.lanes))
    dtype = np.dtype('%s%d' % (_DL_KINDS[dl.dtype.code], dl.dtype.bits // 8))
    shape = tuple(int(dl.shape[i]) for i in range(dl.ndim))
    ld = None
    if dl.strides:
        strides = tuple(int(dl.strides[i]) for i in range(dl.ndim))
        compact, acc = [], 1
        for n in reversed(shape):
            compact.append(acc)
            acc *= n
        compact = tuple(reversed(compact))
        if strides != compact and not all(n <= 1 or a == b for n, a, b in zip(shape, strides, compact)):
            if dl.ndim == 2 and strides[1] == 1 and strides[0] >= shape[1]:
                ld = strides[0]
            else:
                raise ValueError('from_dlpack: only compact row-major or 2-D row-strided tensors (strides %s)' % (strides, ))
    _pyapi.PyCapsule_SetName(cap, b'used_dltensor')  # ownership of the managed tensor is ours now
    owner = _Foreign(ptr)
    return Tensor(shape, dtype, _ptr=int(dl.data or 0) + int(dl.byte_offset), _base=owner, _ld=ld)


# ------------------------------------------------------------------------------------------------- tape hooks
def _tape():
    from . import _autodiff
    return _autodiff.Tape.active(), _autodiff


def _record_copy(src, dst):
    """dst is a copy of src (contig / copy): g_src += g_dst."""
    if src.dtype != np.float32:
        return
    tp, ad = _tape()
    if tp is None:
        return

    def bw():
        if tp.has(dst):
            ad.add_into(tp.grad(src), tp.grad(dst), 1.0)

    tp.record(bw)


def _record_axpby(x, y, a, b, out):
    """out = a x + b y (y may be None): g_x += a g_out, g_y += b g_out."""
    tp, ad = _tape()
    if tp is None:
        return

    def bw():
        if not tp.has(out):
            return
        g = tp.grad(out)
        ad.add_into(tp.grad(x), g, a)
        if y is not None:
            ad.add_into(tp.grad(y), g, b)

    tp.record(bw)


def as_tensor(x, dtype=np.float32):
    """Device tensor from a Tensor / ndarray / nested list (host data is uploaded)."""
    if isinstance(x, Tensor):
        if x.dtype != np.dtype(dtype):
            return Tensor.from_numpy(x.numpy().astype(dtype))
        return x
    return Tensor.from_numpy(np.asarray(x, dtype=dtype))


def launch_count():
    return int(load().vms_launch_count())
