"""ctypes binding of libvms_b200.so (include/vms_b200.h) and a minimal device `Tensor`.

This is the only module that touches the shared library.  There is deliberately NO fallback: if the library is
missing or no CUDA device is present, the first compute call raises (the reference's TF ops are replaced by
sm_100a kernels, not by NumPy).  NumPy is used for host-side staging only.
"""
import ctypes as C
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
    'vms_mc_accept': (None, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_energy_quadratic': (None, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    'vms_adam_step': (None, [c_vp, c_vp, c_int, c_f32, c_vp, c_vp, c_i64, c_i64, c_f64, c_f64, c_f64, c_f64, c_vp]),
    'vms_sum_partials': (None, [c_vp, c_int, c_i64, c_f32, c_vp, c_vp]),
    'vms_elbo_plan_create': (None, [C.POINTER(ElboDesc), C.POINTER(c_vp)]),
    'vms_elbo_plan_destroy': (None, [c_vp]),
    'vms_elbo_param_count': (c_i64, [C.POINTER(ElboDesc)]),
    'vms_elbo_plan_set_mode': (None, [c_vp, c_int]),
    'vms_elbo_plan_is_fused': (c_int, [c_vp]),
    'vms_elbo_plan_tc_status': (None, [c_vp, C.POINTER(c_int)]),
    'vms_elbo_plan_path': (c_int, [c_vp, c_i64]),
    'vms_elbo_plan_set_tc_auto_batch': (None, [c_vp, c_i64]),
    'vms_elbo_plan_set_timing': (None, [c_vp, c_int]),
    'vms_elbo_plan_kernel_ms': (None, [c_vp, C.POINTER(c_f64), C.POINTER(c_int)]),
    'vms_elbo_train_step': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_f64, c_f64, c_f64, c_f64,
                                   c_vp]),
    'vms_peer_buffer_bytes': (c_size, [c_i64]),
    'vms_ipc_get_handle': (None, [c_vp, c_vp]),
    'vms_ipc_open_handle': (None, [c_vp, C.POINTER(c_vp)]),
    'vms_ipc_close_handle': (None, [c_vp]),
    'vms_peer_allreduce_adam': (None, [c_int, c_int, C.POINTER(c_vp), c_i64, C.c_ulonglong, c_f32, c_vp, c_vp, c_vp, c_i64,
                                       c_f64, c_f64, c_f64, c_f64, c_vp, c_vp]),
    'vms_mc_param_count': (c_i64, [C.POINTER(McDesc)]),
    'vms_mc_plan_create': (None, [C.POINTER(McDesc), C.POINTER(c_vp)]),
    'vms_mc_plan_destroy': (None, [c_vp]),
    'vms_mc_run': (None, [c_vp, c_vp, c_vp, c_vp, c_int, c_vp, C.c_ulonglong, C.c_ulonglong, c_vp, c_vp, c_i64, c_int, c_vp,
                          c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_elbo_forward': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    'vms_elbo_forward_backward': (None, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
}
EXPORTS = tuple(sorted(_SIGS))

_ERRORS = {1: ValueError, 2: ValueError, 3: RuntimeError, 4: RuntimeError, 5: NotImplementedError}
_lib = None
_lock = threading.Lock()


class _Lib(object):
    """Attribute access returns checked callables: lib.vms_xxx(...) raises on a non-zero status."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name, (res, args) in _SIGS.items():
            fn = getattr(cdll, name)
            fn.argtypes = args
            fn.restype = c_int if res is None else res
            setattr(self, name, self._checked(fn, name) if res is None else fn)

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
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
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

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

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
            rest = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape[i] = n // rest if rest else 0
        if int(np.prod(shape, dtype=np.int64)) != n:
            raise ValueError('cannot reshape tensor of size %d into shape %s' % (n, tuple(shape)))
        return Tensor(shape, self.dtype, _ptr=self.ptr, _base=self)

    def copy(self):
        t = Tensor(self.shape, self.dtype)
        if self.nbytes:
            c = ctx()
            c.lib.vms_memcpy_d2d(t.ptr, self.ptr, self.nbytes, c.stream)
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
        return out

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self.__mul__(1.0 / k)

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __repr__(self):
        return 'Tensor(shape=%s, dtype=%s, device=cuda:%d)' % (self.shape, self.dtype.name, ctx().device)


def as_tensor(x, dtype=np.float32):
    """Device tensor from a Tensor / ndarray / nested list (host data is uploaded)."""
    if isinstance(x, Tensor):
        if x.dtype != np.dtype(dtype):
            return Tensor.from_numpy(x.numpy().astype(dtype))
        return x
    return Tensor.from_numpy(np.asarray(x, dtype=dtype))


def launch_count():
    return int(load().vms_launch_count())
