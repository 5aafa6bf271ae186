"""Reverse mode for the op-by-op path: a tape of the kernels the host launched, replayed backwards.

The reference trains every composition through TF autodiff (`tf.GradientTape` inside Keras `fit`, tests/test_models.py:
189-262, models.py:85-139).  The fused plans (`csrc/elbo*.cu`) cover the headline family; everything else -- von Mises
encoders, periodic FCDeepNN, blockwise / autoregressive / MAF-flowed decoders, FlowModel NLL training -- runs here: while a
`Tape` is active every device op of `_protocols.py` / `_abi.py` / `mappings.py` / `losses.py` records a closure that
launches its reverse-mode kernel (`vms_dense_backward`, `vms_rqs_apply_backward`, `vms_blockwise_log_prob_backward`,
`vms_blockwise_sample_backward`, `vms_periodic_featurise_backward`, ...).  Gradients live in device tensors shaped like the
ROOT allocation of each tensor; a view's gradient is the same view of its root's gradient, so column slices and reshapes
need no ops of their own and every backward kernel ACCUMULATES.

`Trainer` is the Keras `fit` / `train_step` replacement for this path: forward under a tape, backward, MADE masks on the
masked kernels' gradients, Keras Adam per parameter tensor (`vms_adam_step`).
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import Tensor, ctx


class Tape(object):
    """Records backward closures in forward order; `backward(loss)` seeds d loss = 1 and replays them in reverse."""
    current = None

    def __init__(self):
        self.ops = []
        self.grads = {}   # id(root tensor) -> (root, gradient of the root allocation)
        self._paused = 0

    def __enter__(self):
        self._prev, Tape.current = Tape.current, self
        return self

    def __exit__(self, *exc):
        Tape.current = self._prev
        return False

    # -- recording
    @staticmethod
    def active():
        t = Tape.current
        return t if (t is not None and not t._paused) else None

    def record(self, fn):
        self.ops.append(fn)

    class _Pause(object):
        def __init__(self, tape):
            self.tape = tape

        def __enter__(self):
            self.tape._paused += 1

        def __exit__(self, *exc):
            self.tape._paused -= 1
            return False

    def paused(self):
        return Tape._Pause(self)

    # -- gradients
    @staticmethod
    def _root(t):
        r = t
        while isinstance(r._base, Tensor):
            r = r._base
        return r

    def has(self, t):
        return id(self._root(t)) in self.grads

    def grad(self, t):
        """Gradient tensor shaped (and strided) like `t`: a view into the gradient of t's root allocation."""
        r = self._root(t)
        ent = self.grads.get(id(r))
        if ent is None:
            g = Tensor.zeros(r.shape if r.nbytes else (1, ), np.float32)
            self.grads[id(r)] = ent = (r, g)
        g = ent[1]
        if t is r:
            return g
        return Tensor(t.shape, np.float32, _ptr=g.ptr + (t.ptr - r.ptr), _base=g, _ld=t.ld)

    def release(self):
        """Drops the recorded closures and gradient tensors.  The closures reference the tape (and each other's tensors), a
        reference cycle that only Python's cyclic collector would break -- much later: until then every intermediate and
        gradient tensor of the step stays allocated, the size-class pool of `_abi` runs dry and the next step pays a
        synchronising cudaMalloc per tensor (measured: 4 ms per step instead of 1.8 ms for a 1-D flow model)."""
        self.ops = []
        self.grads = {}

    def backward(self, loss):
        c = ctx()
        g = self.grad(loss)  # (zero-initialised)
        c.lib.vms_add_scalar(g.ptr, max(loss.size, 1), None, 1.0, c.stream)  # d loss = 1, on the device (capturable)
        for fn in reversed(self.ops):
            fn()


def add_into(dst, src, alpha=1.0):
    """dst += alpha * src for same-shaped (possibly column-strided) float32 tensors."""
    c = ctx()
    if dst.ndim == 2:
        c.lib.vms_add_cols(dst.ptr, dst.ld, src.ptr, src.ld, dst.shape[0], dst.shape[1], float(alpha), c.stream)
    else:
        n = dst.size
        c.lib.vms_add_cols(dst.ptr, n, src.ptr, n, 1, n, float(alpha), c.stream)


class Trainer(object):
    """Keras-style training of an arbitrary model of this package: `step(fn)` runs fn() under a tape (fn returns the scalar
    loss Tensor), back-propagates and applies Adam to every weight of `model.weights` that received a gradient."""

    def __init__(self, model, optimizer, group=None):
        self.model, self.opt = model, optimizer
        self.state = {}  # id(weight) -> (weight, m, v)
        self.t = 0
        # data-parallel training of ANY model (north_star: one gradient allreduce per step): every rank runs the step on its
        # shard of the batch, the gradients of all weights travel as ONE flat float32 buffer through an NCCL sum-allreduce
        # on the library's stream, Adam then applies sum / world -- identical on every rank, so replicas stay in step
        self.group = group if (group is not None and group.world > 1) else None
        self._flat = None

    # ---- CUDA-graph replay of the whole step (forward, reverse mode, Adam) for fixed-shape batches
    # The tape path is launch-bound: ~100 small kernels per step, ~10 us of Python each.  After two eager steps the trainer
    # captures one step on the library's stream -- inputs in static device buffers, Adam's step count in device memory
    # (vms_adam_step_multi_dev) -- and from then on a step is: copy the batch into the input buffers, ONE graph launch.
    # Anything that cannot be captured (host-drawn noise, host <-> device copies or synchronisation inside an op: raises
    # _abi.CaptureUnsupported; a failed capture) makes the trainer fall back to eager steps for good.  VMS_TAPE_GRAPH=0
    # disables it.
    graph_warmup = 2

    def step_arrays(self, x, y, loss_of):
        """One optimiser step on host arrays x, y; `loss_of(xb, yb)` builds the scalar loss Tensor from device tensors."""
        import os
        c = ctx()
        key = (x.shape, y.shape)
        graphs = self.__dict__.setdefault('_graphs', {})
        g = graphs.get(key)
        if g is not None:
            for dst, src in ((g['x'], x), (g['y'], y)):
                c.lib.vms_memcpy_h2d(dst.ptr, src.ctypes.data, src.nbytes, c.stream)
            tbuf = np.array([self.t], np.int64)
            c.lib.vms_memcpy_h2d(g['t_dev'].ptr, tbuf.ctypes.data, 8, c.stream)
            c.lib.vms_graph_launch(g['exec'], g['n_kernels'], c.stream)
            self.t += 1
            _abi.bump_param_epoch()
            return g['loss']
        eligible = (self.group is None and self.t >= self.graph_warmup and not self.__dict__.get('_graph_off', False) and
                    os.environ.get('VMS_TAPE_GRAPH', '1') != '0' and len(graphs) < 4)
        xb, yb = Tensor.from_numpy(x), Tensor.from_numpy(y)
        if not eligible:
            return self.step(lambda: loss_of(xb, yb))
        g = self._capture(xb, yb, loss_of)
        if g is None:
            self._graph_off = True
            return self.step(lambda: loss_of(xb, yb))
        graphs[key] = g
        tbuf = np.array([self.t], np.int64)
        c.lib.vms_memcpy_h2d(g['t_dev'].ptr, tbuf.ctypes.data, 8, c.stream)
        c.lib.vms_graph_launch(g['exec'], g['n_kernels'], c.stream)  # the captured kernels have not run yet: this is the step
        self.t += 1
        _abi.bump_param_epoch()
        return g['loss']

    def _capture(self, xb, yb, loss_of):
        c = ctx()
        from . import _protocols
        todo = None
        c.synchronize()
        t_dev, lr_dev = Tensor.zeros((2, ), np.int32), Tensor.zeros((1, ))  # (int64 step count as two int32 words)
        c.lib.vms_graph_begin_capture(c.stream)
        _abi._capturing[0] = True
        ok, tape, loss = True, None, None
        try:
            with Tape() as tape:
                loss = loss_of(xb, yb)
                tape.backward(loss)
            table, keep = [], []
            seen = set()
            for w in self.model.weights:
                if id(w) in seen or not tape.has(w):
                    continue
                seen.add(id(w))
                gr = tape.grad(w)
                if not (w.contiguous and gr.contiguous):
                    raise _abi.CaptureUnsupported('strided weight view')
                st = self.state.get(id(w))
                if st is None:
                    raise _abi.CaptureUnsupported('a weight without optimiser state (it received no gradient in the eager steps)')
                mask = getattr(w, '_grad_mask', None)
                table.append(_abi.AdamTensor(w.ptr, gr.ptr, None if mask is None else mask.ptr, st[1].ptr, st[2].ptr, w.size))
                keep.append(gr)
            o = self.opt
            arr = (_abi.AdamTensor * len(table))(*table)
            c.lib.vms_adam_step_multi_dev(arr, len(table), 1.0, t_dev.ptr, lr_dev.ptr, o.learning_rate, o.beta_1, o.beta_2,
                                          o.epsilon, c.stream)
        except Exception:
            ok = False
        finally:
            _abi._capturing[0] = False
        if not ok:
            c.lib.vms_graph_abort_capture(c.stream)
            if tape is not None:
                tape.release()
            return None
        ex, nk = C.c_void_p(), C.c_int(0)
        try:
            c.lib.vms_graph_end_capture(c.stream, C.byref(ex), C.byref(nk))
        except Exception:
            tape.release()
            return None
        # the graph reads and writes every tensor of the captured step: they stay allocated as long as the graph lives
        return {'exec': ex.value, 'n_kernels': nk.value, 'x': xb, 'y': yb, 'loss': loss, 't_dev': t_dev, 'lr_dev': lr_dev,
                'keep': (tape, keep, table)}

    def step(self, fn):
        c = ctx()
        from . import _protocols
        _protocols._dp_group = self.group  # batch-normalisation layers: cross-replica batch statistics while this step runs
        try:
            with Tape() as tape:
                loss = fn()
                tape.backward(loss)
        finally:
            _protocols._dp_group = None
        self.t += 1
        seen = set()
        todo = []  # (weight, gradient, mask or None)
        for w in self.model.weights:
            if id(w) in seen or not (tape.has(w) or self.group is not None):
                continue
            seen.add(id(w))
            g = tape.grad(w)  # (data-parallel: every rank exchanges every weight, a zero gradient if it received none)
            todo.append((w, g, getattr(w, '_grad_mask', None)))  # mask: tfp AutoregressiveNetwork, masked entries stay zero
        scale = 1.0
        if self.group is not None:
            total = sum(w.size for w, _, _ in todo)
            if self._flat is None or self._flat.size != total:
                self._flat = Tensor((total, ))
            off, flat = 0, []
            for w, g, mask in todo:
                if mask is not None:  # masked before the exchange, so that every rank sends the constrained gradient
                    c.lib.vms_mul_inplace(g.ptr, mask.ptr, g.size, c.stream)
                gc = g if g.contiguous else g.contig()
                dst = Tensor(w.shape, np.float32, _ptr=self._flat.ptr + 4 * off, _base=self._flat)
                c.lib.vms_memcpy_d2d(dst.ptr, gc.ptr, 4 * w.size, c.stream)
                flat.append((w, dst, None))
                off += w.size
            self.group.allreduce_sum_device_(self._flat.ptr, total, c.stream)
            todo, scale = flat, 1.0 / self.group.world
        o = self.opt
        table, keep = [], []
        for w, g, mask in todo:
            st = self.state.get(id(w))
            if st is None:
                st = self.state[id(w)] = (w, Tensor.zeros(w.shape), Tensor.zeros(w.shape))
            if w.contiguous and g.contiguous:
                # every contiguous weight tensor of the model in ONE launch (vms_adam_step_multi: the path is launch-bound)
                table.append(_abi.AdamTensor(w.ptr, g.ptr, None if mask is None else mask.ptr, st[1].ptr, st[2].ptr, w.size))
                keep.append(g)
            else:  # a weight that is a strided view of a flat buffer: update a contiguous copy and write it back
                if mask is not None:
                    c.lib.vms_mul_inplace(g.ptr, mask.ptr, g.size, c.stream)
                wc, gc = w.contig(), g.contig()
                c.lib.vms_adam_step(wc.ptr, gc.ptr, 1, scale, st[1].ptr, st[2].ptr, w.size, self.t, o.learning_rate, o.beta_1,
                                    o.beta_2, o.epsilon, c.stream)
                it = 4
                c.lib.vms_memcpy2d_d2d(w.ptr, w.ld * it, wc.ptr, wc.ld * it, w.shape[-1] * it, w.shape[0], c.stream)
        if table:
            arr = (_abi.AdamTensor * len(table))(*table)
            c.lib.vms_adam_step_multi(arr, len(table), scale, self.t, o.learning_rate, o.beta_1, o.beta_2, o.epsilon, c.stream)
        _abi.bump_param_epoch()
        tape.release()
        return loss
