"""Mappings -- host-side mirror of the hot-path part of `vaemolsim/mappings.py` over sm_100a kernels.

In scope: `FCDeepNN` (mappings.py:18-166) and `DistanceSelection` (mappings.py:308-477), same
constructor keywords, call signatures and error behaviour.  `AttentionBlock`, `ParticleEmbedding`,
`LocalParticleDescriptors` (un-vendored geometric-algebra-attention dependency) and the CG template layers are out of
scope.
"""
import numpy as np

from . import _protocols as P
from ._abi import Tensor, as_tensor, ctx


class RaggedTensor(object):
    """The little of tf.RaggedTensor that `DistanceSelection` consumes: flat values + row_splits."""

    def __init__(self, values, row_splits):
        self.values = np.ascontiguousarray(values, np.float32)
        self.row_splits = np.ascontiguousarray(row_splits, np.int64)
        if self.row_splits.ndim != 1 or self.row_splits[0] != 0 or self.row_splits[-1] != self.values.shape[0]:
            raise ValueError('row_splits must start at 0 and end at len(values)')

    @staticmethod
    def from_rows(rows, inner=None):
        rows = [np.asarray(r, np.float32) for r in rows]
        if inner is None:
            inner = next((r.shape[-1] for r in rows if r.ndim == 2), 3)
        rows = [r.reshape(-1, inner) for r in rows]
        lens = np.array([r.shape[0] for r in rows], np.int64)
        vals = np.concatenate(rows, axis=0) if rows else np.zeros((0, inner), np.float32)
        return RaggedTensor(vals, np.concatenate([[0], np.cumsum(lens)]))

    @staticmethod
    def from_row_lengths(values, row_lengths):
        return RaggedTensor(values, np.concatenate([[0], np.cumsum(np.asarray(row_lengths, np.int64))]))

    @property
    def shape(self):
        return (len(self.row_splits) - 1, None, self.values.shape[-1])


class FCDeepNN(P.Layer):
    """mappings.py:18-166: fully connected network; periodic dofs enter as (cos, sin) pairs (mappings.py:144-149)."""

    def __init__(self, target_shape, hidden_dim=200, periodic_dofs=False, batch_norm=False, name='mapping',
                 activation='relu', kernel_initializer='glorot_uniform', **kwargs):
        super(FCDeepNN, self).__init__(name=name, **kwargs)
        try:
            self.target_shape = tuple(target_shape)
        except TypeError:
            self.target_shape = (target_shape, )
        self.hidden_dim = [hidden_dim] if isinstance(hidden_dim, (int, np.integer)) else hidden_dim
        self.periodic_dofs = periodic_dofs
        self.batch_norm = batch_norm
        self.activation = activation
        self.kernel_initializer = kernel_initializer

    def build(self, input_shape):
        n_in = int(np.prod(input_shape[1:]))
        if isinstance(self.periodic_dofs, bool):
            self.any_periodic = self.periodic_dofs
            self.periodic_dofs = np.array([self.periodic_dofs] * n_in, dtype=bool)
        else:
            if len(self.periodic_dofs) != n_in:
                raise ValueError("Shape of periodic_dofs (%i) should match flattened input (%i)." %
                                 (len(self.periodic_dofs), n_in))
            self.any_periodic = bool(np.any(self.periodic_dofs))
            self.periodic_dofs = np.asarray(self.periodic_dofs, dtype=bool)
        self._periodic_dev = Tensor.from_numpy(self.periodic_dofs.astype(np.uint8)) if self.any_periodic else None
        self.layer_list = []
        width = n_in + int(np.sum(self.periodic_dofs))
        for hd in self.hidden_dim:
            lay = P.Dense(hd, activation=self.activation, kernel_initializer=self.kernel_initializer)
            lay.build((None, width))
            lay.built = True
            self.layer_list.append(lay)
            width = hd
            if self.batch_norm:  # mappings.py:113-114
                bn = P.KerasBatchNormalization()
                bn.build((None, width))
                bn.built = True
                self.layer_list.append(bn)
        last = P.Dense(int(np.prod(self.target_shape)), activation=None, kernel_initializer=self.kernel_initializer)
        last.build((None, width))
        last.built = True
        self.layer_list.append(last)
        self.layer_list.append(P.Reshape(self.target_shape))  # mappings.py:123

    def call(self, inputs, training=False):
        x = as_tensor(inputs).contig()
        out = x.reshape(x.shape[0], -1)
        if self.any_periodic:
            c = ctx()
            feat = Tensor((out.shape[0], out.shape[1] + int(np.sum(self.periodic_dofs))))
            c.lib.vms_periodic_featurise(out.ptr, out.shape[0], out.shape[1], self._periodic_dev.ptr, feat.ptr, c.stream)
            from . import _autodiff
            tp = _autodiff.Tape.active()
            if tp is not None:
                src, n_per, per = out, int(np.sum(self.periodic_dofs)), self._periodic_dev

                def bw():
                    if tp.has(feat):
                        c.lib.vms_periodic_featurise_backward(src.ptr, src.shape[0], src.shape[1], per.ptr, n_per,
                                                              tp.grad(feat).ptr, tp.grad(src).ptr, c.stream)

                tp.record(bw)
            out = feat
        for layer in self.layer_list:
            out = layer.call(out) if isinstance(layer, P.Dense) else layer.call(out, training=training)
        return out

    def get_config(self):
        config = super(FCDeepNN, self).get_config()
        config.update({"target_shape": self.target_shape, "hidden_dim": self.hidden_dim,
                       "periodic_dofs": self.periodic_dofs, "batch_norm": self.batch_norm})
        return config


class DistanceSelection(P.Layer):
    """mappings.py:308-477: the `max_included` nearest particles (minimum image) around each reference site, masked by
    the cutoff and zero padded.  One kernel launch (`csrc/distsel.cu`), bit-exact with TF's float32 op order."""

    def __init__(self, cutoff, max_included=50, box_lengths=None, name='dist_select', **kwargs):
        super(DistanceSelection, self).__init__(name=name, **kwargs)
        self.cutoff = cutoff
        self.sq_cut = cutoff**2
        self.max_included = max_included
        if box_lengths is not None:
            self.box_lengths = np.asarray(box_lengths, np.float32).reshape(1, 1, 3)
            self._box_dev = Tensor.from_numpy(self.box_lengths.reshape(3))
        else:
            self.box_lengths = None

    @staticmethod
    def _ragged(x, inner=None):
        if isinstance(x, RaggedTensor):
            return x
        if isinstance(x, (list, tuple)):
            return RaggedTensor.from_rows(x, inner)
        return None

    def select_from_frame(self, frame, ref, box_lengths=None, particle_info=None, return_indices=False):
        """Extension (no reference counterpart): B reference sites selecting from ONE frame.  `frame` [N, 3] (and
        `particle_info` [N, P]) are the particles of a single configuration, `ref` [B, 3] the sites; the result is exactly
        `call(tile(frame, B), ref, ...)` -- what the reference API requires a caller to materialise -- without the B copies:
        the frame is uploaded and read from HBM once (it stays in L2 across rows) instead of B times.  This is the shape of
        backmapping a simulation box: thousands of coarse-grained sites, one box of particles."""
        c = ctx()
        frame = as_tensor(frame).contig()
        if frame.ndim == 3 and frame.shape[0] == 1:
            frame = frame.reshape(frame.shape[1], 3)
        if frame.ndim != 2 or frame.shape[1] != 3:
            raise ValueError('frame must have shape (N_particles, 3); got %s' % (frame.shape, ))
        N = frame.shape[0]
        ref = as_tensor(ref).contig()
        if ref.size % 3 != 0:
            raise ValueError('ref must have shape (N_sites, 3); got %s' % (ref.shape, ))
        B = ref.size // 3
        box_ptr, per_row = None, 0
        if box_lengths is not None:
            box = as_tensor(box_lengths).contig()
            if box.size == 3:
                box_ptr = box.ptr
            elif box.size == B * 3:
                box_ptr, per_row = box.ptr, 1
            else:
                raise ValueError('box_lengths must have shape (3,) or (N_sites, 3); got %s' % (box.shape, ))
        elif self.box_lengths is not None:
            box_ptr = self._box_dev.ptr
        k = int(self.max_included)
        info_ptr, P_, out_info = None, 0, None
        if particle_info is not None:
            info = as_tensor(particle_info).contig()
            if info.ndim == 3 and info.shape[0] == 1:
                info = info.reshape(info.shape[1], info.shape[2])
            if info.ndim != 2 or info.shape[0] != N:
                raise ValueError('particle_info must have shape (N_particles, P) like frame')
            P_, info_ptr = info.shape[1], info.ptr
            out_info = Tensor((B, k, P_))
        out_xyz = Tensor((B, k, 3))
        out_idx = Tensor((B, k), np.int32) if return_indices else None
        c.lib.vms_dist_select_frame(frame.ptr, N, ref.ptr, B, box_ptr, per_row, float(np.float32(self.sq_cut)), k, info_ptr, P_,
                                    out_xyz.ptr, None if out_info is None else out_info.ptr,
                                    None if out_idx is None else out_idx.ptr, c.stream)
        outs = [out_xyz] + ([out_info] if out_info is not None else []) + ([out_idx] if return_indices else [])
        return outs[0] if len(outs) == 1 else tuple(outs)

    def call(self, coords, ref, box_lengths=None, particle_info=None, return_indices=False):
        c = ctx()
        rag = self._ragged(coords, 3)
        if rag is not None:
            B, N = len(rag.row_splits) - 1, 0
            cvals = Tensor.from_numpy(rag.values.reshape(-1, 3))
            splits = Tensor.from_numpy(rag.row_splits, dtype=np.int64)
        else:
            coords = as_tensor(coords).contig()
            B, N = coords.shape[0], coords.shape[1]
            cvals, splits = coords, None
        ref = as_tensor(ref).contig()
        if ref.size != B * 3:
            raise ValueError('ref must have shape (N_batch, 3) or (N_batch, 1, 3); got %s for batch %d' % (ref.shape, B))
        box_ptr, per_row = None, 0
        if box_lengths is not None:  # per-call box takes precedence over the stored one (mappings.py:408-412)
            box = as_tensor(box_lengths).contig()
            if box.size != B * 3:
                raise ValueError('box_lengths must have shape (N_batch, 3); got %s' % (box.shape, ))
            box_ptr, per_row = box.ptr, 1
        elif self.box_lengths is not None:
            box_ptr = self._box_dev.ptr
        info_ptr, P_ = None, 0
        out_info = None
        k = int(self.max_included)
        if particle_info is not None:
            irag = self._ragged(particle_info)
            if irag is not None:
                if rag is None or not np.array_equal(irag.row_splits, rag.row_splits):
                    raise ValueError('particle_info must be ragged like coords')
                info = Tensor.from_numpy(irag.values)
            else:
                info = as_tensor(particle_info).contig()
                if rag is not None or info.shape[:2] != (B, N):
                    raise ValueError('particle_info must have shape (N_batch, N_particles, P) like coords')
            P_ = info.shape[-1]
            info_ptr = info.ptr
            out_info = Tensor((B, k, P_))
        out_xyz = Tensor((B, k, 3))
        out_idx = Tensor((B, k), np.int32) if return_indices else None
        c.lib.vms_dist_select(cvals.ptr, None if splits is None else splits.ptr, B, N, ref.ptr, box_ptr, per_row,
                              float(np.float32(self.sq_cut)), k, info_ptr, P_, out_xyz.ptr,
                              None if out_info is None else out_info.ptr, None if out_idx is None else out_idx.ptr,
                              c.stream)
        outs = [out_xyz] + ([out_info] if out_info is not None else []) + ([out_idx] if return_indices else [])
        return outs[0] if len(outs) == 1 else tuple(outs)

    def get_config(self):
        config = super(DistanceSelection, self).get_config()
        config.update({"cutoff": self.cutoff, "max_included": self.max_included, "box_lengths": self.box_lengths})
        return config
