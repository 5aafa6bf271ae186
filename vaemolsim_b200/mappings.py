"""Mappings -- host-side mirror of the hot-path part of `vaemolsim/mappings.py` over sm_100a kernels.

In scope: `FCDeepNN` (mappings.py:18-166) and `DistanceSelection` (mappings.py:308-477), same
constructor keywords, call signatures and error behaviour; SURVEY 8f rank 2: `AttentionBlock` (mappings.py:480-561),
`ParticleEmbedding` (:564-688) and `LocalParticleDescriptors` (:691-762) over `VectorAttention`, a mirror of
`geometric_algebra_attention.keras.VectorAttention` for the one configuration the reference builds (rank 2, concat merge
and join; the package is un-vendored and unpinned -- arithmetic restated in oracle/gaa.py, parity unpinned).  The CG
template layers are out of scope.
"""
import ctypes as C
import os

import numpy as np

from . import _protocols as P
from . import _abi
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


# ============================================================================ geometric-algebra attention
def zero_mask(coords):
    """tf.keras.layers.Masking(mask_value=0.0).compute_mask on [B, n, 3] coordinates: uint8 [B, n], 1 = keep."""
    c = ctx()
    coords = as_tensor(coords).contig()
    B, n = coords.shape[0], coords.shape[1]
    mask = Tensor((B, n), np.uint8)
    c.lib.vms_gaa_zero_mask(coords.ptr, B * n, mask.ptr, c.stream)
    return mask


class VectorAttention(P.Layer):
    """`geometric_algebra_attention.keras.VectorAttention(score_net, value_net, reduce, merge_fun, join_fun, rank)` for the
    configuration of mappings.py:518-525 / :633-647: rank 2, `merge_fun='concat'`, `join_fun='concat'`.

    call([coords [B, n, 3], values [B, n, D]], mask=uint8 [B, n] or None) -> [B, n, D] (reduce=False) or [B, D].
    Training (a tape is active): pair tensors in HBM, every step an op-by-op kernel with a reverse mode.  Otherwise ONE
    kernel per layer (`vms_gaa_attention_forward`) when the networks have the reference's structure and widths fit
    (D <= 32, hidden <= 64; `VMS_GAA_FUSED=0` keeps the op-by-op path)."""

    def __init__(self, score_net, value_net, reduce=True, merge_fun='mean', join_fun='mean', rank=2, name='vector_attention',
                 **kwargs):
        super(VectorAttention, self).__init__(name=name, **kwargs)
        if rank != 2 or merge_fun != 'concat' or join_fun != 'concat':
            raise NotImplementedError('VectorAttention: only rank=2, merge_fun="concat", join_fun="concat" (the configuration '
                                      'of vaemolsim/mappings.py:518-525) is built')
        self.score_net, self.value_net = score_net, value_net
        self.reduce, self.merge_fun, self.join_fun, self.rank = reduce, merge_fun, join_fun, rank

    def build(self, input_shape):
        D = int(input_shape[1][-1])
        sd = np.sqrt(2.0 / self.rank / D)  # geometric_algebra_attention: normal(stddev sqrt(2 / rank / n_dim)) projections
        r = P.rng()
        self.merge_kernels = [Tensor.from_numpy(r.normal(0, sd, (D, D)).astype(np.float32)) for _ in range(2)]
        self.join_kernels = [Tensor.from_numpy(r.normal(0, sd, (D, D)).astype(np.float32)) for _ in range(2)]
        self._weights = self.merge_kernels + self.join_kernels
        for net, width in ((self.score_net, D), (self.value_net, 2)):
            if not net.built:
                net.build((None, width))
                net.built = True

    def _fused_weights(self, D):
        """The layer's weights as vms_gaa_weights when score_net = Dense(H, act) Dense(1) and value_net = Dense(H)
        LayerNormalization Activation(act) Dense(D) (mappings.py:505-514), else None."""
        s, v = getattr(self.score_net, 'layers', None), getattr(self.value_net, 'layers', None)
        if not (s and v and len(s) == 2 and len(v) == 4 and isinstance(s[0], P.Dense) and isinstance(s[1], P.Dense) and
                isinstance(v[0], P.Dense) and isinstance(v[1], P.LayerNormalization) and isinstance(v[2], P.Activation) and
                isinstance(v[3], P.Dense)):
            return None
        H = s[0].units
        if not (s[1].units == 1 and s[1].act == 0 and v[0].units == H and v[0].act == 0 and v[3].units == D and
                v[3].act == 0 and s[0].act == v[2].act and all(l.use_bias for l in (s[0], s[1], v[0], v[3]))):
            return None
        w = _abi.GaaWeights(self.merge_kernels[0].ptr, self.merge_kernels[1].ptr, self.join_kernels[0].ptr,
                            self.join_kernels[1].ptr, s[0].kernel.ptr, s[0].bias.ptr, s[1].kernel.ptr, s[1].bias.ptr,
                            v[0].kernel.ptr, v[0].bias.ptr, v[1].gamma.ptr, v[1].beta.ptr, v[3].kernel.ptr, v[3].bias.ptr)
        return w, H, s[0].act, v[1].epsilon

    def call(self, inputs, mask=None):
        c = ctx()
        coords, values = as_tensor(inputs[0]).contig(), as_tensor(inputs[1]).contig()
        if isinstance(mask, (list, tuple)):  # Keras hands a list input its list of masks: the coordinate mask is the first
            mask = mask[0]
        B, n, D = values.shape
        if coords.shape != (B, n, 3):
            raise ValueError('VectorAttention: coords %s do not match values %s' % (coords.shape, values.shape))
        from . import _autodiff
        tp = _autodiff.Tape.active()
        reduce = 1 if self.reduce else 0
        out = Tensor((B, D)) if reduce else Tensor((B, n, D))
        if tp is None and os.environ.get('VMS_GAA_FUSED', '1') != '0':
            fw = self._fused_weights(D)
            if fw is not None and c.lib.vms_gaa_attention_forward_supported(n, D, fw[1]):
                w, H, act, eps = fw
                c.lib.vms_gaa_attention_forward(coords.ptr, values.ptr, D, None if mask is None else mask.ptr, B, n, D, H,
                                                C.byref(w), reduce, act, eps, out.ptr, c.stream)
                return out
        # op-by-op: pair tensors [B n n, .] in HBM
        v2 = values.reshape(B * n, D)
        inv = Tensor((B * n * n, 2))
        c.lib.vms_gaa_pair_invariants(coords.ptr, B, n, inv.ptr, c.stream)
        iv = self.value_net.call(inv)
        u = P.dense_op(v2, self.merge_kernels[0])
        w_ = P.dense_op(v2, self.merge_kernels[1])
        merged = Tensor((B * n * n, D))
        c.lib.vms_gaa_pair_merge(u.ptr, u.ld, w_.ptr, w_.ld, B, n, D, merged.ptr, c.stream)
        if tp is not None:

            def bw_merge():
                if tp.has(merged):
                    gu, gw = tp.grad(u), tp.grad(w_)
                    c.lib.vms_gaa_pair_merge_backward(tp.grad(merged).ptr, B, n, D, gu.ptr, gu.ld, gw.ptr, gw.ld, c.stream)

            tp.record(bw_merge)
        joined = P.dense_op(iv, self.join_kernels[0], cond=merged, cond_kernel=self.join_kernels[1])
        scores = self.score_net.call(joined)
        att = Tensor((B * n * n, ))
        c.lib.vms_gaa_attend(scores.ptr, joined.ptr, None if mask is None else mask.ptr, B, n, D, reduce, out.ptr, att.ptr,
                             c.stream)
        if tp is not None:

            def bw_attend():  # (holds `mask` and `att` until the tape is released)
                if tp.has(out):
                    c.lib.vms_gaa_attend_backward(att.ptr, joined.ptr, None if mask is None else mask.ptr, B, n, D, reduce,
                                                  tp.grad(out).ptr, tp.grad(scores).ptr, tp.grad(joined).ptr, c.stream)

            tp.record(bw_attend)
        return out


def _mlp_ln(hidden_dim, out_dim, activation):
    """Dense(hidden) -> LayerNormalization -> Activation -> Dense(out)   (mappings.py:509-514, :526-531, :638-643)."""
    return P.Sequential([P.Dense(hidden_dim), P.LayerNormalization(), P.Activation(activation), P.Dense(out_dim)])


class AttentionBlock(P.Layer):
    """mappings.py:480-561: geometric-algebra attention (reduce=False) -> Dense / LayerNorm / activation / Dense ->
    residual.  Rotation invariant, permutation equivariant."""

    def __init__(self, hidden_dim=40, name='geom_attn', activation='relu', **kwargs):
        super(AttentionBlock, self).__init__(name=name, **kwargs)
        self.hidden_dim = hidden_dim
        self.activation = activation
        self.supports_masking = True

    def build(self, input_shape):
        working_dim = int(input_shape[1][-1])
        self.score_fun = P.Sequential([P.Dense(self.hidden_dim, activation=self.activation), P.Dense(1)])
        self.value_fun = _mlp_ln(self.hidden_dim, working_dim, self.activation)
        self.attn = VectorAttention(self.score_fun, self.value_fun, reduce=False, merge_fun='concat', join_fun='concat',
                                    rank=2)
        self.nonlinearity = _mlp_ln(self.hidden_dim, working_dim, self.activation)
        self.attn.build(input_shape)
        self.attn.built = True
        self.nonlinearity.build((None, working_dim))
        self.nonlinearity.built = True

    def _sublayers(self):
        return [self.attn, self.nonlinearity]  # (score_fun / value_fun are reached through attn)

    def call(self, inputs, mask=None):
        coords, embedding = as_tensor(inputs[0]), as_tensor(inputs[1]).contig()
        B, n, D = embedding.shape
        new_embed = self.attn.call([coords, embedding], mask=mask)
        new_embed = self.nonlinearity.call(new_embed.reshape(B * n, D)).reshape(B, n, D)
        return new_embed + embedding

    def get_config(self):
        config = super(AttentionBlock, self).get_config()
        config.update({"hidden_dim": self.hidden_dim})
        return config


class ParticleEmbedding(P.Layer):
    """mappings.py:564-688: info_net -> num_blocks AttentionBlocks -> a final permutation-invariant attention
    (reduce=True); with `mask_zero` particles whose coordinates are all zero (DistanceSelection's padding and the site
    itself) are masked out of every attention."""

    def __init__(self, embedding_dim, hidden_dim=40, num_blocks=2, mask_zero=True, name='particle_embedding',
                 activation='relu', **kwargs):
        super(ParticleEmbedding, self).__init__(name=name, **kwargs)
        self.embedding_dim = embedding_dim
        self.hidden_dim = hidden_dim
        self.num_blocks = num_blocks
        self.mask_zero = mask_zero
        self.activation = activation

    def build(self, input_shape):
        self.info_net = P.Dense(self.embedding_dim)  # no activation: linear map to the working dimension (mappings.py:621)
        self.block_list = [AttentionBlock(self.hidden_dim, activation=self.activation) for _ in range(self.num_blocks)]
        self.mask = zero_mask if self.mask_zero else None
        self.final_attn = VectorAttention(
            P.Sequential([P.Dense(self.hidden_dim, activation=self.activation), P.Dense(1)]),
            _mlp_ln(self.hidden_dim, self.embedding_dim, self.activation), reduce=True, merge_fun='concat',
            join_fun='concat', rank=2)

    def call(self, coords, particle_info):
        coords, info = as_tensor(coords).contig(), as_tensor(particle_info).contig()
        B, n, P_ = info.shape
        mask = self.mask(coords) if self.mask_zero else None
        if not self.info_net.built:
            self.info_net.build((None, P_))
            self.info_net.built = True
        E = self.embedding_dim
        embedding = self.info_net.call(info.reshape(B * n, P_)).reshape(B, n, E)
        for block in self.block_list:
            embedding = block([coords, embedding], mask=mask)
        out = self.final_attn([coords, embedding], mask=mask)
        if mask is not None:
            out._keras_mask = mask
        return out

    def get_config(self):
        config = super(ParticleEmbedding, self).get_config()
        config.update({"embedding_dim": self.embedding_dim, "hidden_dim": self.hidden_dim, "num_blocks": self.num_blocks,
                       "mask_zero": self.mask_zero})
        return config


class LocalParticleDescriptors(P.Layer):
    """mappings.py:691-762: distance masking around a reference site, then an embedding of the local point cloud."""

    def __init__(self, mask_fn, embed_fn, name='local_particle_desc', **kwargs):
        super(LocalParticleDescriptors, self).__init__(name=name, **kwargs)
        self.mask_fn = mask_fn
        self.embed_fn = embed_fn

    def call(self, coords, ref, props, box_lengths=None):
        local_coords, local_props = self.mask_fn(coords, ref, particle_info=props, box_lengths=box_lengths)
        return self.embed_fn(local_coords, local_props)

    def get_config(self):
        config = super(LocalParticleDescriptors, self).get_config()
        config.update({"mask_fn": self.mask_fn, "embed_fn": self.embed_fn})
        return config
