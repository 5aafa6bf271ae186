"""Models -- host-side mirror of `vaemolsim/models.py` over sm_100a kernels.

Same names / keywords as the reference (models.py:16 `FlowModel`, :153 `MappingToDistribution`, :242 `VAE`).  Calling a
model builds distribution objects exactly as the reference does (op-by-op kernels); TRAINING (Keras `fit` /
GradientTape / Adam in the reference, tests/test_models.py:181-182) goes through the fused ELBO plan of `csrc/elbo.cu`
for the model family of the reference's tests (IndependentNormal encoder / decoder over one-hidden-layer FCDeepNNs,
N(0, I) or RealNVP-RQS-flowed prior, KLDivergenceEstimate + LogProbLoss) and raises NotImplementedError for
compositions whose backward kernels are not built yet.  `BackmappingOnly` (models.py:470-572) runs and trains
on the tape path.  `VAEDualELBO` (models.py:335, cannot be constructed in the reference) is not mirrored.
"""
import ctypes as C

import numpy as np

from . import _abi, dists, flows, losses, mappings
from . import _protocols as P
from ._abi import Tensor, as_tensor, ctx


class Adam(object):
    """tf.keras.optimizers.Adam hyper-parameters (Keras defaults; tests/test_models.py:181 uses lr 1e-3)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon


class Model(P.Layer):
    """The slice of tf.keras.Model the reference's tests use: compile / fit / evaluate / predict."""

    def compile(self, optimizer=None, loss=None, **kwargs):
        self.optimizer = optimizer if optimizer is not None else Adam()
        self.loss = loss

    def predict_step(self, inputs):
        return self(inputs, training=False).sample()

    # ------------------------------------------------------------------ generic training (tape-based reverse mode)
    # Keras `fit` / `evaluate` for ANY composition of this package's layers (tests/test_models.py:189-228, models.py:85-139):
    # forward op by op under a tape (_autodiff.py), reverse-mode kernels, Adam per weight tensor.  The fused plans of
    # `VAE` take over for the model family they support.
    def _loss_tensor(self, xb, yb, training):
        out = self(xb, training=training)
        loss_fn = getattr(self, 'loss', None) or losses.LogProbLoss()
        total = loss_fn(yb, out)
        for extra in getattr(self, 'losses', []) or []:
            if isinstance(extra, Tensor):
                total = total + extra
            elif extra:
                total = total + float(extra)
        return total

    def _check_trainable(self):
        pass  # every layer of this package has a reverse mode now (batch normalisation included)

    def _trainer(self):
        from . import _autodiff
        opt = getattr(self, 'optimizer', None) or Adam()
        tr = getattr(self, '_generic_trainer', None)
        grp = getattr(self, '_dp_group', None)
        if tr is None or tr.opt is not opt or tr.group is not (grp if (grp is not None and grp.world > 1) else None):
            tr = self._generic_trainer = _autodiff.Trainer(self, opt, group=grp)
        return tr

    def distribute(self, group):
        """Data-parallel training over `group` (`parallel.Group`, one process per GPU): `fit` / `train_on_batch` run on this
        rank's shard of the data and every step sums the gradients of all weights over the ranks (one NCCL allreduce) before
        Adam, so the replicas stay identical.  Batch-normalisation layers use cross-replica batch statistics (the moments of
        the whole batch: two small allreduces per layer forward, one in the reverse mode), i.e. single-device semantics.  Pass None to go back to single-process training."""
        self._dp_group = group
        return self

    def _prep_xy(self, x, y):
        x = np.asarray(x.numpy() if isinstance(x, Tensor) else x, np.float32)
        y = x if y is None else np.asarray(y.numpy() if isinstance(y, Tensor) else y, np.float32)
        return x, y

    def fit(self, x, y=None, epochs=1, batch_size=32, verbose=0, shuffle=True):
        x, y = self._prep_xy(x, y)
        self(Tensor.from_numpy(x[:min(2, len(x))]))  # build
        self._check_trainable()
        tr = self._trainer()
        hist = {'loss': []}
        for _ in range(epochs):
            order = P.rng().permutation(len(x)) if shuffle else np.arange(len(x))
            tot, n = 0.0, 0
            for i in range(0, len(x), batch_size):
                idx = order[i:i + batch_size]
                loss = tr.step_arrays(np.ascontiguousarray(x[idx]), np.ascontiguousarray(y[idx]),
                                      lambda xb, yb: self._loss_tensor(xb, yb, True))
                tot, n = tot + float(loss.numpy()) * len(idx), n + len(idx)
            hist['loss'].append(tot / max(n, 1))
        return hist

    def train_on_batch(self, x, y=None):
        """One optimiser step on one batch; returns the loss (the inner call of `fit`)."""
        x, y = self._prep_xy(x, y)
        self._check_trainable()
        return float(self._trainer().step_arrays(np.ascontiguousarray(x), np.ascontiguousarray(y),
                                                 lambda xb, yb: self._loss_tensor(xb, yb, True)).numpy())

    def evaluate(self, x, y=None, batch_size=32, verbose=0):
        x, y = self._prep_xy(x, y)
        tot, n = 0.0, 0
        for i in range(0, len(x), batch_size):
            xb, yb = Tensor.from_numpy(x[i:i + batch_size]), Tensor.from_numpy(y[i:i + batch_size])
            tot, n = tot + float(self._loss_tensor(xb, yb, False).numpy()) * xb.shape[0], n + xb.shape[0]
        return tot / max(n, 1)

    # ------------------------------------------------------------------ weights in / out (Keras `get_weights` protocol)
    # The reference's users move trained weights with Keras (`model.save_weights`, `model.get_weights()`, models.py:470-572;
    # the `order_seed` of RQSSplineMAF exists so that a re-built flow matches saved weights, flows.py:572-575).  TF / h5py
    # are not available here, so the exchange format is the Keras VARIABLE ORDER itself: `get_weights()` / `set_weights()`
    # list every variable in the order Keras tracks them for the same composition (layer creation order; per Dense kernel
    # then bias; per MADE layer kernel, bias[, conditional kernel]; batch norm gamma, beta, moving mean, moving variance),
    # and `save_weights` / `load_weights` store that list as an .npz -- on the TF side
    # `np.savez(path, *keras_model.get_weights())` produces a file `load_weights` reads, and
    # `keras_model.set_weights([f[k] for k in f.files])` consumes one written here.  MADE kernels are re-masked on import.
    def _unique_weights(self):
        out = []
        for w in self.weights:
            if not any(w is u for u in out):
                out.append(w)
        return out

    def get_weights(self):
        return [w.numpy() for w in self._unique_weights()]

    def set_weights(self, arrays):
        ws = self._unique_weights()
        arrays = list(arrays)
        if len(arrays) != len(ws):
            raise ValueError('set_weights: the model has %d variables, got %d arrays (build the model by calling it once on '
                             'data first)' % (len(ws), len(arrays)))
        c = ctx()
        for w, a in zip(ws, arrays):
            a = np.ascontiguousarray(a, np.float32)
            if tuple(a.shape) != tuple(w.shape):
                raise ValueError('set_weights: shape %s does not match variable of shape %s' % (a.shape, w.shape))
            mask = getattr(w, '_grad_mask', None)
            if mask is not None:
                a = np.ascontiguousarray(a * mask.numpy())
            if w.contiguous:
                c.lib.vms_memcpy_h2d(w.ptr, a.ctypes.data, a.nbytes, c.stream)
                c.synchronize()
            else:
                w.assign_cols(0, Tensor.from_numpy(a))
        _abi.bump_param_epoch()
        f = getattr(self, '_fused', None)
        if f is not None:
            f.invalidate()

    def save_weights(self, path):
        arrs = self.get_weights()
        np.savez(path, **{'arr_%d' % i: a for i, a in enumerate(arrs)})

    def load_weights(self, path):
        with np.load(path) as f:
            keys = sorted(f.files, key=lambda k: int(k.split('_')[-1]) if k.split('_')[-1].isdigit() else 0)
            self.set_weights([f[k] for k in keys])

    def predict(self, x, batch_size=32, verbose=0):
        x = np.asarray(x.numpy() if isinstance(x, Tensor) else x, np.float32)
        outs = [self.predict_step(Tensor.from_numpy(x[i:i + batch_size])).numpy() for i in range(0, len(x), batch_size)]
        return np.concatenate(outs, axis=0)


class FlowModel(Model):
    """models.py:16-150: (optional mapping) -> FlowedDistribution."""

    def __init__(self, flow, latent_dist, mapping=None, name='flow_model', **kwargs):
        super(FlowModel, self).__init__(name=name, **kwargs)
        self.flowed_dist = dists.FlowedDistribution(flow, latent_dist)
        static = isinstance(latent_dist, P.DistributionLambda)  # models.py:74-83
        if mapping is None:
            self.mapping = None if static else mappings.FCDeepNN(self.flowed_dist.params_size())
        elif static:
            print("Warning: for static distribution (DistributionLambda as latent distribution), cannot have "
                  "mapping, so setting to None.")
            self.mapping = None
        else:
            self.mapping = mapping  # the reference forgets this assignment

    def call(self, inputs, training=False):
        mapped = self.mapping(inputs, training=training) if self.mapping is not None else inputs
        if self.flowed_dist.conditional:
            return self.flowed_dist(mapped, conditional_input=inputs, training=training)
        return self.flowed_dist(mapped, training=training)

    def get_config(self):
        config = super(FlowModel, self).get_config()
        config.update({"flow": self.flowed_dist.flow, "latent_dist": self.flowed_dist.latent_dist,
                       "mapping": self.mapping})
        return config


class MappingToDistribution(Model):
    """models.py:153-239: mapping network (FCDeepNN by default) followed by a distribution layer."""

    def __init__(self, distribution, mapping=None, name='map_to_dist', **kwargs):
        super(MappingToDistribution, self).__init__(name=name, **kwargs)
        self.distribution = distribution
        self.conditional = getattr(self.distribution, 'conditional', False)
        if mapping is None:
            if isinstance(self.distribution, P.DistributionLambda):
                self.mapping = mappings.FCDeepNN(self.distribution.params_size(self.distribution.event_size))
            else:
                self.mapping = mappings.FCDeepNN(self.distribution.params_size())
        else:
            self.mapping = mapping

    def call(self, inputs, training=False):
        mapped = self.mapping(inputs, training=training)
        if self.conditional:
            return self.distribution(mapped, training=training, conditional_input=inputs)
        return self.distribution(mapped, training=training)

    def get_config(self):
        config = super(MappingToDistribution, self).get_config()
        config.update({"distribution": self.distribution, "mapping": self.mapping})
        return config


class BackmappingOnly(Model):
    """models.py:470-572: [CG sites to decode (B, 1, 3), all other coordinates (B, (N), 3) -- dense or ragged --, particle
    properties (B, (N), P)] -> local descriptors (`mask_and_embed`, e.g. `LocalParticleDescriptors`) -> decoding
    distribution (`decode_dist`, e.g. `MappingToDistribution`).  `fit` / `evaluate` / `predict` take that list as `x`
    (tests/test_models.py:265-308) and train on the tape path."""

    def __init__(self, mask_and_embed, decode_dist, name='backmapping', **kwargs):
        super(BackmappingOnly, self).__init__(name=name, **kwargs)
        self.mask_and_embed = mask_and_embed
        self.decode_dist = decode_dist

    def call(self, inputs, training=False):
        cg_to_decode, other_coords, other_particle_props = inputs[0], inputs[1], inputs[2]
        local_descriptors = self.mask_and_embed(other_coords, cg_to_decode, other_particle_props)
        return self.decode_dist(local_descriptors, training=training)

    # -- Keras fit / evaluate / predict over a LIST of inputs whose second and third entries may be ragged
    @staticmethod
    def _n_rows(inputs):
        return int(np.asarray(inputs[0].numpy() if isinstance(inputs[0], Tensor) else inputs[0]).shape[0])

    @staticmethod
    def _take(a, idx):
        if isinstance(a, mappings.RaggedTensor):
            return mappings.RaggedTensor.from_rows([a.values[a.row_splits[i]:a.row_splits[i + 1]] for i in idx],
                                                   inner=a.values.shape[-1])
        if isinstance(a, (list, tuple)):
            return [a[i] for i in idx]
        a = a.numpy() if isinstance(a, Tensor) else np.asarray(a, np.float32)
        return np.ascontiguousarray(a[idx])

    def _batch(self, inputs, idx):
        return [self._take(a, idx) for a in inputs]

    def fit(self, x, y=None, epochs=1, batch_size=32, verbose=0, shuffle=True):
        n = self._n_rows(x)
        y = np.asarray(y.numpy() if isinstance(y, Tensor) else y, np.float32)
        self(self._batch(x, np.arange(min(2, n))))  # build
        tr = self._trainer()
        hist = {'loss': []}
        for _ in range(epochs):
            order = P.rng().permutation(n) if shuffle else np.arange(n)
            tot, cnt = 0.0, 0
            for i in range(0, n, batch_size):
                idx = order[i:i + batch_size]
                xb, yb = self._batch(x, idx), Tensor.from_numpy(y[idx])
                loss = tr.step(lambda: self._loss_tensor(xb, yb, True))
                tot, cnt = tot + float(loss.numpy()) * len(idx), cnt + len(idx)
            hist['loss'].append(tot / max(cnt, 1))
        return hist

    def train_on_batch(self, x, y=None):
        yb = Tensor.from_numpy(np.asarray(y.numpy() if isinstance(y, Tensor) else y, np.float32))
        return float(self._trainer().step(lambda: self._loss_tensor(x, yb, True)).numpy())

    def evaluate(self, x, y=None, batch_size=32, verbose=0):
        n = self._n_rows(x)
        y = np.asarray(y.numpy() if isinstance(y, Tensor) else y, np.float32)
        tot = 0.0
        for i in range(0, n, batch_size):
            idx = np.arange(i, min(i + batch_size, n))
            tot += float(self._loss_tensor(self._batch(x, idx), Tensor.from_numpy(y[idx]), False).numpy()) * len(idx)
        return tot / max(n, 1)

    def predict(self, x, batch_size=32, verbose=0):
        n = self._n_rows(x)
        outs = [self.predict_step(self._batch(x, np.arange(i, min(i + batch_size, n)))).numpy()
                for i in range(0, n, batch_size)]
        return np.concatenate(outs, axis=0)

    def get_config(self):
        config = super(BackmappingOnly, self).get_config()
        config.update({"mask_and_embed": self.mask_and_embed, "decode_dist": self.decode_dist})
        return config


class VAE(Model):
    """models.py:242-332: encoder -> sample -> prior -> regulariser -> decoder."""

    def __init__(self, encoder, decoder, prior, regularizer=None, name='vae', **kwargs):
        super(VAE, self).__init__(name=name, **kwargs)
        self.encoder = encoder
        self.decoder = decoder
        self.prior = prior
        self.regularizer = regularizer if regularizer is not None else losses.KLDivergenceEstimate()
        self.losses = []
        self.metrics = {}
        self._fused = None

    def call(self, inputs, training=False):
        inputs = as_tensor(inputs)
        encode_dist = self.encoder(inputs, training=training)
        encode_sample = encode_dist.sample()
        prior_dist = self.prior(encode_sample, training=training)
        reg_loss = self.regularizer(encode_dist, prior_dist, encode_sample)
        self.losses = [reg_loss]
        self.metrics = {'kl_div': reg_loss / float(self.regularizer.weight), 'regularizer_loss': reg_loss}
        return self.decoder(encode_sample, training=training)

    # ------------------------------------------------------------------ fused ELBO path (csrc/elbo.cu)
    def fused(self, max_batch=4096):
        """The fused ELBO plan for this model (built on first use; rebuilt if max_batch grows)."""
        if self._fused is None or self._fused.max_batch < max_batch:
            old = self._fused
            self._fused = FusedELBO(self, max_batch)
            if old is not None:
                # re-planning for a larger batch must not reset the optimiser: carry Adam's moments and step count over
                # (the weights travel through the layers' own tensors, re-bound by the new plan)
                c = ctx()
                c.lib.vms_memcpy_d2d(self._fused.m.ptr, old.m.ptr, old.m.nbytes, c.stream)
                c.lib.vms_memcpy_d2d(self._fused.v.ptr, old.v.ptr, old.v.nbytes, c.stream)
                c.synchronize()
                self._fused.t = old.t
                old.close()
        return self._fused

    def train_step(self, x, eps=None):
        """One Adam step on a batch; returns dict(loss, nll, kl) as floats."""
        x = as_tensor(x)
        f = self.fused(x.shape[0])
        if eps is None:
            eps = P.rng().standard_normal((x.shape[0], f.dz), dtype=np.float32)
        scal = f.train_step(x, as_tensor(eps), getattr(self, 'optimizer', None) or Adam())
        return dict(zip(('loss', 'nll', 'kl'), scal.numpy()[:3].tolist()))

    def _fused_or_none(self, max_batch):
        """The fused plan when this composition belongs to its family (and the loss is LogProbLoss), else None: the
        tape-based generic path of `Model` trains everything else."""
        if getattr(self, 'loss', None) is not None and not isinstance(self.loss, losses.LogProbLoss):
            return None
        try:
            return self.fused(max_batch)
        except NotImplementedError:
            return None

    def fit(self, x, y=None, epochs=1, batch_size=32, verbose=0, shuffle=True):
        xa = np.asarray(x.numpy() if isinstance(x, Tensor) else x, np.float32)
        if not self.built:
            self(Tensor.from_numpy(xa[:min(2, len(xa))]))
        f = self._fused_or_none(min(int(batch_size), len(xa)))
        if f is None:
            return Model.fit(self, x, y, epochs=epochs, batch_size=batch_size, verbose=verbose, shuffle=shuffle)
        x = xa.reshape(len(xa), -1)
        hist = {'loss': [], 'kl_div': []}
        opt = getattr(self, 'optimizer', None) or Adam()
        for _ in range(epochs):
            order = P.rng().permutation(len(x)) if shuffle else None
            scal = f.train_loop(x, opt, batch_size, order=order)  # pipelined: copies, noise and read-backs overlap the steps
            w = np.minimum(batch_size, len(x) - np.arange(0, len(x), min(int(batch_size), len(x)))).astype(np.float64)
            hist['loss'].append(float(np.sum(scal[:, 0] * w) / w.sum()))
            hist['kl_div'].append(float(np.sum(scal[:, 2] * w) / w.sum()))
            if f.tc_status():  # a tensor-core completion wait ran into its bound: this epoch's gradients are invalid
                raise RuntimeError('VAE.fit: the tensor-core plan reported an MMA completion time-out; parameters of '
                                   'this epoch are invalid')
        return hist

    def evaluate(self, x, y=None, batch_size=32, verbose=0):
        x = np.asarray(x.numpy() if isinstance(x, Tensor) else x, np.float32)
        f = self._fused_or_none(min(batch_size, len(x)))
        if f is None:
            return Model.evaluate(self, x, y, batch_size=batch_size, verbose=verbose)
        tot, n = 0.0, 0
        for i in range(0, len(x), batch_size):
            xb = Tensor.from_numpy(x[i:i + batch_size])
            eps = Tensor.from_numpy(P.rng().standard_normal((xb.shape[0], f.dz), dtype=np.float32))
            tot, n = tot + float(f.forward(xb, eps)['scalars'].numpy()[0]) * xb.shape[0], n + xb.shape[0]
        return tot / n

    def get_config(self):
        config = super(VAE, self).get_config()
        config.update({"encoder": self.encoder, "decoder": self.decoder, "prior": self.prior,
                       "regularizer": self.regularizer})
        return config


class FusedELBO(object):
    """Binds a VAE of the supported family to a `vms_elbo_plan`: ONE flat parameter buffer (the layers' kernels / biases
    become views into it), whole-step CUDA graph, flat gradient, Adam state.

    Supported: encoder / decoder = MappingToDistribution(IndependentNormal, FCDeepNN with one relu hidden layer, no
    periodic dofs); prior = DistributionLambda -> StandardNormal, or FlowedDistribution(RQSSplineRealNVP without batch
    norm / before / after transforms, StandardNormal latent); regulariser KLDivergenceEstimate sampled from dist_a.
    """

    def __init__(self, vae, max_batch):
        self.vae = vae
        self.max_batch = int(max_batch)
        enc_layers, self.dx, self.dz, hidden = self._check_mlp(vae.encoder, 'encoder')
        dec_layers, dz2, dx2, hidden2 = self._check_mlp(vae.decoder, 'decoder')
        if (dz2, dx2, hidden2) != (self.dz, self.dx, hidden):
            raise NotImplementedError('fused ELBO: encoder / decoder shapes must mirror each other')
        reg = vae.regularizer
        if type(reg) is not losses.KLDivergenceEstimate or reg.sample_dist != 'dist_a':
            raise NotImplementedError('fused ELBO: only KLDivergenceEstimate (samples from dist_a) is built')
        blocks, nbins, fh, lo, hi = self._check_prior(vae.prior, self.dz)
        self.desc = _abi.ElboDesc(self.dx, self.dz, hidden, len(blocks), nbins, fh, lo, hi, float(reg.weight),
                                  self.max_batch)
        c = ctx()
        self.n_params = int(c.lib.vms_elbo_param_count(C.byref(self.desc)))
        self.theta = Tensor((self.n_params, ))
        self.grad = Tensor((self.n_params, ))
        self.m = Tensor.zeros((self.n_params, ))
        self.v = Tensor.zeros((self.n_params, ))
        self.scalars = Tensor((4, ))
        self.t = 0
        # flat order: enc.0 enc.1 dec.0 dec.1, then per flow block d1, heads (include/vms_b200.h)
        off = 0
        for lay in enc_layers + dec_layers + [l for b in blocks for l in (b.d1, b.heads)]:
            W, b = lay.get_weights()
            kW = Tensor(W.shape, _ptr=self.theta.ptr + 4 * off, _base=self.theta)
            off += W.size
            kb = Tensor(b.shape, _ptr=self.theta.ptr + 4 * off, _base=self.theta)
            off += b.size
            lay.rebind(kW, kb)
            lay.assign(W, b)
            lay._on_assign = self.invalidate  # host-side weight assignment: the plan's packed weight images are stale
        if off != self.n_params:
            raise RuntimeError('fused ELBO: parameter layout mismatch (%d != %d)' % (off, self.n_params))
        h = C.c_void_p()
        c.lib.vms_elbo_plan_create(C.byref(self.desc), C.byref(h))
        self.handle = h.value

    def _release_loop(self):
        """Frees the pipelined loop's copy stream, events and pinned buffers (`train_loop`)."""
        L, self._loop = getattr(self, '_loop', None), None
        if L is None:
            return
        lib = _abi.load()
        lib.vms_stream_synchronize(L['copy_stream'])
        for e in L['h2d'] + L['free']:
            lib.vms_event_destroy(e)
        for p in L['stage'] + [L['host_ring']]:
            lib.vms_free_host(p)
        lib.vms_stream_destroy(L['copy_stream'])

    def close(self):
        """Releases the plan and the loop resources; the parameter / moment tensors stay valid (layers view theta)."""
        try:
            self._release_loop()
            if getattr(self, 'handle', None):
                _abi.load().vms_elbo_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def __del__(self):
        self.close()

    @staticmethod
    def _check_mlp(m, what):
        ok = (isinstance(m, MappingToDistribution) and isinstance(m.distribution, P.IndependentNormal)
              and isinstance(m.mapping, mappings.FCDeepNN))
        if not ok:
            raise NotImplementedError('fused ELBO: %s must be MappingToDistribution(IndependentNormal, FCDeepNN)' % what)
        fc = m.mapping
        if not fc.built:
            raise RuntimeError('fused ELBO: call the model once on data before training so its layers are built')
        dense = [l for l in fc.layer_list if isinstance(l, P.Dense)]
        if fc.any_periodic or fc.batch_norm or len(dense) != 2 or dense[0].act != P.ACT['relu']:
            raise NotImplementedError('fused ELBO: %s FCDeepNN must have one relu hidden layer, no periodic dofs, no '
                                      'batch normalisation' % what)
        l0, l1 = dense
        return [l0, l1], l0.kernel.shape[0], l1.units // 2, l0.units

    @staticmethod
    def _is_standard_normal(latent, dz):
        """True only if `latent` (a DistributionLambda) really produces N(0, I_dz): it is called once on a dummy latent
        batch, exactly as the reference calls the prior at models.py:306-308 / mcmc.py:101,107.  An IndependentNormal
        layer (a DistributionLambda subclass) or a lambda returning a non-unit Normal must NOT be taken for N(0, I)."""
        if type(latent) is not P.DistributionLambda:
            return False
        try:
            d = latent(Tensor.zeros((2, dz)))
        except Exception:
            return False
        return type(d) is P.StandardNormal and d.event_size == dz

    @classmethod
    def _check_prior(cls, prior, dz):
        if isinstance(prior, dists.FlowedDistribution):
            flow, latent = prior.flow, prior.latent_dist
            if not isinstance(flow, flows.RQSSplineRealNVP) or flow.batch_norm or flow.before_flow_transform is not None \
                    or flow.after_flow_transform is not None:
                raise NotImplementedError('fused ELBO: prior flow must be a plain RQSSplineRealNVP')
            if not flow.built:
                raise RuntimeError('fused ELBO: call the model once on data before training so the flow is built')
            if not cls._is_standard_normal(latent, dz):
                raise NotImplementedError('fused ELBO: the flow\'s latent distribution must be a static N(0, I) '
                                          '(DistributionLambda -> StandardNormal of the latent size)')
            blocks = [b.bijector_fn for b in flow.chain.bijectors[::-1]]  # chain list is reversed (flows.py:323)
            b0 = blocks[0]
            return blocks, b0.num_bins, b0.hidden_dim, float(b0.bin_min), float(b0.bin_max)
        if cls._is_standard_normal(prior, dz):
            return [], 2, 1, -1.0, 1.0
        raise NotImplementedError('fused ELBO: prior must be a static N(0, I) DistributionLambda (-> StandardNormal of the '
                                  'latent size) or a FlowedDistribution over one')

    def invalidate(self):
        """Tell the plan that theta changed behind its back (its packed weight images must be rebuilt)."""
        if getattr(self, 'handle', None):
            ctx().lib.vms_elbo_plan_invalidate(self.handle)

    def set_mode(self, mode):
        """0 = auto (whole-step tensor-core kernel for training steps up to three waves of tiles, the FFMA fused kernel for forward
        evaluation, the tensor-core plan above), 1 = force the unfused per-layer float32-FFMA graph path, 2 = unfused plan
        with the coupling blocks as fused tcgen05 kernels (the large-batch configuration), 3 = force the whole-step
        tensor-core kernel, 4 = force the single FFMA fused kernel."""
        ctx().lib.vms_elbo_plan_set_mode(self.handle, int(mode))

    def set_tc_auto_batch(self, batch):
        """Batch from which auto mode (0) prefers the tensor-core plan over the fused kernels (default: above one wave of
        32-row tiles for the FFMA kernel, above three waves for the whole-step tensor-core kernel)."""
        ctx().lib.vms_elbo_plan_set_tc_auto_batch(self.handle, int(batch))

    def path(self, batch):
        """Implementation a step of `batch` rows takes: 'fused' (one persistent kernel), 'ffma' (per-layer float32
        plan) or 'tensor-core' (large-batch plan: tcgen05 coupling blocks + streaming MLP kernels)."""
        return ('fused', 'ffma', 'tensor-core', 'tensor-core-fused')[
            ctx().lib.vms_elbo_plan_path(self.handle, int(batch))]

    def tc_status(self):
        """True if a tensor-core kernel of the mode-2 plan gave up waiting for an MMA completion (results invalid)."""
        err = C.c_int(0)
        ctx().lib.vms_elbo_plan_tc_status(self.handle, C.byref(err))
        return bool(err.value)

    @property
    def is_fused(self):
        return bool(ctx().lib.vms_elbo_plan_is_fused(self.handle))

    def forward(self, x, eps, want=('z', 'logq', 'logpz', 'logpx')):
        c = ctx()
        B = x.shape[0]
        out = {'z': Tensor((B, self.dz)) if 'z' in want else None}
        for k in ('logq', 'logpz', 'logpx'):
            out[k] = Tensor((B, )) if k in want else None
        out['scalars'] = Tensor((4, ))
        P_ = P._ptr
        c.lib.vms_elbo_forward(self.handle, self.theta.ptr, x.ptr, eps.ptr, B, P_(out['z']), P_(out['logq']),
                               P_(out['logpz']), P_(out['logpx']), out['scalars'].ptr, c.stream)
        return out

    def forward_backward(self, x, eps, grad_ptr=None):
        """Writes the flat gradient into self.grad (or to the device address `grad_ptr`, e.g. a slot of the peer
        exchange buffer) and {loss, nll, kl} into self.scalars (device)."""
        c = ctx()
        c.lib.vms_elbo_forward_backward(self.handle, self.theta.ptr, x.ptr, eps.ptr, x.shape[0],
                                        self.grad.ptr if grad_ptr is None else grad_ptr, self.scalars.ptr, c.stream)
        return self.scalars

    def adam_step(self, opt, grad_scale=1.0):
        _abi.bump_param_epoch()
        c = ctx()
        self.t += 1
        c.lib.vms_adam_step(self.theta.ptr, self.grad.ptr, 1, grad_scale, self.m.ptr, self.v.ptr, self.n_params, self.t,
                            opt.learning_rate, opt.beta_1, opt.beta_2, opt.epsilon, c.stream)

    def train_loop(self, x_host, opt, batch_size, order=None, eps_host=None, n_steps=None, seed=None, exchange=None):
        """The inner loop of `VAE.fit` (the Keras training loop of tests/test_models.py:181-182) as a device pipeline:
        batch i + 1 travels host -> device on a copy stream while batch i trains (double-buffered inputs, events in both
        directions), the reparameterisation noise is drawn on the device (Philox, `vms_standard_normal`) unless `eps_host`
        is given, and the per-step scalars {loss, nll, kl} land in a small device ring that is read back every 8 steps,
        so the host never waits for a step.  x_host [N, dx] float32 (pinned memory avoids a staging copy when `order` is
        None); order: row permutation (shuffle); n_steps: number of batches (default: one epoch, ragged last batch
        dropped into its own step).  exchange: a `parallel.PeerExchange` -- data-parallel training: the step becomes
        forward + backward into this rank's slot of the exchange buffer followed by the fused NVLink allreduce + Adam kernel
        (every rank must run the same number of steps); the scalars are this rank's shard means.  Returns the
        [n_steps, 3] scalars."""
        _abi.bump_param_epoch()
        c = ctx()
        lib = c.lib
        N = x_host.shape[0]
        bs = int(min(batch_size, N))
        starts = list(range(0, N, bs))
        if n_steps is not None:
            starts = [starts[i % len(starts)] for i in range(int(n_steps))]
        R = 8
        if getattr(self, '_loop', None) is not None and self._loop['bs'] < bs:
            self._release_loop()
        if getattr(self, '_loop', None) is None:
            mk = lambda: C.c_void_p()
            st = mk()
            lib.vms_stream_create(C.byref(st))
            evs = []
            for _ in range(4):
                e = mk()
                lib.vms_event_create(C.byref(e))
                evs.append(e.value)
            hp = mk()
            lib.vms_malloc_host(C.byref(hp), R * 4 * 4 * 2)
            stage = []
            for _ in range(2):
                p = mk()
                lib.vms_malloc_host(C.byref(p), bs * (self.dx + self.dz) * 4)
                stage.append(p.value)
            self._loop = dict(bs=bs, copy_stream=st.value, h2d=evs[:2], free=evs[2:], xd=[Tensor((bs, self.dx)) for _ in range(2)],
                              ed=[Tensor((bs, self.dz)) for _ in range(2)], ring=Tensor((R, 4)), host_ring=hp.value,
                              stage=stage, noise_offset=0,
                              noise_seed=int(P.rng().integers(0, 2**63 - 1)) if seed is None else int(seed))
        L = self._loop
        if seed is not None:
            L['noise_seed'], L['noise_offset'] = int(seed), 0
        cs = L['copy_stream']
        x_host = np.ascontiguousarray(x_host, np.float32)
        if eps_host is not None:
            eps_host = np.ascontiguousarray(eps_host, np.float32)
        out = np.empty((len(starts), 4), np.float32)
        ring_host = np.frombuffer((C.c_byte * (R * 4 * 4 * 2)).from_address(L['host_ring']), np.float32).reshape(2, R, 4)
        pending = []  # (first step, count, half) of ring read-backs in flight
        used = [False, False]

        def drain(keep):
            while len(pending) > keep:
                first, cnt, half, ev = pending.pop(0)
                lib.vms_event_synchronize(ev)
                out[first:first + cnt] = ring_host[half, :cnt]

        ring_ev = []
        for _ in range(2):
            e = C.c_void_p()
            lib.vms_event_create(C.byref(e))
            ring_ev.append(e.value)
        for i, s0 in enumerate(starts):
            b = i & 1
            n = min(bs, N - s0)
            # ---- copy stream: inputs of step i into buffer b (once step i - 2 released it)
            if used[b]:
                lib.vms_stream_wait_event(cs, L['free'][b])
            if order is None:
                src_x = x_host.ctypes.data + s0 * self.dx * 4
            else:
                if used[b]:
                    lib.vms_event_synchronize(L['h2d'][b])  # the staging buffer's previous copy has left the host
                sx = np.frombuffer((C.c_byte * (n * self.dx * 4)).from_address(L['stage'][b]), np.float32).reshape(n, self.dx)
                np.take(x_host, order[s0:s0 + n], axis=0, out=sx)
                src_x = L['stage'][b]
            lib.vms_memcpy_h2d(L['xd'][b].ptr, src_x, n * self.dx * 4, cs)
            if eps_host is not None:
                lib.vms_memcpy_h2d(L['ed'][b].ptr, eps_host.ctypes.data + s0 * self.dz * 4, n * self.dz * 4, cs)
            lib.vms_event_record(L['h2d'][b], cs)
            # ---- compute stream: noise, the step, release of the buffer
            if eps_host is None:
                lib.vms_standard_normal(L['noise_seed'], L['noise_offset'], n * self.dz, L['ed'][b].ptr, c.stream)
                L['noise_offset'] += n * self.dz
            lib.vms_stream_wait_event(c.stream, L['h2d'][b])
            slot = i % R
            if exchange is None:
                self.t += 1
                lib.vms_elbo_train_step(self.handle, self.theta.ptr, L['xd'][b].ptr, L['ed'][b].ptr, n, self.grad.ptr,
                                        L['ring'].ptr + 16 * slot, self.m.ptr, self.v.ptr, self.t, opt.learning_rate,
                                        opt.beta_1, opt.beta_2, opt.epsilon, c.stream)
            else:
                exchange.train_step(self, L['xd'][b], L['ed'][b], n, opt, scalars_ptr=L['ring'].ptr + 16 * slot)  # advances self.t
            lib.vms_event_record(L['free'][b], c.stream)
            used[b] = True
            if slot == R - 1 or i == len(starts) - 1:
                half = (i // R) & 1
                drain(1)  # at most one read-back outstanding before its host half is reused
                lib.vms_memcpy_d2h(L['host_ring'] + half * R * 16, L['ring'].ptr, (slot + 1) * 16, c.stream)
                lib.vms_event_record(ring_ev[half], c.stream)
                pending.append((i - slot, slot + 1, half, ring_ev[half]))
        drain(0)
        lib.vms_stream_synchronize(cs)
        for e in ring_ev:
            lib.vms_event_destroy(e)
        lib.vms_memcpy_d2d(self.scalars.ptr, L['ring'].ptr + 16 * ((len(starts) - 1) % R), 16, c.stream)
        return out[:, :3]

    def train_step(self, x, eps, opt):
        """Forward + backward + Adam in one C call (`vms_elbo_train_step`: 2 kernel launches on the fused path)."""
        _abi.bump_param_epoch()
        c = ctx()
        self.t += 1
        c.lib.vms_elbo_train_step(self.handle, self.theta.ptr, x.ptr, eps.ptr, x.shape[0], self.grad.ptr, self.scalars.ptr,
                                  self.m.ptr, self.v.ptr, self.t, opt.learning_rate, opt.beta_1, opt.beta_2, opt.epsilon,
                                  c.stream)
        return self.scalars
