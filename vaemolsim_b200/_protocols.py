"""Host-side objects that speak the TFP Bijector / Distribution protocols the reference's layers return and its
losses / models / MCMC driver consume, backed by the sm_100a kernels behind include/vms_b200.h.

Nothing here does arithmetic on the host except parameter bookkeeping (masks, offsets, RNG for noise that is an INPUT
to the kernels).  Every log_prob / forward / inverse / log-det call is one or a few kernel launches.

Reference counterparts (third-party objects constructed at these call sites):
  tfp.bijectors.RationalQuadraticSpline   flows.py:204-207, :512-515
  tfp.bijectors.RealNVP                   flows.py:310-315
  tfp.bijectors.MaskedAutoregressiveFlow  flows.py:626-637
  tfp.bijectors.AutoregressiveNetwork     flows.py:454-487, dists.py:301-305
  tfp.bijectors.Chain / Shift / Scale     flows.py:53-58, :323, :644
  tfp.distributions.TransformedDistribution  flows.py:353, :684
  tfp.distributions.Blockwise / Autoregressive / Independent(Normal|VonMises|Deterministic)
                                          dists.py:217, :336-340, :604-610, :701
  tf.keras.layers.Dense                   flows.py:136-152, mappings.py:107-121
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import Tensor, as_tensor, ctx

ACT = {None: 0, 'linear': 0, 'relu': 1, 'tanh': 2}
DIST_NORMAL, DIST_VONMISES = 0, 1
SCALE_IDENTITY, SCALE_SOFTPLUS, SCALE_SOFTPLUS_EPS = 0, 1, 2

_rng = np.random.default_rng()


def set_seed(seed):
    """Seeds the host generator used for weight initialisation and sampling noise."""
    global _rng
    _rng = np.random.default_rng(seed)


def rng():
    if _abi._capturing[0]:  # host-drawn noise cannot be part of a captured training step
        raise _abi.CaptureUnsupported('host random numbers')
    return _rng


def _act_code(activation):
    if callable(activation):
        activation = getattr(activation, '__name__', None)
    if activation not in ACT:
        raise ValueError('unsupported activation %r (supported: None, relu, tanh)' % (activation,))
    return ACT[activation]


def _tape():
    from . import _autodiff
    return _autodiff.Tape.active()


def _ad():
    from . import _autodiff
    return _autodiff


def _ptr(t):
    return None if t is None else t.ptr


# data-parallel group of the training step in progress (set by _autodiff.Trainer.step): batch-normalisation layers then use
# CROSS-REPLICA batch statistics, so that sharded training equals training on the whole batch
_dp_group = None


def _sync_group():
    g = _dp_group
    return g if (g is not None and getattr(g, 'world', 1) > 1) else None


# ================================================================================================ layers
class Layer(object):
    """Minimal stand-in for tf.keras.layers.Layer: lazy build on first call, named, get_config, weights."""

    def __init__(self, name=None, **kwargs):
        if kwargs:
            raise TypeError('unexpected keyword arguments %s' % sorted(kwargs))
        self.name = name if name is not None else type(self).__name__.lower()
        self.built = False

    def build(self, input_shape):
        pass

    @staticmethod
    def _shape_of(x):
        if isinstance(x, (list, tuple)) and len(x) and not np.isscalar(x[0]):
            return [Layer._shape_of(v) for v in x]
        if hasattr(x, 'shape'):
            return tuple(x.shape)
        if isinstance(x, Distribution):
            return (x.batch, x.event_size)
        return None

    def __call__(self, *args, **kwargs):
        if not self.built:
            self.build(self._shape_of(args[0]) if args else None)
            self.built = True
        if 'training' in kwargs and not self._takes_training():
            kwargs.pop('training')  # Keras drops `training` for layers whose call does not take it
        return self.call(*args, **kwargs)

    def _takes_training(self):
        t = getattr(type(self), '_takes_training_cache', None)
        if t is None or t[0] is not type(self).call:
            import inspect
            ps = inspect.signature(type(self).call).parameters
            ok = 'training' in ps or any(p.kind == p.VAR_KEYWORD for p in ps.values())
            type(self)._takes_training_cache = t = (type(self).call, ok)
        return t[1]

    def get_config(self):
        return {'name': self.name}

    def _sublayers(self):
        seen = []
        for v in self.__dict__.values():
            vs = v if isinstance(v, (list, tuple)) else [v]
            for u in vs:
                if isinstance(u, Layer) and u not in seen:
                    seen.append(u)
                elif isinstance(u, Bijector):
                    for w in u._layers():
                        if w not in seen:
                            seen.append(w)
        return seen

    @property
    def weights(self):
        out = list(getattr(self, '_weights', []))
        for sub in self._sublayers():
            out += sub.weights
        return out

    trainable_weights = weights

    def count_params(self):
        return int(sum(w.size for w in self.weights))


def _init_kernel(kind, fan_in, fan_out):
    r = rng()
    if kind == 'glorot_uniform':
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return r.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)
    if kind == 'truncated_normal':  # Keras string initialiser: stddev 0.05, resampled beyond 2 sigma
        out = r.normal(0.0, 0.05, size=(fan_in, fan_out))
        bad = np.abs(out) > 0.1
        while bad.any():
            out[bad] = r.normal(0.0, 0.05, size=int(bad.sum()))
            bad = np.abs(out) > 0.1
        return out.astype(np.float32)
    if kind == 'zeros':
        return np.zeros((fan_in, fan_out), np.float32)
    raise ValueError('unsupported kernel_initializer %r' % (kind,))


class Dense(Layer):
    """tf.keras.layers.Dense on the device: kernel [in, units], bias [units] (vms_dense_forward)."""

    def __init__(self, units, activation=None, name='dense', kernel_initializer='glorot_uniform', use_bias=True):
        super(Dense, self).__init__(name=name)
        self.units = int(units)
        self.act = _act_code(activation)
        self.kernel_initializer = kernel_initializer
        self.use_bias = use_bias
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        fan_in = max(int(input_shape[-1]), 1)
        self.kernel = Tensor.from_numpy(_init_kernel(self.kernel_initializer, fan_in, self.units))
        self.bias = Tensor.zeros((self.units,)) if self.use_bias else None
        self._weights = [self.kernel] + ([self.bias] if self.use_bias else [])

    def set_weights(self, arrays):
        self.kernel = Tensor.from_numpy(np.asarray(arrays[0], np.float32))
        if self.use_bias:
            self.bias = Tensor.from_numpy(np.asarray(arrays[1], np.float32))
        self._weights = [self.kernel] + ([self.bias] if self.use_bias else [])
        self.built = True
        _abi.bump_param_epoch()

    def get_weights(self):
        return [w.numpy() for w in self._weights]

    def assign(self, kernel, bias=None):
        """Overwrite the weights IN PLACE (keeps device pointers, e.g. views into a flat parameter buffer)."""
        c = ctx()
        k = np.ascontiguousarray(kernel, np.float32)
        if k.shape != self.kernel.shape:
            raise ValueError('kernel shape %s != %s' % (k.shape, self.kernel.shape))
        c.lib.vms_memcpy_h2d(self.kernel.ptr, k.ctypes.data, k.nbytes, c.stream)
        if bias is not None and self.use_bias:
            b = np.ascontiguousarray(bias, np.float32)
            c.lib.vms_memcpy_h2d(self.bias.ptr, b.ctypes.data, b.nbytes, c.stream)
        c.synchronize()
        _abi.bump_param_epoch()
        hook = getattr(self, '_on_assign', None)
        if hook is not None:
            hook()

    def rebind(self, kernel, bias):
        """Point the layer at externally owned storage (flat parameter buffer views)."""
        self.kernel, self.bias = kernel, bias
        _abi.bump_param_epoch()
        self._weights = [self.kernel] + ([self.bias] if self.use_bias else [])

    def call(self, x, ones_input=False, cond=None, cond_kernel=None):
        if not ones_input and x.shape[1] != self.kernel.shape[0]:
            raise ValueError('Dense %s expects last dimension %d, got %d' % (self.name, self.kernel.shape[0], x.shape[1]))
        return dense_op(x, self.kernel, self.bias, self.act, ones_input=ones_input, cond=cond, cond_kernel=cond_kernel)


def dense_op(x, W, bias=None, act=0, ones_input=False, cond=None, cond_kernel=None):
    """act(x @ W + bias + cond @ cond_kernel) on the device (vms_dense_forward), recorded on the active tape
    (vms_dense_backward: input, kernel, bias and conditional-input / conditional-kernel gradients, as TF autodiff)."""
    c = ctx()
    B = x.shape[0]
    K, N = W.shape[0], W.shape[1]
    out = Tensor((B, N))
    xp, ldx = (None, 1) if ones_input else (x.ptr, x.ld)
    Cn = 0 if cond is None else cond.shape[1]
    c.lib.vms_dense_forward(xp, ldx, W.ptr, _ptr(bias), B, K, N, act, _ptr(cond), 0 if cond is None else cond.ld,
                            _ptr(cond_kernel), Cn, out.ptr, out.ld, c.stream)
    tp = _tape()
    if tp is not None:

        def bw():
            if not tp.has(out):
                return
            g_out = tp.grad(out)
            ws = Tensor((max(int(c.lib.vms_dense_backward_workspace(B, K, N, Cn)) // 4, 1), ))
            g_x = None if ones_input else tp.grad(x)
            g_c = Tensor((B, Cn)) if cond is not None else None  # (the kernel overwrites g_cond: accumulate below)
            c.lib.vms_dense_backward(xp, ldx, W.ptr, B, K, N, act, out.ptr, out.ld, g_out.ptr, g_out.ld, _ptr(cond),
                                     0 if cond is None else cond.ld, _ptr(cond_kernel), Cn, _ptr(g_x),
                                     0 if g_x is None else g_x.ld, 1, tp.grad(W).ptr,
                                     None if bias is None else tp.grad(bias).ptr, _ptr(g_c), Cn,
                                     None if cond_kernel is None else tp.grad(cond_kernel).ptr, 1, ws.ptr, c.stream)
            if g_c is not None:
                _ad().add_into(tp.grad(cond), g_c, 1.0)

        tp.record(bw)
    return out


class LayerNormalization(Layer):
    """tf.keras.layers.LayerNormalization() over the last axis (Keras defaults: epsilon 1e-3, gamma 1, beta 0), optionally
    with the `tf.keras.layers.Activation` that follows it in the reference's networks (mappings.py:511-512) folded into the
    same kernel (vms_layernorm_forward / _backward)."""

    def __init__(self, epsilon=1e-3, name='layer_normalization'):
        super(LayerNormalization, self).__init__(name=name)
        self.epsilon = float(epsilon)
        self.gamma = self.beta = None

    def build(self, input_shape):
        H = int(input_shape[-1])
        self.gamma, self.beta = Tensor.from_numpy(np.ones(H, np.float32)), Tensor.zeros((H,))
        self._weights = [self.gamma, self.beta]

    def call(self, x, activation=0):
        c = ctx()
        R, H = x.shape
        y, stats = Tensor((R, H)), Tensor((R, 2))
        c.lib.vms_layernorm_forward(x.ptr, x.ld, R, H, self.gamma.ptr, self.beta.ptr, self.epsilon, activation, y.ptr, y.ld,
                                    stats.ptr, c.stream)
        tp = _tape()
        if tp is not None:
            gamma, beta = self.gamma, self.beta

            def bw():
                if not tp.has(y):
                    return
                ws = Tensor((max(int(c.lib.vms_layernorm_backward_workspace(R, H)) // 4, 1), ))
                g_y, g_x = tp.grad(y), tp.grad(x)
                c.lib.vms_layernorm_backward(x.ptr, x.ld, R, H, gamma.ptr, stats.ptr, activation, y.ptr, y.ld, g_y.ptr, g_y.ld,
                                             g_x.ptr, g_x.ld, tp.grad(gamma).ptr, tp.grad(beta).ptr, ws.ptr, c.stream)

            tp.record(bw)
        return y


class Activation(Layer):
    """tf.keras.layers.Activation(fn); inside a `Sequential` it is folded into the LayerNormalization before it."""

    def __init__(self, activation, name='activation'):
        super(Activation, self).__init__(name=name)
        self.act = _act_code(activation)
        self._weights = []

    def call(self, x):
        if self.act == 0:
            return x
        raise NotImplementedError('a stand-alone Activation layer is only supported directly after LayerNormalization')


class Sequential(Layer):
    """tf.keras.models.Sequential over Dense / LayerNormalization / Activation layers acting on the last axis of a 2-D
    tensor (the score / value / nonlinearity networks of mappings.py:505-532)."""

    def __init__(self, layers, name='sequential'):
        super(Sequential, self).__init__(name=name)
        self.layers = list(layers)

    def build(self, input_shape):
        shape = tuple(input_shape)
        for lay in self.layers:
            if not lay.built:
                lay.build(shape)
                lay.built = True
            if isinstance(lay, Dense):
                shape = shape[:-1] + (lay.units,)

    def call(self, x):
        out, skip = x, False
        for i, lay in enumerate(self.layers):
            if skip:
                skip = False
                continue
            nxt = self.layers[i + 1] if i + 1 < len(self.layers) else None
            if isinstance(lay, LayerNormalization) and isinstance(nxt, Activation):
                out, skip = lay.call(out, activation=nxt.act), True
            else:
                out = lay.call(out)
        return out


class _BatchNormState(object):
    """gamma / beta / moving statistics of one batch-normalisation layer over D features and the device plumbing both
    variants share (csrc/batchnorm.cu).  Keras defaults: momentum 0.99, epsilon 1e-3, gamma 1, beta 0, moving mean 0,
    moving variance 1 (what `tf.keras.layers.BatchNormalization()` at mappings.py:114 and the default layer inside
    `tfp.bijectors.BatchNormalization` at flows.py:309 / :624 create)."""

    def __init__(self, D, momentum=0.99, epsilon=1e-3):
        self.D, self.momentum, self.epsilon = int(D), float(momentum), float(epsilon)
        self.gamma, self.beta = Tensor.from_numpy(np.ones(D, np.float32)), Tensor.zeros((D,))
        self.moving_mean, self.moving_variance = Tensor.zeros((D,)), Tensor.from_numpy(np.ones(D, np.float32))
        self._scale, self._shift, self._ldj = Tensor((D,)), Tensor((D,)), Tensor((1,))
        self._mean, self._var = Tensor((D,)), Tensor((D,))

    def weights(self):
        return [self.gamma, self.beta, self.moving_mean, self.moving_variance]

    def statistics(self, x, training):
        """(mean, var) to normalise with: the batch moments (moving statistics updated) when training, else the
        moving statistics."""
        if not training:
            return self.moving_mean, self.moving_variance
        c = ctx()
        B = x.shape[0]
        ws = Tensor(((int(c.lib.vms_batch_moments_workspace(B, self.D)) + 3) // 4,))
        c.lib.vms_batch_moments(x.ptr, x.ld, B, self.D, self._mean.ptr, self._var.ptr, ws.ptr, c.stream)
        grp = _sync_group()
        if grp is not None:  # global moments: two small allreduces (means, then the parallel-variance combination)
            D = self.D
            buf, mean_r = Tensor((D + 1,)), self._mean.copy()
            c.lib.vms_bn_sync_pack(mean_r.ptr, None, None, B, D, buf.ptr, c.stream)
            grp.allreduce_sum_device_(buf.ptr, D + 1, c.stream)
            c.lib.vms_bn_sync_unpack(buf.ptr, D, self._mean.ptr, c.stream)
            c.lib.vms_bn_sync_pack(mean_r.ptr, self._var.ptr, self._mean.ptr, B, D, buf.ptr, c.stream)
            grp.allreduce_sum_device_(buf.ptr, D + 1, c.stream)
            c.lib.vms_bn_sync_unpack(buf.ptr, D, self._var.ptr, c.stream)
        for mov, cur in ((self.moving_mean, self._mean), (self.moving_variance, self._var)):
            c.lib.vms_axpby(mov.ptr, cur.ptr, self.momentum, 1.0 - self.momentum, self.D, mov.ptr, c.stream)
        c.synchronize()  # the workspace is released when this frame returns
        return self._mean, self._var

    def apply(self, x, mean, var, denormalize, want_ldj=False):
        c = ctx()
        B = x.shape[0]
        c.lib.vms_batchnorm_coeffs(mean.ptr, var.ptr, self.gamma.ptr, self.beta.ptr, self.D, self.epsilon,
                                   1 if denormalize else 0, self._scale.ptr, self._shift.ptr,
                                   self._ldj.ptr if want_ldj else None, c.stream)
        out = Tensor((B, self.D))
        c.lib.vms_affine_cols(x.ptr, x.ld, B, self.D, self._scale.ptr, self._shift.ptr, 0, out.ptr, out.ld, c.stream)
        ldj = None
        if want_ldj:
            ldj = Tensor((B,))
            c.lib.vms_broadcast_scalar(self._ldj.ptr, B, ldj.ptr, c.stream)
        tp = _tape()
        if tp is not None:
            if denormalize:
                raise NotImplementedError('reverse mode through the de-normalising direction of the batch-norm bijector '
                                          '(sampling direction) is not built')
            batch_stats = mean is self._mean  # statistics() handed out the batch moments: gradients flow through them
            m_, v_ = (mean.copy(), var.copy()) if batch_stats else (mean, var)  # the layer's buffers are reused per call
            D, eps, gamma, beta = self.D, self.epsilon, self.gamma, self.beta

            def bw():  # TF autodiff through tf.nn.batch_normalization + tf.nn.moments (and the bijector's log-det)
                has_o, has_l = tp.has(out), (ldj is not None and tp.has(ldj))
                if not (has_o or has_l):
                    return
                g_out = tp.grad(out) if has_o else Tensor.zeros((B, D))
                G = None
                if has_l:
                    G = Tensor.zeros((1, ))
                    c.lib.vms_sum_all(tp.grad(ldj).ptr, B, 1.0, G.ptr, c.stream)
                ws = Tensor((int(c.lib.vms_batchnorm_backward_workspace(D)) // 4 + 1, ))
                g_x = tp.grad(x)
                grp = _sync_group() if batch_stats else None
                if grp is not None:  # cross-replica statistics: the reverse mode needs the column sums of ALL ranks
                    loc, glob = Tensor((2 * D + 2,)), Tensor((2 * D + 2,))
                    c.lib.vms_batchnorm_backward_sums(x.ptr, x.ld, B, D, m_.ptr, v_.ptr, eps, g_out.ptr, g_out.ld, _ptr(G), loc.ptr,
                                                      ws.ptr, c.stream)
                    c.lib.vms_memcpy_d2d(glob.ptr, loc.ptr, 4 * (2 * D + 2), c.stream)
                    grp.allreduce_sum_device_(glob.ptr, 2 * D + 2, c.stream)
                    c.lib.vms_batchnorm_backward_apply(x.ptr, x.ld, B, D, m_.ptr, v_.ptr, gamma.ptr, eps, g_out.ptr, g_out.ld,
                                                       loc.ptr, glob.ptr, g_x.ptr, g_x.ld, tp.grad(gamma).ptr, tp.grad(beta).ptr,
                                                       c.stream)
                    return
                c.lib.vms_batchnorm_backward(x.ptr, x.ld, B, D, m_.ptr, v_.ptr, gamma.ptr, eps, 1 if batch_stats else 0,
                                             g_out.ptr, g_out.ld, _ptr(G), g_x.ptr, g_x.ld, tp.grad(gamma).ptr,
                                             tp.grad(beta).ptr, ws.ptr, c.stream)

            tp.record(bw)
        return out, ldj


class KerasBatchNormalization(Layer):
    """tf.keras.layers.BatchNormalization() over the last axis of a [B, D] tensor (mappings.py:113-114): batch moments
    and a moving-average update when training, moving statistics otherwise; reverse mode through the batch statistics by
    `vms_batchnorm_backward` when a tape is active."""

    def __init__(self, name='batch_normalization'):
        super(KerasBatchNormalization, self).__init__(name=name)
        self.state = None

    def build(self, input_shape):
        self.state = _BatchNormState(int(input_shape[-1]))
        self._weights = self.state.weights()

    def call(self, x, training=False):
        if self.state is None:
            self.build(x.shape)
            self.built = True
        mean, var = self.state.statistics(x, training)
        return self.state.apply(x, mean, var, denormalize=False)[0]


class Reshape(Layer):
    """tf.keras.layers.Reshape(target_shape): the last entry of FCDeepNN.layer_list (mappings.py:123)."""

    def __init__(self, target_shape, name='reshape'):
        super(Reshape, self).__init__(name=name)
        self.target_shape = tuple(target_shape)
        self._weights = []

    def call(self, x, training=False):
        return x.reshape((x.shape[0],) + self.target_shape)


# ---------------------------------------------------------------------------------------- MADE (AutoregressiveNetwork)
def _input_order(event_size, input_order):
    if isinstance(input_order, str):
        if input_order == 'left-to-right':
            return np.arange(1, event_size + 1)
        if input_order == 'right-to-left':
            return np.arange(event_size, 0, -1)
        if input_order == 'random':
            o = np.arange(1, event_size + 1)
            rng().shuffle(o)
            return o
        raise ValueError('Invalid input order: "%s".' % input_order)
    order = np.array(input_order)
    if order.shape != (event_size,) or not np.all(np.sort(order) == np.arange(1, event_size + 1)):
        raise ValueError('Invalid input order: %r' % (input_order,))
    return order


def made_masks(params, event_size, hidden_units, input_order):
    """tfp `_make_dense_autoregressive_masks` (hidden_degrees='equal'): list of [in, out] boolean masks."""
    deg = [_input_order(event_size, input_order)]
    for units in hidden_units:
        min_degree = min(int(np.min(deg[-1])), event_size - 1)
        deg.append(np.maximum(min_degree, np.ceil(np.arange(1, units + 1) * (event_size - 1) / float(units + 1)).astype(
            np.int32)))
    masks = [a[:, None] <= b[None, :] for a, b in zip(deg[:-1], deg[1:])]
    last = deg[-1][:, None] < deg[0][None, :]
    last = np.reshape(np.tile(last[..., None], [1, 1, params]), [last.shape[0], event_size * params])
    return masks + [last]


class AutoregressiveNetwork(Layer):
    """tfp.bijectors.AutoregressiveNetwork (MADE): pre-masked Dense layers; the optional conditional input enters
    bias-free into EVERY layer (`conditional_input_layers='all_layers'`; pinned by the notebook parameter counts)."""

    def __init__(self, params, event_shape=None, conditional=False, conditional_event_shape=None,
                 input_order='left-to-right', hidden_units=None, activation=None, name='autoregressive_network',
                 kernel_initializer='glorot_uniform'):
        super(AutoregressiveNetwork, self).__init__(name=name)
        self.params = int(params)
        self._event_size = None if event_shape is None else int(np.prod(event_shape))
        self.conditional = conditional
        if conditional and conditional_event_shape is None:
            raise ValueError('`conditional_event_shape` must be provided when `conditional` is True')
        self.cond_size = int(np.prod(conditional_event_shape)) if conditional else 0
        self.input_order = input_order
        self.hidden_units = list(hidden_units) if hidden_units is not None else []
        self.act = _act_code(activation)
        self.kernel_initializer = kernel_initializer
        self.layers = []
        self.cond_kernels = []

    @property
    def event_shape(self):
        return (self._event_size,)

    def build(self, input_shape):
        if self._event_size is None:
            self._event_size = int(input_shape[-1])
        D = self._event_size
        self.masks = made_masks(self.params, D, self.hidden_units, self.input_order)
        sizes = [D] + self.hidden_units + [D * self.params]
        self._weights = []
        for k, (a, b) in enumerate(zip(sizes[:-1], sizes[1:])):
            lay = Dense(b, activation=None, name='%s_dense_%d' % (self.name, k))
            W = _init_kernel(self.kernel_initializer, a, b) * self.masks[k].astype(np.float32)
            lay.set_weights([W, np.zeros(b, np.float32)])
            lay.kernel._grad_mask = Tensor.from_numpy(self.masks[k].astype(np.float32))  # applied to the kernel's gradient
            lay.act = self.act if k + 1 < len(sizes) - 1 else 0
            self.layers.append(lay)
            self._weights += lay._weights
            if self.conditional:
                Wc = Tensor.from_numpy(_init_kernel(self.kernel_initializer, self.cond_size, b))
                self.cond_kernels.append(Wc)
                self._weights.append(Wc)

    def _sublayers(self):
        return []

    def set_weights(self, arrays):
        """arrays per layer: kernel (masked here, as tfp's constraint does), bias, [conditional kernel]."""
        it = iter(arrays)
        self._weights = []
        for k, lay in enumerate(self.layers):
            W = np.asarray(next(it), np.float32) * self.masks[k].astype(np.float32)
            lay.set_weights([W, np.asarray(next(it), np.float32)])
            lay.kernel._grad_mask = Tensor.from_numpy(self.masks[k].astype(np.float32))
            self._weights += lay._weights
            if self.conditional:
                self.cond_kernels[k] = Tensor.from_numpy(np.asarray(next(it), np.float32))
                self._weights.append(self.cond_kernels[k])
        _abi.bump_param_epoch()

    def call(self, x, conditional_input=None):
        if self.conditional and conditional_input is None:
            raise ValueError('`conditional_input` must be passed as a named argument.')
        cond = as_tensor(conditional_input) if self.conditional else None
        if cond is not None and cond.ndim != 2:
            cond = cond.reshape(cond.shape[0], -1)
        out = x
        for k, lay in enumerate(self.layers):
            out = lay.call(out, cond=cond, cond_kernel=self.cond_kernels[k] if self.conditional else None)
        return out.reshape(x.shape[0], self._event_size, self.params)


# ================================================================================================ bijectors
class Bijector(object):
    """TFP bijector protocol.  Subclasses implement `_fwd(x, **kw) -> (y, fldj[B])` and `_inv(y, **kw) -> (x, ildj[B])`
    (log-dets summed over the event, i.e. event_ndims=1) in a single fused pass."""
    name = 'bijector'

    def forward(self, x, **kw):
        return self._fwd(as_tensor(x), **kw)[0]

    def inverse(self, y, **kw):
        return self._inv(as_tensor(y), **kw)[0]

    def forward_log_det_jacobian(self, x, event_ndims=1, **kw):
        return self._fwd(as_tensor(x), **kw)[1]

    def inverse_log_det_jacobian(self, y, event_ndims=1, **kw):
        return self._inv(as_tensor(y), **kw)[1]

    def __call__(self, v, **kw):
        if isinstance(v, Distribution):
            return TransformedDistribution(v, self)
        return self.forward(v, **kw)

    def _layers(self):
        return []


class RationalQuadraticSpline(Bijector):
    """RQS bijector over the RAW conditioner outputs (activations fused in the kernel, flows.py:86-101).

    raw_w, raw_h: Tensor views [B, n_dims*K]; raw_s: [B, n_dims*(K-1)]; range = [range_min, range_max]."""

    def __init__(self, raw_w, raw_h, raw_s, n_dims, num_bins, range_min, range_max, name='rqs'):
        self.raw_w, self.raw_h, self.raw_s = raw_w, raw_h, raw_s
        self.n_dims, self.num_bins = int(n_dims), int(num_bins)
        self.range_min, self.range_max = float(range_min), float(range_max)
        self.name = name

    def apply_into(self, v, out, ldj_sum, accumulate, inverse, ldj=None):
        c = ctx()
        B = v.shape[0]
        if v.shape[1] != self.n_dims or self.raw_w.shape[0] != B:
            raise ValueError('RationalQuadraticSpline: input shape %s does not match parameters (%d, %d)' %
                             (v.shape, self.raw_w.shape[0], self.n_dims))
        a = _abi.RqsArgs(B, self.n_dims, self.num_bins, self.range_min, self.range_max, v.ptr, v.ld, self.raw_w.ptr,
                         self.raw_w.ld, self.raw_h.ptr, self.raw_h.ld, self.raw_s.ptr, self.raw_s.ld, out.ptr, out.ld,
                         _ptr(ldj), _ptr(ldj_sum), 1 if accumulate else 0, 1 if inverse else 0)
        c.lib.vms_rqs_apply(C.byref(a), c.stream)
        tp = _tape()
        if tp is not None:
            rw, rh, rs, nd, K = self.raw_w, self.raw_h, self.raw_s, self.n_dims, self.num_bins

            def bw():  # reverse mode of the spline (vms_rqs_apply_backward overwrites its outputs: temporaries + accumulate)
                has_out, has_l = tp.has(out), (ldj_sum is not None and tp.has(ldj_sum))
                if ldj is not None and tp.has(ldj):
                    raise NotImplementedError('reverse mode through per-dimension log-dets (event_ndims=0) is not built')
                if not (has_out or has_l):
                    return
                g_out = tp.grad(out) if has_out else Tensor.zeros((B, nd))
                g_in, g_w, g_h, g_s = Tensor((B, nd)), Tensor((B, nd * K)), Tensor((B, nd * K)), Tensor((B, nd * (K - 1)))
                ba = _abi.RqsBwdArgs(a, g_out.ptr, g_out.ld, tp.grad(ldj_sum).ptr if has_l else None, g_in.ptr, g_in.ld,
                                     g_w.ptr, g_w.ld, g_h.ptr, g_h.ld, g_s.ptr, g_s.ld)
                c.lib.vms_rqs_apply_backward(C.byref(ba), c.stream)
                ad = _ad()
                ad.add_into(tp.grad(v), g_in, 1.0)
                ad.add_into(tp.grad(rw), g_w, 1.0)
                ad.add_into(tp.grad(rh), g_h, 1.0)
                ad.add_into(tp.grad(rs), g_s, 1.0)

            tp.record(bw)

    def _apply(self, v, inverse):
        out = Tensor(v.shape)
        ldj = Tensor((v.shape[0],))
        self.apply_into(v, out, ldj, False, inverse)
        return out, ldj

    def _fwd(self, x, **kw):
        return self._apply(x, False)

    def _inv(self, y, **kw):
        return self._apply(y, True)

    def forward_log_det_jacobian(self, x, event_ndims=1, **kw):
        if event_ndims == 0:
            x = as_tensor(x)
            out, l = Tensor(x.shape), Tensor(x.shape)
            self.apply_into(x, out, None, False, False, ldj=l)
            return l
        return self._fwd(as_tensor(x))[1]


class RealNVP(Bijector):
    """tfp.bijectors.RealNVP with a spline `bijector_fn`: num_masked > 0 conditions on the first columns, < 0 on the
    last (flows.py:290-306); the conditioner sees the masked columns, the spline transforms the rest."""

    def __init__(self, num_masked, bijector_fn, name='real_nvp'):
        self.num_masked = int(num_masked)
        self.bijector_fn = bijector_fn
        self.name = name

    def _split(self, D):
        m = self.num_masked
        if m >= 0:
            return (0, m), (m, D)
        return (D + m, D), (0, D + m)

    def _apply(self, v, inverse, **kw):
        D = v.shape[1]
        (c0, c1), (t0, t1) = self._split(D)
        cond = v.cols(c0, c1)
        bij = self.bijector_fn(cond, t1 - t0, **kw)
        tp = _tape()
        if tp is None:
            out = v.contig().copy() if c1 > c0 else Tensor(v.shape)
        else:
            vc = v.contig()
            with tp.paused():  # the pass-through copy is differentiated here: only the conditioner columns flow straight back
                out = vc.copy() if c1 > c0 else Tensor(v.shape)
            if c1 > c0:
                def bw():
                    if tp.has(out):
                        _ad().add_into(tp.grad(vc).cols(c0, c1), tp.grad(out).cols(c0, c1), 1.0)

                tp.record(bw)
            v = vc
        ldj = Tensor((v.shape[0],))
        bij.apply_into(v.cols(t0, t1), out.cols(t0, t1), ldj, False, inverse)
        return out, ldj

    def _fwd(self, x, **kw):
        return self._apply(x, False, **kw)

    def _inv(self, y, **kw):
        return self._apply(y, True, **kw)

    def _layers(self):
        return [self.bijector_fn] if isinstance(self.bijector_fn, Layer) else []


class MaskedAutoregressiveFlow(Bijector):
    """tfp.bijectors.MaskedAutoregressiveFlow with a spline `bijector_fn`: inverse (density direction) is ONE
    conditioner pass; forward (sampling) runs event_size passes starting from zeros."""

    def __init__(self, bijector_fn, name='masked_autoregressive_flow'):
        self.bijector_fn = bijector_fn
        self.name = name

    def _inv(self, y, **kw):
        return self.bijector_fn(y, **kw)._apply(y, True)

    def _fwd(self, x, **kw):
        y = Tensor.zeros(x.shape)
        ldj = None
        for _ in range(x.shape[1]):
            y, ldj = self.bijector_fn(y, **kw)._apply(x, False)
        return y, ldj

    def _layers(self):
        return [self.bijector_fn] if isinstance(self.bijector_fn, Layer) else []


class Affine(Bijector):
    """Column-wise affine bijector: y = (x + shift) * scale (shift_first) or y = x * scale + shift."""

    def __init__(self, shift=None, scale=None, shift_first=False, name='affine'):
        self.shift_np = None if shift is None else np.asarray(shift, np.float32).reshape(-1)
        self.scale_np = None if scale is None else np.asarray(scale, np.float32).reshape(-1)
        self.shift_first = shift_first
        self.name = name
        self._dev = None

    def _params(self, inverse):
        if self._dev is None:
            sh, sc = self.shift_np, self.scale_np
            one = np.float32(1)
            fwd = (None if sc is None else Tensor.from_numpy(sc), None if sh is None else Tensor.from_numpy(sh))
            # inverse of y = x*sc + sh : x = (y - sh) / sc = (y + (-sh)) * (1/sc), applied shift-first
            inv = (None if sc is None else Tensor.from_numpy(one / sc), None if sh is None else Tensor.from_numpy(-sh))
            self._dev = (fwd, inv)
        return self._dev[1 if inverse else 0]

    def _run(self, v, inverse):
        c = ctx()
        B, D = v.shape
        sc, sh = self._params(inverse)
        out = Tensor((B, D))
        shift_first = (not self.shift_first) if inverse else self.shift_first
        c.lib.vms_affine_cols(v.ptr, v.ld, B, D, _ptr(sc), _ptr(sh), 1 if shift_first else 0, out.ptr, out.ld, c.stream)
        l = 0.0 if self.scale_np is None else float(np.sum(np.log(np.abs(self.scale_np)), dtype=np.float32))
        ldj = Tensor.from_numpy(np.full(B, -l if inverse else l, np.float32))
        return out, ldj

    def _fwd(self, x, **kw):
        return self._run(x, False)

    def _inv(self, y, **kw):
        return self._run(y, True)


def Shift(shift, name='shift'):
    return Affine(shift=shift, name=name)


def Scale(scale, name='scale'):
    return Affine(scale=scale, name=name)


class Chain(Bijector):
    """tfp.bijectors.Chain: forward applies bijectors[-1] first; inverse applies bijectors[0]^-1 first.
    Keyword arguments keyed by bijector name are routed to that bijector (flows.py:671-690)."""

    def __init__(self, bijectors, name='chain'):
        self.bijectors = list(bijectors)
        self.name = name

    def _walk(self, v, seq, method, kw):
        total = None
        for b in seq:
            v, l = getattr(b, method)(v, **kw.get(b.name, {}))
            total = l if total is None else total + l
        if total is None:
            total = Tensor.zeros((v.shape[0],))
        return v, total

    def _fwd(self, x, **kw):
        return self._walk(x, self.bijectors[::-1], '_fwd', kw)

    def _inv(self, y, **kw):
        return self._walk(y, self.bijectors, '_inv', kw)

    def _layers(self):
        out = []
        for b in self.bijectors:
            out += b._layers()
        return out


class BatchNormalization(Bijector):
    """tfp.bijectors.BatchNormalization(training=...) as placed between flow blocks (flows.py:308-309, :623-624).
    [TFP-recalled] The bijector's INVERSE is the normalisation (batch moments when `training`, which also updates the
    moving statistics; moving statistics otherwise), its forward the de-normalisation with the moving statistics;
    ildj = sum_d log gamma_d - 0.5 log(var_d + eps) for every row, fldj the negative with the moving variance."""

    def __init__(self, training=False, name='batch_normalization'):
        self.training = training
        self.name = name
        self.state = None
        self._holder = _BatchNormWeights(self)

    def _state(self, v):
        if self.state is None:
            self.state = _BatchNormState(v.shape[1])
        return self.state

    def _fwd(self, x, **kw):
        st = self._state(x)
        return st.apply(x, st.moving_mean, st.moving_variance, denormalize=True, want_ldj=True)

    def _inv(self, y, **kw):
        st = self._state(y)
        mean, var = st.statistics(y, self.training)
        return st.apply(y, mean, var, denormalize=False, want_ldj=True)

    def _layers(self):
        return [self._holder]  # gamma / beta are trainable variables of the model, as in tfp


class _BatchNormWeights(Layer):
    """Exposes the variables of a batch-norm bijector (created lazily at its first call) to `Model.weights`."""

    def __init__(self, owner):
        super(_BatchNormWeights, self).__init__(name=owner.name + '_variables')
        self.owner_ref = [owner]  # (a list: keep the bijector out of Layer._sublayers' attribute scan)

    def _sublayers(self):
        return []

    @property
    def weights(self):
        st = self.owner_ref[0].state
        return st.weights() if st is not None else []

    trainable_weights = weights


# ================================================================================================ distributions
class Distribution(object):
    """TFP distribution protocol: sample([n]), log_prob(x), experimental_sample_and_log_prob()."""
    batch = None  # number of batch rows, or None for an unbatched distribution
    event_size = 1

    def _rows(self):
        return 1 if self.batch is None else self.batch

    # subclasses: _sample_rows() -> Tensor [rows, D];  _log_prob_rows(x [rows, D]) -> Tensor [rows]
    def sample(self, sample_shape=None, **kw):
        if sample_shape is None or sample_shape == () or sample_shape == []:
            s = self._sample_rows(**kw)
            return s if self.batch is not None else s.reshape(self.event_size)
        n = int(np.prod(sample_shape))
        draws = np.stack([self._sample_rows(**kw).numpy() for _ in range(n)])
        if self.batch is None:
            draws = draws.reshape(n, self.event_size)
        return Tensor.from_numpy(draws)

    def log_prob(self, x, **kw):
        x = as_tensor(x)
        rows, D = self._rows(), self.event_size
        if x.ndim == 1:
            return self._log_prob_rows(x.reshape(1, D), **kw).reshape(())
        if x.ndim == 2 and x.shape[0] == rows:
            return self._log_prob_rows(x, **kw)
        if x.ndim == 2 and rows == 1:  # unbatched distribution, [n, D] samples
            return self._broadcast_log_prob(x, **kw)
        if x.ndim == 3 and x.shape[1] == rows:
            xs = x.numpy()
            return Tensor.from_numpy(np.stack([self._log_prob_rows(Tensor.from_numpy(xs[i]), **kw).numpy()
                                               for i in range(xs.shape[0])]))
        raise ValueError('log_prob: sample shape %s incompatible with batch %s / event %d' % (x.shape, self.batch, D))

    def _broadcast_log_prob(self, x, **kw):
        xs = x.numpy()
        return Tensor.from_numpy(np.concatenate([self._log_prob_rows(Tensor.from_numpy(xs[i:i + 1]), **kw).numpy()
                                                 for i in range(xs.shape[0])]))

    def experimental_sample_and_log_prob(self, sample_shape=None, **kw):
        s = self.sample(sample_shape, **kw)
        return s, self.log_prob(s, **kw)


class StandardNormal(Distribution):
    """Independent(Normal(zeros, ones)) -- the static latent / base distribution of the reference's tests and
    notebooks (tests/test_models.py:172-175)."""

    def __init__(self, batch, event_size):
        self.batch = None if batch is None else int(batch)
        self.event_size = int(event_size)

    def _sample_rows(self, **kw):
        return Tensor.from_numpy(rng().standard_normal((self._rows(), self.event_size), dtype=np.float32))

    def _log_prob_any(self, x):
        c = ctx()
        lp = Tensor((x.shape[0],))
        c.lib.vms_std_normal_log_prob(x.ptr, x.ld, x.shape[0], self.event_size, lp.ptr, 0, c.stream)
        tp = _tape()
        if tp is not None:
            D = self.event_size

            def bw():
                if tp.has(lp):
                    g_x = tp.grad(x)
                    c.lib.vms_std_normal_log_prob_backward(x.ptr, x.ld, x.shape[0], D, tp.grad(lp).ptr, g_x.ptr, g_x.ld,
                                                           c.stream)

            tp.record(bw)
        return lp

    def _log_prob_rows(self, x, **kw):
        return self._log_prob_any(x)

    def _broadcast_log_prob(self, x, **kw):
        return self._log_prob_any(x)


class Blockwise(Distribution):
    """Per-degree-of-freedom Normal / von Mises distributions over one parameter tensor (vms_blockwise_log_prob)."""

    def __init__(self, params, kinds, loc_off, loc2_off, scale_off, scale_mode):
        self.params = params  # Tensor [B, P]
        self._n = params.shape[0]
        self.batch = self._n
        self.kinds = [int(k) for k in kinds]
        self.event_size = len(self.kinds)
        arr = lambda a: (C.c_int32 * self.event_size)(*[int(v) for v in a])
        self._kind, self._loc, self._loc2, self._scale = arr(kinds), arr(loc_off), arr(loc2_off), arr(scale_off)
        self.loc_off, self.loc2_off, self.scale_off = list(loc_off), list(loc2_off), list(scale_off)
        self.scale_mode = int(scale_mode)

    def _lp_call(self, x, n, ld_p):
        """log_prob of n rows; ld_p = 0 broadcasts the single parameter row of an unbatched distribution to every row."""
        c = ctx()
        lp = Tensor((n,))
        params, D = self.params, self.event_size
        c.lib.vms_blockwise_log_prob(x.ptr, x.ld, params.ptr, ld_p, n, D, self._kind, self._loc, self._loc2, self._scale,
                                     self.scale_mode, lp.ptr, 0, c.stream)
        tp = _tape()
        if tp is not None:
            def bw():  # TF autodiff through Normal / VonMises log_prob and the parameter transforms (dists.py:56-78)
                if not tp.has(lp):
                    return
                g_p = tp.grad(params) if ld_p else None  # (a broadcast parameter row is a constant of the model)
                g_x = tp.grad(x)
                c.lib.vms_blockwise_log_prob_backward(x.ptr, x.ld, params.ptr, ld_p, n, D, self._kind, self._loc, self._loc2,
                                                      self._scale, self.scale_mode, tp.grad(lp).ptr, g_x.ptr, g_x.ld, _ptr(g_p),
                                                      0 if g_p is None else g_p.ld, c.stream)

            tp.record(bw)
        return lp

    def _log_prob_rows(self, x, **kw):
        return self._lp_call(x, self._n, self.params.ld)

    def _broadcast_log_prob(self, x, **kw):
        if self._n != 1:
            return Distribution._broadcast_log_prob(self, x, **kw)
        return self._lp_call(x, x.shape[0], 0)

    def _planar_normal(self):
        D = self.event_size
        return (all(k == DIST_NORMAL for k in self.kinds) and self.loc_off == list(range(self.loc_off[0], self.loc_off[0] + D))
                and self.scale_off == list(range(self.scale_off[0], self.scale_off[0] + D)))

    def constrained_params(self):
        """(loc [B, D], scale-or-concentration [B, D]) after the parameter transforms, computed on the device."""
        c = ctx()
        loc, scale = Tensor((self._n, self.event_size)), Tensor((self._n, self.event_size))
        c.lib.vms_blockwise_params(self.params.ptr, self.params.ld, self._n, self.event_size, self._kind, self._loc,
                                   self._loc2, self._scale, self.scale_mode, loc.ptr, scale.ptr, c.stream)
        return loc, scale

    def sample_with_noise(self, eps, want_log_prob=True):
        """Reparameterised Normal sample z = eps * scale + loc (and its log_prob) for planar all-Normal layouts."""
        c = ctx()
        z = Tensor((self._n, self.event_size))
        lp = Tensor((self._n,)) if want_log_prob else None
        c.lib.vms_normal_sample_log_prob(self.params.ptr, self.params.ld, self.loc_off[0], self.scale_off[0],
                                         self.scale_mode, eps.ptr, self._n, self.event_size, z.ptr, z.ld, _ptr(lp),
                                         c.stream)
        self._record_sample(z)
        if lp is not None and _tape() is not None:
            raise NotImplementedError('reverse mode through the fused sample + log_prob call is not built: use sample() and '
                                      'log_prob() separately under a tape')
        return z, lp

    def _record_sample(self, z):
        """Reparameterised sample z: pathwise gradient for Normal dofs and for the von Mises location, tfp's implicit
        reparameterisation for the von Mises concentration (vms_blockwise_sample_backward)."""
        tp = _tape()
        if tp is None:
            return
        c = ctx()
        params = self.params

        def bw():
            if not tp.has(z):
                return
            g_z, g_p = tp.grad(z), tp.grad(params)
            c.lib.vms_blockwise_sample_backward(params.ptr, params.ld, self._n, self.event_size, self._kind, self._loc,
                                                self._loc2, self._scale, self.scale_mode, z.ptr, z.ld, g_z.ptr, g_z.ld,
                                                g_p.ptr, g_p.ld, c.stream)

        tp.record(bw)

    def _sample_rows(self, eps=None, **kw):
        if self._planar_normal():
            if eps is None:
                eps = Tensor.from_numpy(rng().standard_normal((self._n, self.event_size), dtype=np.float32))
            return self.sample_with_noise(as_tensor(eps), want_log_prob=False)[0]
        # interleaved layouts / von Mises dofs: one device kernel for every dof (vms_blockwise_sample: Normal from eps or
        # Philox, von Mises by tfp's Best-Fisher rejection sampler); the host only draws the 64-bit stream seed
        c = ctx()
        out = Tensor((self._n, self.event_size))
        e = None if eps is None else as_tensor(eps)
        # (all-Normal dofs with given noise use no device stream: the host generator is left where the caller's noise ended)
        needs_stream = e is None or any(k != DIST_NORMAL for k in self.kinds)
        seed = int(rng().integers(0, 2**63 - 1)) if needs_stream else 0
        c.lib.vms_blockwise_sample(self.params.ptr, self.params.ld, self._n, self.event_size, self._kind, self._loc,
                                   self._loc2, self._scale, self.scale_mode, _ptr(e), 0 if e is None else e.ld, seed,
                                   out.ptr, out.ld, c.stream)
        self._record_sample(out)
        return out

    def experimental_sample_and_log_prob(self, sample_shape=None, **kw):
        if sample_shape is None and self._planar_normal():
            eps = Tensor.from_numpy(rng().standard_normal((self._n, self.event_size), dtype=np.float32))
            return self.sample_with_noise(eps)
        return Distribution.experimental_sample_and_log_prob(self, sample_shape, **kw)


def _as_rows(a):
    a = np.asarray(a.numpy() if isinstance(a, Tensor) else a, np.float32)
    return a.reshape(1, -1) if a.ndim <= 1 else a.reshape(a.shape[0], -1)


class Normal(Blockwise):
    """tfp.distributions.Normal(loc, scale) reinterpreted as Independent over the last axis (conftest.py:11-22).
    Also the `dist_classes` marker for IndependentBlockwise (dists.py:116-196)."""
    kind = DIST_NORMAL
    num_params = 2  # preferred parameters: loc, scale

    def __init__(self, loc, scale):
        unbatched = np.ndim(loc.numpy() if isinstance(loc, Tensor) else loc) <= 1
        loc, scale = np.broadcast_arrays(_as_rows(loc), _as_rows(scale))
        D = loc.shape[1]
        Blockwise.__init__(self, Tensor.from_numpy(np.concatenate([loc, scale], axis=1)), [self.kind] * D, range(D),
                           [-1] * D, range(D, 2 * D), SCALE_IDENTITY)
        if unbatched:
            self.batch = None


class VonMises(Normal):
    """tfp.distributions.VonMises(loc, concentration), Independent over the last axis (conftest.py:25-29)."""
    kind = DIST_VONMISES
    num_params = 2  # loc, concentration (+1 raw parameter for the sine / cosine pair, dists.py:171-172)

    def __init__(self, loc, concentration):
        Normal.__init__(self, loc, concentration)


class Deterministic(Distribution):
    """Independent(Deterministic(loc)) (dists.py:701-704): sampling returns loc; log_prob is 0 at loc, -inf elsewhere."""

    def __init__(self, loc):
        self.loc = loc
        self.batch, self.event_size = loc.shape[0], loc.shape[1]

    def _sample_rows(self, **kw):
        return self.loc.copy()

    def _log_prob_rows(self, x, **kw):
        c = ctx()
        x = as_tensor(x)
        lp = Tensor((self.batch,))
        c.lib.vms_deterministic_log_prob(x.ptr, x.ld, self.loc.ptr, self.loc.ld, self.batch, self.event_size, lp.ptr,
                                         c.stream)
        return lp


class TransformedDistribution(Distribution):
    """tfp.distributions.TransformedDistribution: log_prob(y) = base.log_prob(inv(y)) + ildj(y);
    sample = fwd(base.sample()); experimental_sample_and_log_prob = (fwd(x), base_lp(x) - fldj(x))."""

    def __init__(self, distribution, bijector, kwargs_split_fn=None, name='transformed_distribution'):
        self.distribution = distribution
        self.bijector = bijector
        self.batch, self.event_size = distribution.batch, distribution.event_size
        self._split = kwargs_split_fn or (lambda kw: (kw.get('distribution_kwargs', {}), kw.get('bijector_kwargs', {})))
        self.name = name

    def _log_prob_rows(self, y, **kw):
        dkw, bkw = self._split(kw)
        x, ildj = self.bijector._inv(y, **bkw)
        return self.distribution.log_prob(x, **dkw) + ildj

    def _broadcast_log_prob(self, y, **kw):
        dkw, bkw = self._split(kw)
        x, ildj = self.bijector._inv(y, **bkw)
        return self.distribution.log_prob(x, **dkw) + ildj

    def _sample_rows(self, **kw):
        dkw, bkw = self._split(kw)
        x = self.distribution._sample_rows(**dkw)
        return self.bijector._fwd(x, **bkw)[0]

    def sample(self, sample_shape=None, **kw):
        if sample_shape is None or self.batch is not None:
            return Distribution.sample(self, sample_shape, **kw)
        dkw, bkw = self._split(kw)  # unbatched base: draw [n, D] rows and push them through in one pass
        x = self.distribution.sample(sample_shape, **dkw)
        return self.bijector._fwd(x, **bkw)[0]

    def experimental_sample_and_log_prob(self, sample_shape=None, **kw):
        dkw, bkw = self._split(kw)
        x, lp0 = self.distribution.experimental_sample_and_log_prob(sample_shape, **dkw)
        if x.ndim != 2:
            return Distribution.experimental_sample_and_log_prob(self, sample_shape, **kw)
        y, fldj = self.bijector._fwd(x, **bkw)
        return y, lp0 - fldj


class Autoregressive(Distribution):
    """tfp.distributions.Autoregressive (dists.py:338-340): log_prob(x) = distribution_fn(x).log_prob(x) (one pass);
    sampling = num_steps + 1 passes from sample0 with the SAME noise at every pass."""

    def __init__(self, distribution_fn, sample0, num_steps):
        self.distribution_fn = distribution_fn
        self.sample0 = sample0
        self.num_steps = int(num_steps)
        self.batch, self.event_size = sample0.shape[0], sample0.shape[1]

    def _log_prob_rows(self, x, **kw):
        return self.distribution_fn(x).log_prob(x)

    def _sample_rows(self, eps=None, **kw):
        # tfp re-uses ONE seed for every pass, i.e. the same noise per dof at every step: the Normal noise is drawn once on
        # the host generator ([B, D] float32, reproducible by the oracle) and handed to every pass; the generator state is
        # restored before each pass so that von Mises dofs (device Philox stream, seeded from it) repeat their noise too
        if eps is None:
            eps = Tensor.from_numpy(rng().standard_normal((self.batch, self.event_size), dtype=np.float32))
        state = rng().bit_generator.state
        s = self.sample0
        for _ in range(self.num_steps + 1):
            rng().bit_generator.state = state
            s = self.distribution_fn(s)._sample_rows(eps=eps)
        return s


class DistributionLambda(Layer):
    """tfp.layers.DistributionLambda: a layer whose call builds a distribution from its input."""

    def __init__(self, make_distribution_fn, name='distribution_lambda'):
        super(DistributionLambda, self).__init__(name=name)
        self.make_distribution_fn = make_distribution_fn

    def call(self, inputs, **kw):
        return self.make_distribution_fn(inputs)


class IndependentNormal(DistributionLambda):
    """tfp.layers.IndependentNormal(event_shape): params [B, 2 D] = [loc | raw], scale = softplus(raw)
    (tests/test_models.py:167-170, MC notebook cell 11)."""

    def __init__(self, event_shape=(), name='independent_normal'):
        self.event_size = int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape)
        super(IndependentNormal, self).__init__(self.new, name=name)

    def new(self, params):
        params = as_tensor(params)
        D = self.event_size
        if params.shape[-1] != 2 * D:
            raise ValueError('IndependentNormal(%d) needs %d parameters, got %d' % (D, 2 * D, params.shape[-1]))
        return Blockwise(params, [DIST_NORMAL] * D, range(D), [-1] * D, range(D, 2 * D), SCALE_SOFTPLUS)

    @staticmethod
    def params_size(event_shape=()):
        return 2 * (int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape))
