"""Normalizing flows -- host-side mirror of `vaemolsim/flows.py` over sm_100a kernels.

Same public names, constructor keywords, call signatures and error behaviour as the reference module
(flows.py:15 `make_domain_transform`, :63 `SplineBijector`, :221 `RQSSplineRealNVP`, :363 `MaskedSplineBijector`,
:531 `RQSSplineMAF`).  Layers return bijector / distribution objects speaking the TFP protocols (`_protocols.py`);
the spline arithmetic (activations + bin search + RQS + log-det) runs in `csrc/rqs.cu`, conditioners in
`csrc/dense.cu`.
"""
import numpy as np

from . import _protocols as P
from ._abi import Tensor, as_tensor


def make_domain_transform(domain_list, target, from_target=False):
    """flows.py:15-60: chained Shift -> Scale -> Shift bijector mapping each domain onto `target` (or back)."""
    t_l = target[1] - target[0]
    t_mean = 0.5 * (target[1] + target[0])
    d_l = np.array([(b - a) for a, b in domain_list], dtype='float32')
    d_mean = np.array([0.5 * (a + b) for a, b in domain_list], dtype='float32')
    if from_target:
        shift1 = -t_mean * np.ones_like(d_mean)
        scale = d_l / t_l
        shift2 = d_mean
    else:
        shift1 = -d_mean
        scale = t_l / d_l
        shift2 = t_mean * np.ones_like(d_mean)
    # Chain applies its list in reverse order: shift1 first, then scale, then shift2
    return P.Chain(bijectors=[P.Shift(shift2, name='shift2'), P.Scale(scale, name='scale'), P.Shift(shift1, name='shift1')])


class _HeadView(object):
    """One of the three Dense heads of `SplineBijector` (flows.py:140-152), stored as a column block of the fused
    heads layer so that ONE GEMM produces all raw spline parameters.  `.weights` / `.get_weights()` expose the
    reference's per-head kernel [hidden, cols] and bias [cols]."""

    def __init__(self, owner, name, c0, c1):
        self._owner, self.name, self._c0, self._c1 = owner, name, c0, c1
        self.units = c1 - c0

    def get_weights(self):
        W, b = self._owner.heads.get_weights()
        return [W[:, self._c0:self._c1].copy(), b[self._c0:self._c1].copy()]

    @property
    def weights(self):
        return [Tensor.from_numpy(w) for w in self.get_weights()]

    def set_weights(self, arrays):
        W, b = self._owner.heads.get_weights()
        W[:, self._c0:self._c1] = np.asarray(arrays[0], np.float32)
        b[self._c0:self._c1] = np.asarray(arrays[1], np.float32)
        self._owner.heads.assign(W, b)


class SplineBijector(P.Layer):
    """flows.py:63-218.  Dense(tanh) conditioner + three linear heads -> RationalQuadraticSpline bijector.

    The three heads (`bin_widths`, `bin_heights`, `knot_slopes`) keep their reference shapes but live as column blocks
    [w | h | s] of one fused Dense layer: one GEMM writes the raw-parameter buffer the RQS kernel reads in place.
    """

    def __init__(self, data_dim, name='rqs', bin_range=[-10.0, 10.0], num_bins=32, hidden_dim=200,
                 kernel_initializer='truncated_normal', **kwargs):
        super(SplineBijector, self).__init__(name=name, **kwargs)
        self.data_dim = data_dim
        self.bin_min = bin_range[0]
        self.bin_max = bin_range[1]
        self.num_bins = num_bins
        self.hidden_dim = hidden_dim
        self.kernel_initializer = kernel_initializer

    def build(self, input_shape):
        ki = self.kernel_initializer
        nw = self.data_dim * self.num_bins
        ns = self.data_dim * (self.num_bins - 1)
        self.d1 = P.Dense(self.hidden_dim, name='d1', activation='tanh', kernel_initializer=ki)
        self.heads = P.Dense(2 * nw + ns, activation=None, name='heads', kernel_initializer=ki)
        din = max(int(input_shape[-1]), 1)  # empty conditioner input -> ones((B, 1)), flows.py:184-185
        self.d1.build((None, din))
        self.heads.build((None, self.hidden_dim))
        self.d1.built = self.heads.built = True
        self.bin_widths = _HeadView(self, 'w', 0, nw)
        self.bin_heights = _HeadView(self, 'h', nw, 2 * nw)
        self.knot_slopes = _HeadView(self, 's', 2 * nw, 2 * nw + ns)

    def call(self, input_tensor, nunits=None):
        del nunits  # nets are created beforehand (flows.py:172)
        x = as_tensor(input_tensor)
        if x.ndim <= 1:
            x = x.reshape(1, -1)
        d1_out = self.d1.call(x, ones_input=(x.shape[-1] == 0))
        raw = self.heads.call(d1_out)
        nw = self.data_dim * self.num_bins
        return P.RationalQuadraticSpline(raw.cols(0, nw), raw.cols(nw, 2 * nw), raw.cols(2 * nw, raw.shape[1]),
                                         self.data_dim, self.num_bins, self.bin_min, self.bin_max)

    def get_config(self):
        config = super(SplineBijector, self).get_config()
        config.update({"data_dim": self.data_dim, "bin_range": [self.bin_min, self.bin_max],
                       "num_bins": self.num_bins, "hidden_dim": self.hidden_dim,
                       "kernel_initializer": self.kernel_initializer})
        return config


class RQSSplineRealNVP(P.Layer):
    """flows.py:221-360: chain of RealNVP blocks with rational-quadratic-spline couplings."""

    def __init__(self, num_blocks=4, rqs_params={}, batch_norm=False, before_flow_transform=None,
                 after_flow_transform=None, name='rqs_realNVP', **kwargs):
        super(RQSSplineRealNVP, self).__init__(name=name, **kwargs)
        self.num_blocks = num_blocks
        self.rqs_params = rqs_params
        self.batch_norm = batch_norm
        self.conditional = False
        self.before_flow_transform = before_flow_transform
        self.after_flow_transform = after_flow_transform

    def build(self, input_shape):
        self.data_dim = input_shape[-1]
        block_list = []
        if self.before_flow_transform is not None:
            block_list.append(self.before_flow_transform)
        for i in range(self.num_blocks):
            if self.data_dim == 1:
                this_mask = 0
                num_transform = 1
            elif i % 2 == 0:
                this_mask = self.data_dim // 2
                num_transform = self.data_dim - self.data_dim // 2
            else:
                this_mask = -(self.data_dim - self.data_dim // 2)
                num_transform = self.data_dim // 2
            if i != 0 and self.batch_norm:
                block_list.append(P.BatchNormalization(training=False))
            block_list.append(
                P.RealNVP(num_masked=this_mask, name='block_%i' % i,
                          bijector_fn=SplineBijector(num_transform, **self.rqs_params)))
        if self.after_flow_transform is not None:
            block_list.append(self.after_flow_transform)
        self.chain = P.Chain(block_list[::-1])  # Chain operates in reverse order (flows.py:321-323)

    def call(self, inputs, training=False):
        if self.batch_norm:
            for bij in self.chain.bijectors:
                if isinstance(bij, P.BatchNormalization):
                    bij.training = training
        if isinstance(inputs, P.Distribution):
            return P.TransformedDistribution(inputs, self.chain)
        return self.chain(as_tensor(inputs))

    def get_config(self):
        config = super(RQSSplineRealNVP, self).get_config()
        config.update({"num_blocks": self.num_blocks, "rqs_params": self.rqs_params, "batch_norm": self.batch_norm})
        return config


class MaskedSplineBijector(P.Layer):
    """flows.py:363-528: three MADE networks (K, K, K-1 parameters per dof) -> RationalQuadraticSpline bijector."""

    def __init__(self, name='rqs', bin_range=[-10.0, 10.0], num_bins=32, hidden_dim=200,
                 kernel_initializer='truncated_normal', conditional=False, conditional_event_shape=None,
                 input_order='left-to-right', **kwargs):
        super(MaskedSplineBijector, self).__init__(name=name, **kwargs)
        self.bin_min = bin_range[0]
        self.bin_max = bin_range[1]
        self.num_bins = num_bins
        self.hidden_dim = hidden_dim
        self.kernel_initializer = kernel_initializer
        self.conditional = conditional
        self.conditional_event_shape = conditional_event_shape
        self.input_order = input_order

    def build(self, input_shape):
        self.data_dim = input_shape[-1]
        mk = lambda params, nm: P.AutoregressiveNetwork(
            params, event_shape=self.data_dim, conditional=self.conditional,
            conditional_event_shape=self.conditional_event_shape, input_order=self.input_order,
            hidden_units=[self.hidden_dim], activation='tanh', name=nm, kernel_initializer=self.kernel_initializer)
        self.bin_widths = mk(self.num_bins, 'w')
        self.bin_heights = mk(self.num_bins, 'h')
        self.knot_slopes = mk(self.num_bins - 1, 's')

    def call(self, input_tensor, conditional_input=None):
        x = as_tensor(input_tensor).contig()
        B, D = x.shape
        bw = self.bin_widths(x, conditional_input=conditional_input).reshape(B, D * self.num_bins)
        bh = self.bin_heights(x, conditional_input=conditional_input).reshape(B, D * self.num_bins)
        ks = self.knot_slopes(x, conditional_input=conditional_input).reshape(B, D * (self.num_bins - 1))
        return P.RationalQuadraticSpline(bw, bh, ks, D, self.num_bins, self.bin_min, self.bin_max)

    def get_config(self):
        config = super(MaskedSplineBijector, self).get_config()
        config.update({"bin_range": [self.bin_min, self.bin_max], "num_bins": self.num_bins,
                       "hidden_dim": self.hidden_dim, "kernel_initializer": self.kernel_initializer,
                       "conditional": self.conditional, "conditional_event_shape": self.conditional_event_shape,
                       "input_order": self.input_order})
        return config


class RQSSplineMAF(P.Layer):
    """flows.py:531-700: chain of masked-autoregressive-flow blocks with rational-quadratic-spline transforms."""

    def __init__(self, num_blocks=2, order_seed=None, rqs_params={}, batch_norm=False, before_flow_transform=None,
                 after_flow_transform=None, name='rqs_MAF', **kwargs):
        super(RQSSplineMAF, self).__init__(name=name, **kwargs)
        self.num_blocks = num_blocks
        self.order_seed = order_seed
        self.rqs_params = rqs_params
        self.batch_norm = batch_norm
        self.conditional = rqs_params.get('conditional', False)
        self.before_flow_transform = before_flow_transform
        self.after_flow_transform = after_flow_transform

    def build(self, input_shape):
        self.data_dim = input_shape[-1]
        block_list = []
        if self.before_flow_transform is not None:
            block_list.append(self.before_flow_transform)
        rng = np.random.default_rng(self.order_seed)
        for i in range(self.num_blocks):
            if i == 0:
                order = 'right-to-left'
            elif i == (self.num_blocks - 1):
                order = 'left-to-right'
            else:
                order = np.arange(start=1, stop=self.data_dim + 1)
                rng.shuffle(order)
            if i != 0 and self.batch_norm:
                block_list.append(P.BatchNormalization(training=False))
            if "input_order" in self.rqs_params:
                fn = MaskedSplineBijector(**self.rqs_params)
            else:
                fn = MaskedSplineBijector(input_order=order, **self.rqs_params)
            block_list.append(P.MaskedAutoregressiveFlow(bijector_fn=fn, name='block_%i' % i))
        if self.after_flow_transform is not None:
            block_list.append(self.after_flow_transform)
        self.chain = P.Chain(block_list[::-1])

    def call(self, inputs, training=False, conditional_input=None):
        cond_dict = {}
        for bij in self.chain.bijectors:
            if isinstance(bij, P.MaskedAutoregressiveFlow):
                cond_dict[bij.name] = {'conditional_input': conditional_input}
            elif isinstance(bij, P.BatchNormalization):
                bij.training = training
        if isinstance(inputs, P.Distribution):
            return P.TransformedDistribution(
                inputs, self.chain,
                kwargs_split_fn=lambda kwargs: (kwargs.get('distribution_kwargs', {}),
                                                kwargs.get('bijector_kwargs', cond_dict)))
        return self.chain(as_tensor(inputs), **cond_dict)

    def get_config(self):
        config = super(RQSSplineMAF, self).get_config()
        config.update({"num_blocks": self.num_blocks, "order_seed": self.order_seed, "rqs_params": self.rqs_params,
                       "batch_norm": self.batch_norm})
        return config
