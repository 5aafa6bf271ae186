"""Distribution-producing layers -- host-side mirror of `vaemolsim/dists.py` over sm_100a kernels.

Same public names / keywords / error behaviour as the reference (dists.py:28 `make_param_transform`, :97
`IndependentBlockwise`, :246 `AutoregressiveBlockwise`, :369 `FlowedDistribution`, :478 `StaticFlowedDistribution`,
:545 `IndependentVonMises`, :642 `IndependentDeterministic`).  The layers return distribution objects whose
`log_prob` / `sample` / `experimental_sample_and_log_prob` launch `csrc/logprob.cu` kernels; parameter transforms
(softplus, softplus + eps32, atan2) are fused into those kernels.  `JointDistribution` (dists.py:755, WIP in the
reference, references undefined names) is out of scope.
"""
import numpy as np

from . import _protocols as P
from ._abi import Tensor, as_tensor
from ._protocols import (Autoregressive, Blockwise, Deterministic, Distribution, DistributionLambda, IndependentNormal,
                         Normal, StandardNormal, TransformedDistribution, VonMises)

__all__ = ['make_param_transform', 'IndependentBlockwise', 'AutoregressiveBlockwise', 'FlowedDistribution',
           'StaticFlowedDistribution', 'IndependentVonMises', 'IndependentDeterministic', 'Normal', 'VonMises',
           'StandardNormal', 'IndependentNormal', 'DistributionLambda', 'Distribution', 'Blockwise', 'Autoregressive',
           'Deterministic', 'TransformedDistribution']


def _identity(x):
    return x


def make_param_transform(dist_class=None, transform_fn=_identity):
    """dists.py:28-87.  Returns a function mapping raw parameters `x[..., i]` to a dict of constrained parameters.

    Normal: `parameter_properties` bijectors => loc identity, scale Softplus(low=eps32).  VonMises: loc = atan2(x0, x1),
    concentration = SoftClip(low=eps32, high=sqrt(max32)/2)(x2), evaluated as softplus(x2) + eps32 (the soft upper clip
    at 9.2e18 is unreachable in float32 practice).  The transform runs on the device (vms_blockwise_params).
    """
    if dist_class is None:
        return transform_fn
    if not (isinstance(dist_class, type) and issubclass(dist_class, Normal)):
        raise NotImplementedError('make_param_transform: only Normal and VonMises are implemented, got %r' % (dist_class,))
    vm = dist_class.__name__ == 'VonMises'

    def _fn(x):
        x = as_tensor(x)
        lead = x.shape[:-1]
        flat = x.reshape(-1, x.shape[-1]) if x.ndim != 2 else x
        if vm:
            d = Blockwise(flat, [P.DIST_VONMISES], [0], [1], [2], P.SCALE_SOFTPLUS_EPS)
        else:
            d = Blockwise(flat, [P.DIST_NORMAL], [0], [-1], [1], P.SCALE_SOFTPLUS_EPS)
        loc, scale = d.constrained_params()
        return {'loc': loc.reshape(lead), ('concentration' if vm else 'scale'): scale.reshape(lead)}

    _fn.dist_kind = P.DIST_VONMISES if vm else P.DIST_NORMAL
    return _fn


class IndependentBlockwise(P.Layer):
    """dists.py:97-243: one independent distribution per degree of freedom, parameters split contiguously."""

    def __init__(self, num_dofs, dist_classes, param_nums=None, param_transforms=None, name='independent_blockwise',
                 **kwargs):
        super(IndependentBlockwise, self).__init__(name=name, **kwargs)
        self.num_dofs = num_dofs
        if not isinstance(dist_classes, (list, tuple)):
            if not (isinstance(dist_classes, type) and issubclass(dist_classes, Distribution)):
                raise TypeError("Expected a distribution class object, but got %s" % type(dist_classes).__name__)
            self.dist_classes = [dist_classes] * self.num_dofs
        else:
            if len(dist_classes) != self.num_dofs:
                raise ValueError("If specifying a list of distribution classes, must be of same length as number of "
                                 "degrees of freedom (%i) but got %i." % (self.num_dofs, len(dist_classes)))
            self.dist_classes = dist_classes
        if param_nums is None:
            # preferred parameters of the class, +1 for von Mises (sine / cosine pair for the location)
            self.param_nums = [d.num_params + (1 if d.__name__ == 'VonMises' else 0) for d in self.dist_classes]
        elif not isinstance(param_nums, (list, tuple)):
            self.param_nums = [param_nums] * self.num_dofs
        else:
            if len(param_nums) != self.num_dofs:
                raise ValueError("If specifying a list of parameter numbers, must be of same length as number of "
                                 "degrees of freedom (%i) but got %i." % (self.num_dofs, len(param_nums)))
            self.param_nums = param_nums
        if param_transforms is None:
            self.param_transforms = [make_param_transform(dist_class=d) for d in self.dist_classes]
        elif not isinstance(param_transforms, (list, tuple)):
            self.param_transforms = [_identity] * self.num_dofs
        else:
            if len(param_transforms) != self.num_dofs:
                raise ValueError("If specifying a list of parameter transformations, must be of same length as number "
                                 "of degrees of freedom (%i) but got %i." % (self.num_dofs, len(param_transforms)))
            self.param_transforms = param_transforms
        for d in self.dist_classes:
            if not (isinstance(d, type) and issubclass(d, Normal)):
                raise NotImplementedError('only Normal and VonMises degrees of freedom have device kernels, got %r' % (d,))

    def _layout(self, starts):
        """kinds / column offsets of every dof given the first column of its parameter group."""
        kinds, loc, loc2, scale = [], [], [], []
        mode = None
        for i, d in enumerate(self.dist_classes):
            vm = d.__name__ == 'VonMises'
            tf = self.param_transforms[i]
            ident = tf is _identity
            # the reference calls self.param_transforms[i](params[i]) (dists.py:213); here the transform is fused into the
            # kernels, so only the two transforms that HAVE a kernel are accepted -- anything else must fail, not be
            # silently replaced by the default
            want_kind = P.DIST_VONMISES if vm else P.DIST_NORMAL
            if not ident and getattr(tf, 'dist_kind', None) != want_kind:
                raise NotImplementedError('param_transforms[%d]: only the identity and make_param_transform(%s) have device '
                                          'kernels; got %r' % (i, d.__name__, tf))
            this_mode = P.SCALE_IDENTITY if ident else P.SCALE_SOFTPLUS_EPS
            if mode is not None and this_mode != mode:
                raise NotImplementedError('mixing identity and default parameter transforms is not supported')
            mode = this_mode
            kinds.append(P.DIST_VONMISES if vm else P.DIST_NORMAL)
            s = starts[i]
            if vm and not ident:
                loc.append(s), loc2.append(s + 1), scale.append(s + 2)
            else:
                loc.append(s), loc2.append(-1), scale.append(s + 1)
        return kinds, loc, loc2, scale, mode

    def call(self, inputs):
        params = as_tensor(inputs)
        if params.ndim != 2:
            params = params.reshape(params.shape[0], -1)
        starts = np.concatenate([[0], np.cumsum(self.param_nums)[:-1]]).astype(int)
        kinds, loc, loc2, scale, mode = self._layout(starts)
        return Blockwise(params, kinds, loc, loc2, scale, mode)

    def params_size(self):
        return sum(self.param_nums)

    def get_config(self):
        config = super(IndependentBlockwise, self).get_config()
        config.update({"num_dofs": self.num_dofs, "dist_classes": self.dist_classes, "param_nums": self.param_nums,
                       "param_transforms": self.param_transforms})
        return config


class AutoregressiveBlockwise(IndependentBlockwise):
    """dists.py:246-366: Autoregressive distribution over a Blockwise one; a MADE network shifts the raw parameters."""

    def __init__(self, *args, conditional=False, conditional_event_shape=None, auto_net_params={},
                 name='autoregressive_blockwise', **kwargs):
        super(AutoregressiveBlockwise, self).__init__(*args, name=name, **kwargs)
        self.conditional = conditional
        self.conditional_event_shape = conditional_event_shape
        self.auto_net_params = auto_net_params

    def build(self, input_shape):
        if tuple(input_shape[-2:]) != (self.num_dofs, max(self.param_nums)):
            raise ValueError("Last (assuming only non-batch) dimension is of size %s, but must match number of "
                             "specified degrees of freedom, %s." %
                             (str(tuple(input_shape[-2:])), str((self.num_dofs, max(self.param_nums)))))
        self.auto_net = P.AutoregressiveNetwork(max(self.param_nums), self.num_dofs, conditional=self.conditional,
                                                conditional_event_shape=self.conditional_event_shape,
                                                **self.auto_net_params)
        self.auto_net.build((None, self.num_dofs))  # Keras builds sub-layers on first use; weights exist from here on
        self.auto_net.built = True

    def call(self, inputs, conditional_input=None):
        inputs = as_tensor(inputs)
        B = inputs.shape[0]
        pmax = max(self.param_nums)
        flat_in = inputs.reshape(B, self.num_dofs * pmax)
        kinds, loc, loc2, scale, mode = self._layout([i * pmax for i in range(self.num_dofs)])
        if self.conditional and conditional_input is None:
            raise ValueError('`conditional_input` must be passed as a named argument.')
        cond = as_tensor(conditional_input) if self.conditional else None

        def _make_dist(samples):
            shift = self.auto_net(as_tensor(samples).contig(), conditional_input=cond)
            raw_params = flat_in + shift.reshape(B, self.num_dofs * pmax)
            return Blockwise(raw_params, kinds, loc, loc2, scale, mode)

        cache = self.__dict__.setdefault('_ones', {})  # (a device constant per batch size: no upload inside a training step,
        sample0 = cache.get(B)                         # which keeps the step capturable in a CUDA graph)
        if sample0 is None:
            sample0 = cache[B] = Tensor.from_numpy(np.ones((B, self.num_dofs), np.float32))
        return Autoregressive(_make_dist, sample0=sample0, num_steps=self.num_dofs)

    def params_size(self):
        return (self.num_dofs, max(self.param_nums))

    def get_config(self):
        config = super(AutoregressiveBlockwise, self).get_config()
        config.update({"conditional": self.conditional, "conditional_event_shape": self.conditional_event_shape,
                       "auto_net_params": self.auto_net_params})
        return config


class FlowedDistribution(P.Layer):
    """dists.py:369-475: latent distribution layer followed by a flow => TransformedDistribution."""

    def __init__(self, flow, latent_dist, name='flowed_dist', **kwargs):
        super(FlowedDistribution, self).__init__(name=name, **kwargs)
        self.flow = flow
        self.latent_dist = latent_dist
        self.conditional = self.flow.conditional

    def call(self, inputs, training=False, **kwargs):
        start_dist = self.latent_dist(inputs)
        return self.flow(start_dist, training=training, **kwargs)

    def params_size(self):
        if isinstance(self.latent_dist, DistributionLambda) and hasattr(self.latent_dist, 'event_size'):
            return self.latent_dist.params_size(self.latent_dist.event_size)
        return self.latent_dist.params_size()

    def get_config(self):
        config = super(FlowedDistribution, self).get_config()
        config.update({"flow": self.flow, "latent_dist": self.latent_dist})
        return config


class StaticFlowedDistribution(P.Layer):
    """dists.py:478-542: a static (input-independent) latent distribution transformed by a flow."""

    def __init__(self, flow, latent_dist, name='static_flowed_dist', **kwargs):
        super(StaticFlowedDistribution, self).__init__(name=name, **kwargs)
        self.flow = flow
        self.latent_dist = latent_dist

    def __call__(self, inputs, training=False):
        return self.flow(self.latent_dist, training=training)

    def get_config(self):
        config = super(StaticFlowedDistribution, self).get_config()
        config.update({"flow": self.flow, "latent_dist": self.latent_dist})
        return config


class IndependentVonMises(DistributionLambda):
    """dists.py:545-639: params [B, 3 D] = [sine | cosine | raw concentration]; loc = atan2, concentration = softplus."""

    def __init__(self, event_shape=(), name='independent_von_mises', **kwargs):
        self.event_size = int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape)
        self._event_shape = event_shape
        super(IndependentVonMises, self).__init__(lambda t: IndependentVonMises.new(t, event_shape), name=name, **kwargs)

    @staticmethod
    def new(params, event_shape=(), validate_args=False, name=None):
        params = as_tensor(params)
        D = int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape)
        if params.shape[-1] != 3 * D:
            raise ValueError('IndependentVonMises(%d) needs %d parameters, got %d' % (D, 3 * D, params.shape[-1]))
        return Blockwise(params, [P.DIST_VONMISES] * D, range(D), range(D, 2 * D), range(2 * D, 3 * D), P.SCALE_SOFTPLUS)

    @staticmethod
    def params_size(event_shape=(), name=None):
        return np.int32(3) * (int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape))


class IndependentDeterministic(DistributionLambda):
    """dists.py:642-733: a 'distribution' that deterministically returns its parameters."""

    def __init__(self, event_shape=(), name='independent_deterministic', **kwargs):
        self.event_size = int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape)
        self._event_shape = event_shape
        super(IndependentDeterministic, self).__init__(lambda t: IndependentDeterministic.new(t, event_shape), name=name,
                                                       **kwargs)

    @staticmethod
    def new(params, event_shape=(), validate_args=False, name=None):
        params = as_tensor(params)
        return Deterministic(params if params.ndim == 2 else params.reshape(params.shape[0], -1))

    @staticmethod
    def params_size(event_shape=(), name=None):
        return np.int32(1) * (int(np.prod(event_shape)) if np.ndim(event_shape) else int(event_shape))
