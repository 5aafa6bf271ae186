"""Losses and information regularisers -- host-side mirror of `vaemolsim/losses.py` over sm_100a kernels.

Same names, keywords and error behaviour as the reference (losses.py:26 `LogProbLoss`, :69
`PotentialEnergyLogProbLoss`, :128 `InfoRegularizer`, :201 `NonRegularizer`, :226 `KLDivergenceEstimate`, :256
`LogProbRegularizer`, :299 `ReverseKLDivergenceEstimate`).  Batch means are the deterministic warp-shuffle reductions
of `csrc/reduce.cu` (`vms_kl_mean`, `vms_scaled_mean`); results are 0-d device tensors with `.numpy()`.
"""
import numpy as np

from ._abi import Tensor, as_tensor, ctx


def _mean_diff(a, b, sign=1.0):
    """sign * mean(a - b) over the batch (b may be None)."""
    c = ctx()
    a = a.contig()
    out = Tensor(())
    if b is not None:
        b = b.contig()
    if b is None:
        c.lib.vms_scaled_mean(a.ptr, a.size, sign, out.ptr, c.stream)
    else:
        c.lib.vms_kl_mean(a.ptr, b.ptr, a.size, sign, out.ptr, c.stream)
    from . import _autodiff
    tp = _autodiff.Tape.active()
    if tp is not None:
        n = a.size

        def bw():  # d mean / d a_i = sign / n (Keras mean reduction, losses.py:58, :253)
            if not tp.has(out):
                return
            g = tp.grad(out)
            c.lib.vms_add_scalar(tp.grad(a).ptr, n, g.ptr, float(sign) / n, c.stream)
            if b is not None:
                c.lib.vms_add_scalar(tp.grad(b).ptr, n, g.ptr, -float(sign) / n, c.stream)

        tp.record(bw)
    return out


class Loss(object):
    """Minimal tf.keras.losses.Loss: `__call__(y_true, y_pred)` applies the batch-mean reduction unless
    reduction='none' (tests/test_losses.py:29-36)."""

    def __init__(self, name=None, reduction='auto', **kwargs):
        if kwargs:
            raise TypeError('unexpected keyword arguments %s' % sorted(kwargs))
        self.name = name
        self.reduction = reduction

    def __call__(self, y_true, y_pred, sample_weight=None):
        per_sample = self.call(y_true, y_pred)
        if self.reduction in ('none', None):
            return per_sample
        return _mean_diff(per_sample, None, 1.0)

    def get_config(self):
        return {'name': self.name, 'reduction': self.reduction}


class LogProbLoss(Loss):
    """losses.py:26-66: per-sample negative log-probability of `samples` under the decoder distribution."""

    def __init__(self, name='log_prob_loss', **kwargs):
        super(LogProbLoss, self).__init__(name=name, **kwargs)

    def call(self, samples, decoder):
        return -decoder.log_prob(samples)


class PotentialEnergyLogProbLoss(Loss):
    """losses.py:69-125: potential(samples) - decoder.log_prob(samples); samples drawn from the decoder if None."""

    def __init__(self, potential, name='pot_log_prob_loss', **kwargs):
        super(PotentialEnergyLogProbLoss, self).__init__(name=name, **kwargs)
        self.potential = potential

    def call(self, samples, decoder):
        if samples is None:
            samples = decoder.sample()
        return as_tensor(self.potential(samples)) - decoder.log_prob(samples)

    def get_config(self):
        config = super(PotentialEnergyLogProbLoss, self).get_config()
        config.update({"potential": self.potential})
        return config


class InfoRegularizer(object):
    """losses.py:128-198: weight * call(dist_a, dist_b, samples); samples default to a draw from `sample_dist`."""

    def __init__(self, weight=1.0, sample_dist='dist_a', name='info_reg', **kwargs):
        if kwargs:
            raise TypeError('unexpected keyword arguments %s' % sorted(kwargs))
        self.name = name
        self.weight = np.float32(weight)
        if sample_dist in ['dist_a', 'dist_b']:
            self.sample_dist = sample_dist
        else:
            raise ValueError("sample_dist must be one of 'dist_a' or 'dist_b'.")

    def __call__(self, dist_a, dist_b, samples=None):
        if samples is None:
            if self.sample_dist == 'dist_a':
                samples = dist_a.sample()
            elif self.sample_dist == 'dist_b':
                samples = dist_b.sample()
        return float(self.weight) * self.call(dist_a, dist_b, samples)

    def call(self, dist_a, dist_b, samples):
        raise NotImplementedError(
            "In any subclass, a 'call' method must be implemented, taking the arguments dist_a, dist_b, samples.")


class NonRegularizer(InfoRegularizer):
    """losses.py:201-223: no regularisation."""

    def call(self, dist_a, dist_b, samples):
        return 0.0


class KLDivergenceEstimate(InfoRegularizer):
    """losses.py:226-253: single-sample Monte-Carlo KL, mean(log a(s) - log b(s))."""

    def call(self, dist_a, dist_b, samples):
        return _mean_diff(dist_a.log_prob(samples), dist_b.log_prob(samples))


class LogProbRegularizer(InfoRegularizer):
    """losses.py:256-296: mean(-log b(s)) (dist_a only supplies the samples)."""

    def call(self, dist_a, dist_b, samples):
        return _mean_diff(dist_b.log_prob(samples), None, -1.0)


class ReverseKLDivergenceEstimate(InfoRegularizer):
    """losses.py:299-330: mean(log b(s) - log a(s)) with samples from dist_b."""

    def __init__(self, **kwargs):
        super(ReverseKLDivergenceEstimate, self).__init__(**kwargs)
        self.sample_dist = 'dist_b'

    def call(self, dist_a, dist_b, samples):
        return _mean_diff(dist_b.log_prob(samples), dist_a.log_prob(samples))
