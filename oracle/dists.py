"""Distribution oracle (log_prob, reparameterised sample from given noise) -- NumPy, test infrastructure only.

Follows
  * `vaemolsim/dists.py:28-87`   (`make_param_transform`: parameter_properties bijectors; von Mises atan2 + SoftClip)
  * `vaemolsim/dists.py:197-217` (`IndependentBlockwise.call`: split by param_nums, per-dof dist, Blockwise)
  * `vaemolsim/dists.py:307-340` (`AutoregressiveBlockwise.call`: raw = inputs + MADE(samples, cond); Autoregressive)
  * `vaemolsim/dists.py:589-610` (`IndependentVonMises.new`: split 3, atan2, softplus concentration)
  * `vaemolsim/dists.py:688-704` (`IndependentDeterministic.new`)
  * `vaemolsim/dists.py:414-439` (`FlowedDistribution.call`) with TFP `TransformedDistribution`
and tensorflow-probability v0.23.0 `distributions/normal.py::_log_prob`, `von_mises.py::_log_prob`,
`blockwise.py`, `autoregressive.py::_sample_n/_log_prob`, `layers/distribution_layer.py::IndependentNormal.new`
(third-party, not vendored).
"""
import numpy as np
from scipy import special

from . import nets
from .rqs import softplus_tf

EPS32 = np.float32(np.finfo(np.float32).eps)
HALF_LOG_2PI = 0.5 * np.log(2.0 * np.pi)
LOG_2PI = np.log(2.0 * np.pi)


def i0e(x):
    from scipy.special import i0e as _i0e
    return _i0e(np.asarray(x, dtype=np.float64))


# ----------------------------------------------------------------------------- Normal
def normal_log_prob(x, loc, scale):
    """TFP Normal._log_prob, elementwise: -0.5 (x/s - m/s)^2 - (0.5 log 2pi + log s)."""
    dt = np.result_type(x, loc, scale)
    x, loc, scale = (np.asarray(a, dtype=dt) for a in (x, loc, scale))
    z = x / scale - loc / scale
    return (dt.type(-0.5) * z * z - (dt.type(HALF_LOG_2PI) + np.log(scale))).astype(dt)


def independent_normal_params(params, event_size):
    """tfp.layers.IndependentNormal.new: loc, scale = split(params, 2); scale = softplus(raw)."""
    loc, raw = params[..., :event_size], params[..., event_size:2 * event_size]
    return loc, softplus_tf(raw)


def independent_normal_log_prob(x, params):
    D = x.shape[-1]
    loc, scale = independent_normal_params(params, D)
    return normal_log_prob(x, loc, scale).sum(axis=-1).astype(x.dtype)


def normal_sample(loc, scale, eps):
    """TFP Normal._sample_n: eps * scale + loc."""
    return (eps * scale + loc).astype(loc.dtype)


# ----------------------------------------------------------------------------- von Mises
def vonmises_log_prob(x, loc, conc):
    """TFP VonMises._log_prob: conc (cos(x - loc) - 1) - log(2 pi) - log(i0e(conc))."""
    dt = np.result_type(x, loc, conc)
    z = np.asarray(x, dt) - np.asarray(loc, dt)
    return (np.asarray(conc, dt) * (np.cos(z) - 1) - dt.type(LOG_2PI) - np.log(i0e(conc)).astype(dt)).astype(dt)


def independent_vonmises_params(params, event_size):
    """dists.py:602-607: sine, cosine, scale = split(params, 3); loc = atan2(sine, cosine); conc = softplus."""
    s, c, k = (params[..., i * event_size:(i + 1) * event_size] for i in range(3))
    return np.arctan2(s, c).astype(params.dtype), softplus_tf(k)


def independent_vonmises_log_prob(x, params):
    loc, conc = independent_vonmises_params(params, x.shape[-1])
    return vonmises_log_prob(x, loc, conc).sum(axis=-1).astype(x.dtype)


# ----------------------------------------------------------------------------- make_param_transform / Blockwise
def param_transform(kind, p):
    """dists.py:56-78.  kind in {'normal','vonmises'}; p [..., n_params] -> dict of constrained parameters.

    Normal: parameter_properties => loc identity, scale Softplus(low=eps) i.e. softplus(x) + eps32.
    VonMises: loc = atan2(p0, p1); concentration = SoftClip(low=eps32, high=sqrt(max32)/2)(p2), restated as
    softplus(x) + eps32 -- the soft upper clip at 9.2e18 is unreachable in fp32 practice [unverified vs TFP].
    """
    if kind == 'normal':
        return dict(loc=p[..., 0], scale=(softplus_tf(p[..., 1]) + EPS32).astype(p.dtype))
    if kind == 'vonmises':
        return dict(loc=np.arctan2(p[..., 0], p[..., 1]).astype(p.dtype),
                    concentration=(softplus_tf(p[..., 2]) + EPS32).astype(p.dtype))
    raise ValueError(kind)


PARAM_NUMS = {'normal': 2, 'vonmises': 3}  # dists.py:164-173 (+1 for von Mises)


def blockwise_log_prob_split(x, params_per_dof, kinds):
    """x [B, D]; params_per_dof: list of D arrays [B, n_i].  Blockwise log_prob = sum over dofs."""
    lp = np.zeros(x.shape[0], x.dtype)
    for i, kind in enumerate(kinds):
        t = param_transform(kind, params_per_dof[i])
        if kind == 'normal':
            lp = lp + normal_log_prob(x[:, i], t['loc'], t['scale'])
        else:
            lp = lp + vonmises_log_prob(x[:, i], t['loc'], t['concentration'])
    return lp.astype(x.dtype)


def independent_blockwise_log_prob(x, params, kinds):
    """dists.py:210-217: params [B, sum(param_nums)] split contiguously per dof."""
    nums = [PARAM_NUMS[k] for k in kinds]
    offs = np.concatenate([[0], np.cumsum(nums)])
    return blockwise_log_prob_split(x, [params[:, offs[i]:offs[i + 1]] for i in range(len(kinds))], kinds)


def autoregressive_blockwise_log_prob(x, inputs, made_layers, kinds, cond=None, activation=None):
    """dists.py:326-340 + TFP Autoregressive._log_prob: distribution_fn(x).log_prob(x).

    inputs [B, D, Pmax]; raw = inputs + MADE(x, cond); dof i uses raw[:, i, :param_nums[i]] -- note the
    reference passes the full Pmax-wide slice and the transform reads only the leading entries.
    `activation` is the AutoregressiveNetwork hidden activation (TFP default None = linear; the MC notebook
    passes only hidden_units, `MC_Moves_with_VAEs.ipynb` cell 17).
    """
    pmax = inputs.shape[-1]
    raw = inputs + nets.made_forward(x, made_layers, pmax, cond, activation=activation)
    return blockwise_log_prob_split(x, [raw[:, i, :] for i in range(len(kinds))], kinds)


def autoregressive_blockwise_sample_normal(inputs, made_layers, eps, cond=None, activation=None):
    """TFP Autoregressive._sample_n for all-Normal dofs with FIXED noise eps [B, D] reused every step:
    samples0 = dist_fn(ones).sample(); then num_steps = D times samples = dist_fn(samples).sample()."""
    B, D, pmax = inputs.shape
    s = np.ones((B, D), inputs.dtype)
    for _ in range(D + 1):
        raw = inputs + nets.made_forward(s, made_layers, pmax, cond, activation=activation)
        loc = raw[..., 0]
        scale = softplus_tf(raw[..., 1]) + EPS32
        s = normal_sample(loc, scale, eps)
    return s


# ----------------------------------------------------------------------------------- von Mises sampling gradient
def i1e(x):
    return special.i1e(x)


def vonmises_cdf_and_dconcentration(x, conc):
    """TFP v0.23 `von_mises.von_mises_cdf` and its derivative with respect to the concentration [TFP-recalled, checked
    against finite differences of scipy's CDF in tests/test_oracle.py]: backward-recurrence series (20 terms, Hill 1977
    table I, D = 8) below concentration 10.5, corrected Normal approximation above; the clipped series value has zero
    derivative.  x in [-pi, pi] (centred sample), float64."""
    x = np.asarray(x, np.float64)
    conc = np.asarray(conc, np.float64) + 0.0 * x
    # series
    rn = np.zeros_like(x); drn = np.zeros_like(x); vn = np.zeros_like(x); dvn = np.zeros_like(x)
    for n in range(20, 0, -1):
        den = 2.0 * n / conc + rn
        dden = -2.0 * n / conc**2 + drn
        rn = 1.0 / den
        drn = -dden / den**2
        mult = np.sin(n * x) / n + vn
        dvn = drn * mult + rn * dvn
        vn = rn * mult
    cdf_s = 0.5 + x / (2.0 * np.pi) + vn / np.pi
    dcdf_s = (dvn / np.pi) * ((cdf_s >= 0.0) & (cdf_s <= 1.0))
    cdf_s = np.clip(cdf_s, 0.0, 1.0)
    # corrected Normal approximation, differentiated by hand (TFP uses value_and_gradient)
    i0 = special.i0e(conc)
    ratio = special.i1e(conc) / i0 - 1.0            # d log i0e / d conc
    z = np.sqrt(2.0 / np.pi) / i0 * np.sin(0.5 * x)
    dz = -z * ratio
    z2, z3, z4 = z**2, z**3, z**4
    c = 24.0 * conc
    a = (c - 2.0 * z2 - 16.0) / 3.0
    bnum = z4 + 1.75 * z2 + 83.5
    bden = c - 56.0 - z2 + 3.0
    d = a - bnum / bden
    xi = z - z3 / d**2
    da = (24.0 - 4.0 * z * dz) / 3.0
    dbnum = (4.0 * z3 + 3.5 * z) * dz
    dbden = 24.0 - 2.0 * z * dz
    dd = da - (dbnum * bden - bnum * dbden) / bden**2
    dxi = dz - (3.0 * z2 * dz * d**2 - z3 * 2.0 * d * dd) / d**4
    cdf_n = 0.5 * (1.0 + special.erf(xi / np.sqrt(2.0)))
    dcdf_n = np.exp(-0.5 * xi**2) / np.sqrt(2.0 * np.pi) * dxi
    use_series = conc < 10.5
    return np.where(use_series, cdf_s, cdf_n), np.where(use_series, dcdf_s, dcdf_n)


def vonmises_sample_dconcentration(s, conc):
    """d sample / d concentration of a centred von Mises sample by implicit reparameterisation (TFP
    `_von_mises_sample_bwd`): -dF/dconc / p(s) with 1 / p = exp(-conc (cos s - 1)) 2 pi i0e(conc)."""
    _, dcdf = vonmises_cdf_and_dconcentration(s, conc)
    inv_prob = np.exp(-conc * (np.cos(s) - 1.0)) * (2.0 * np.pi * special.i0e(conc))
    return -dcdf * inv_prob
