"""RealNVP / MAF rational-quadratic-spline flow oracle -- NumPy, test infrastructure only.

Follows
  * `vaemolsim/flows.py:154-207` (`SplineBijector.call`: Dense(tanh) -> three heads -> activations -> RQS;
                                   empty conditioner input replaced by ones `:184-185`)
  * `vaemolsim/flows.py:281-323` (`RQSSplineRealNVP.build`: masking pattern per block, chain = blocks[::-1])
  * `vaemolsim/flows.py:489-515` (`MaskedSplineBijector.call`: three MADE nets -> activations -> RQS)
  * `vaemolsim/flows.py:597-644` (`RQSSplineMAF.build`: block input orders, chain = blocks[::-1])
  * `vaemolsim/flows.py:15-60`   (`make_domain_transform`: Shift, Scale, Shift chain)
and tensorflow-probability v0.23.0 `real_nvp.py` (`_forward/_inverse/_*_log_det_jacobian`, reverse mask
for negative `num_masked`), `masked_autoregressive.py` (`MaskedAutoregressiveFlow._forward` = D passes
from zeros, `_inverse` = one pass), `chain.py`, `transformed_distribution.py` (third-party, not vendored).
"""
import numpy as np

from . import nets, rqs


# ----------------------------------------------------------------------------- SplineBijector (RealNVP conditioner)
def spline_net_init(rng, din, data_dim, num_bins=32, hidden_dim=200, dtype=np.float32):
    """flows.py:134-152: d1 Dense(H, tanh) then heads w, h (data_dim*K) and s (data_dim*(K-1)); truncated_normal."""
    din_eff = max(din, 1)  # empty input -> ones((B,1)), flows.py:184-185
    tn = nets.truncated_normal
    return dict(
        d1=(tn(rng, din_eff, hidden_dim, dtype=dtype), np.zeros(hidden_dim, dtype)),
        w=(tn(rng, hidden_dim, data_dim * num_bins, dtype=dtype), np.zeros(data_dim * num_bins, dtype)),
        h=(tn(rng, hidden_dim, data_dim * num_bins, dtype=dtype), np.zeros(data_dim * num_bins, dtype)),
        s=(tn(rng, hidden_dim, data_dim * (num_bins - 1), dtype=dtype), np.zeros(data_dim * (num_bins - 1), dtype)),
    )


def spline_net_raw(x_cond, p, num_bins):
    """x_cond [B, din] (din may be 0) -> raw_w, raw_h [B, Dt, K], raw_s [B, Dt, K-1], hidden [B, H]."""
    B = x_cond.shape[0]
    dt = p['d1'][0].dtype
    if x_cond.shape[-1] == 0:
        x_cond = np.ones((B, 1), dtype=dt)
    hid = nets.dense(x_cond.astype(dt), p['d1'][0], p['d1'][1], 'tanh')
    rw = nets.dense(hid, *p['w']).reshape(B, -1, num_bins)
    rh = nets.dense(hid, *p['h']).reshape(B, -1, num_bins)
    rs = nets.dense(hid, *p['s']).reshape(B, -1, num_bins - 1)
    return rw, rh, rs, hid


def realnvp_split(i, D):
    """flows.py:290-306 + TFP RealNVP reverse mask.  Returns (cond_slice, trans_slice) for block i."""
    if D == 1:
        return slice(0, 0), slice(0, 1)
    if i % 2 == 0:
        m = D // 2
        return slice(0, m), slice(m, D)
    m = D - D // 2  # num_masked = -m: condition on the LAST m dims, transform the first D//2
    return slice(D - m, D), slice(0, D - m)


def realnvp_block(v, p, i, num_bins, bin_range, inverse):
    """One RealNVP block applied forward (x->y, fldj) or inverse (y->x, ildj); ldj summed over event."""
    D = v.shape[-1]
    cs, ts = realnvp_split(i, D)
    rw, rh, rs, _ = spline_net_raw(v[:, cs], p, num_bins)
    fn = rqs.rqs_inverse_raw if inverse else rqs.rqs_forward_raw
    out_t, ldj = fn(v[:, ts], rw, rh, rs, bin_range[0], bin_range[1])
    out = v.copy()
    out[:, ts] = out_t
    return out, ldj.sum(axis=-1).astype(v.dtype)


def realnvp_init(rng, D, num_blocks=4, num_bins=32, hidden_dim=200, dtype=np.float32):
    blocks = []
    for i in range(num_blocks):
        cs, ts = realnvp_split(i, D)
        blocks.append(spline_net_init(rng, cs.stop - cs.start, ts.stop - ts.start, num_bins, hidden_dim, dtype))
    return blocks


def realnvp_forward(x, blocks, num_bins, bin_range):
    """chain.forward: block_0 first (flows.py:323).  Returns (y, fldj [B])."""
    ldj = np.zeros(x.shape[0], x.dtype)
    for i, p in enumerate(blocks):
        x, l = realnvp_block(x, p, i, num_bins, bin_range, inverse=False)
        ldj = ldj + l
    return x, ldj


def realnvp_inverse(y, blocks, num_bins, bin_range):
    """chain.inverse: block_{n-1}^-1 first.  Returns (x, ildj [B])."""
    ldj = np.zeros(y.shape[0], y.dtype)
    for i in reversed(range(len(blocks))):
        y, l = realnvp_block(y, blocks[i], i, num_bins, bin_range, inverse=True)
        ldj = ldj + l
    return y, ldj


# ----------------------------------------------------------------------------- MaskedSplineBijector / MAF
def maf_block_init(rng, D, num_bins=32, hidden_dim=200, input_order='left-to-right', cond_size=0, dtype=np.float32):
    """flows.py:454-487: three MADE nets (K, K, K-1 params), one hidden layer, truncated_normal kernels."""
    mk = lambda params: nets.made_init(rng, params, D, [hidden_dim], input_order, cond_size, 'truncated_normal', dtype)
    return dict(w=mk(num_bins), h=mk(num_bins), s=mk(num_bins - 1))


def maf_raw(v, p, num_bins, cond=None):
    return (nets.made_forward(v, p['w'], num_bins, cond), nets.made_forward(v, p['h'], num_bins, cond),
            nets.made_forward(v, p['s'], num_bins - 1, cond))


def maf_block_inverse(y, p, num_bins, bin_range, cond=None):
    """MaskedAutoregressiveFlow._inverse: params from y, one pass.  Returns (x, ildj [B])."""
    rw, rh, rs = maf_raw(y, p, num_bins, cond)
    x, l = rqs.rqs_inverse_raw(y, rw, rh, rs, bin_range[0], bin_range[1])
    return x, l.sum(axis=-1).astype(y.dtype)


def maf_block_forward(x, p, num_bins, bin_range, cond=None):
    """MaskedAutoregressiveFlow._forward: y0 = 0; D passes y <- RQS(params(y)).forward(x).  (y, fldj [B])."""
    y = np.zeros_like(x)
    for _ in range(x.shape[-1]):
        rw, rh, rs = maf_raw(y, p, num_bins, cond)
        y, l = rqs.rqs_forward_raw(x, rw, rh, rs, bin_range[0], bin_range[1])
    # fldj(x) = -ildj(forward(x)); params(y) are those of the last pass for autoregressive-consistent y
    return y, l.sum(axis=-1).astype(x.dtype)


def maf_init(rng, D, num_blocks=2, order_seed=None, num_bins=32, hidden_dim=200, cond_size=0, dtype=np.float32):
    orders = nets.maf_block_orders(num_blocks, D, order_seed)
    return [maf_block_init(rng, D, num_bins, hidden_dim, o, cond_size, dtype) for o in orders]


def maf_forward(x, blocks, num_bins, bin_range, cond=None):
    ldj = np.zeros(x.shape[0], x.dtype)
    for p in blocks:
        x, l = maf_block_forward(x, p, num_bins, bin_range, cond)
        ldj = ldj + l
    return x, ldj


def maf_inverse(y, blocks, num_bins, bin_range, cond=None):
    ldj = np.zeros(y.shape[0], y.dtype)
    for p in reversed(blocks):
        y, l = maf_block_inverse(y, p, num_bins, bin_range, cond)
        ldj = ldj + l
    return y, ldj


# ----------------------------------------------------------------------------- domain transform
def domain_transform_params(domain_list, target, from_target=False):
    """flows.py:31-47: (shift1, scale, shift2), applied as ((x + shift1) * scale) + shift2."""
    t_l = target[1] - target[0]
    t_mean = 0.5 * (target[1] + target[0])
    d_l = np.array([(b - a) for a, b in domain_list], dtype='float32')
    d_mean = np.array([0.5 * (a + b) for a, b in domain_list], dtype='float32')
    if from_target:
        return -t_mean * np.ones_like(d_mean), d_l / t_l, d_mean
    return -d_mean, t_l / d_l, t_mean * np.ones_like(d_mean)


def domain_transform_forward(x, prm):
    s1, sc, s2 = prm
    return ((x + s1) * sc + s2).astype(np.float32)


def domain_transform_inverse(y, prm):
    s1, sc, s2 = prm
    return ((y - s2) / sc - s1).astype(np.float32)


def domain_transform_fldj(prm):
    return np.float32(np.sum(np.log(np.abs(prm[1]))))
