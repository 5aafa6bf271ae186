"""Rational-quadratic-spline (RQS) oracle -- NumPy restatement, test infrastructure only.

Follows
  * `vaemolsim/flows.py:86-93`   (`SplineBijector._bin_positions`: softmax * (max-min-K*1e-2) + 1e-2)
  * `vaemolsim/flows.py:95-101`  (`SplineBijector._slopes`: softplus + 1e-2)
  * `vaemolsim/flows.py:394-409` (same activations for `MaskedSplineBijector`, no reshape)
  * `vaemolsim/flows.py:204-207`, `:512-515` (construction of `tfp.bijectors.RationalQuadraticSpline`)
and the published algorithm of tensorflow-probability v0.23.0
`bijectors/rational_quadratic_spline.py` (`_compute_shared`, `_forward`, `_inverse`,
`_forward_log_det_jacobian`; third-party, not vendored under /root/reference; pinned by
`pyproject.toml:28`) -- **parity unpinned** against real TFP output, see `oracle/__init__.py`.

All functions are dtype-generic: float32 arrays reproduce TF's op-by-op float32 rounding
(sequential cumsum, separate mul/add), float64 arrays give a high-precision reference used for
finite-difference / gradient checks.
"""
import numpy as np

MIN_BIN = 1e-2  # flows.py:92 (+1e-2 on widths/heights)
MIN_SLOPE = 1e-2  # flows.py:101 (+1e-2 on slopes)


def softplus_tf(x):
    """tf.math.softplus: Eigen's thresholded form.

    threshold = log(eps) + 2;  x > -threshold -> x;  x < threshold -> exp(x);  else log1p(exp(x)).
    """
    x = np.asarray(x)
    thr = np.asarray(np.log(np.finfo(x.dtype).eps) + 2.0, dtype=x.dtype)
    with np.errstate(over='ignore'):
        ex = np.exp(x)
        mid = np.log1p(ex)
    return np.where(x > -thr, x, np.where(x < thr, ex, mid)).astype(x.dtype)


def sigmoid(x):
    x = np.asarray(x)
    with np.errstate(over='ignore'):
        return (1.0 / (1.0 + np.exp(-x))).astype(x.dtype)


def softmax_tf(x):
    """tf.math.softmax over the last axis: exp(x - max) / sum(exp(x - max))."""
    x = np.asarray(x)
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return (e / np.sum(e, axis=-1, keepdims=True)).astype(x.dtype)


def bin_positions(raw, bin_min, bin_max):
    """flows.py:90-93 / :400-402.  raw [..., K] -> widths or heights [..., K]."""
    raw = np.asarray(raw)
    dt = raw.dtype
    k = raw.shape[-1]
    # Python float arithmetic on the scale factor exactly as the reference writes it, then cast.
    scale = dt.type(bin_max - bin_min - k * 1e-2)
    return (softmax_tf(raw) * scale + dt.type(MIN_BIN)).astype(dt)


def slopes(raw):
    """flows.py:100-101 / :408-409.  raw [..., K-1] -> interior knot slopes [..., K-1]."""
    raw = np.asarray(raw)
    return (softplus_tf(raw) + raw.dtype.type(MIN_SLOPE)).astype(raw.dtype)


def knot_positions(bin_sizes, range_min):
    """TFP `_knot_positions`: [range_min, cumsum(bin_sizes) + range_min]  ([..., K] -> [..., K+1])."""
    dt = bin_sizes.dtype
    cs = np.cumsum(bin_sizes, axis=-1, dtype=dt) + dt.type(range_min)
    lhs = np.full(cs.shape[:-1] + (1,), range_min, dtype=dt)
    return np.concatenate([lhs, cs], axis=-1)


def _gather(a, idx):
    return np.take_along_axis(a, idx[..., None], axis=-1)[..., 0]


def compute_shared(bw, bh, ks, range_min, x=None, y=None):
    """TFP `RationalQuadraticSpline._compute_shared`.

    bw, bh: [..., K]; ks: [..., K-1]; x or y: [...].
    """
    dt = bw.dtype
    kx = knot_positions(bw, range_min)
    ky = knot_positions(bh, range_min)
    ones = np.ones(ks.shape[:-1] + (1,), dtype=dt)
    kd = np.concatenate([ones, ks, ones], axis=-1)
    kk = kx if y is None else ky
    v = np.asarray(x if y is None else y, dtype=dt)
    kmin, kmax = kk[..., 0], kk[..., -1]
    oob = (v <= kmin) | (v >= kmax)
    v = np.where(oob, kmin, v)
    # searchsorted(kk[..., :-1], v, side='right') - 1, floored at 0
    idx = np.sum(kk[..., :-1] <= v[..., None], axis=-1) - 1
    idx = np.maximum(idx, 0).astype(np.int64)
    x_k, x_k1 = _gather(kx, idx), _gather(kx, idx + 1)
    y_k, y_k1 = _gather(ky, idx), _gather(ky, idx + 1)
    d_k, d_k1 = _gather(kd, idx), _gather(kd, idx + 1)
    h_k = y_k1 - y_k
    w_k = x_k1 - x_k
    s_k = h_k / w_k
    return dict(oob=oob, idx=idx, x_k=x_k, y_k=y_k, d_k=d_k, d_k1=d_k1, h_k=h_k, w_k=w_k, s_k=s_k)


def forward(x, bw, bh, ks, range_min):
    """TFP `_forward` -> y."""
    x = np.asarray(x, dtype=bw.dtype)
    d = compute_shared(bw, bh, ks, range_min, x=x)
    relx = (x - d['x_k']) / d['w_k']
    num = d['h_k'] * (d['s_k'] * relx**2 + d['d_k'] * relx * (1 - relx))
    den = d['s_k'] + (d['d_k1'] + d['d_k'] - 2 * d['s_k']) * relx * (1 - relx)
    return np.where(d['oob'], x, d['y_k'] + num / den).astype(bw.dtype)


def forward_log_det_jacobian(x, bw, bh, ks, range_min):
    """TFP `_forward_log_det_jacobian` -> per-element log dy/dx (not yet summed over the event)."""
    dt = bw.dtype
    x = np.asarray(x, dtype=dt)
    d = compute_shared(bw, bh, ks, range_min, x=x)
    relx = (x - d['x_k']) / d['w_k']
    relx = np.where(d['oob'], dt.type(0.5), relx)
    s, dk, dk1 = d['s_k'], d['d_k'], d['d_k1']
    with np.errstate(invalid='ignore', divide='ignore'):
        grad = (2 * np.log(s) + np.log(dk1 * relx**2 + 2 * s * relx * (1 - relx) + dk * (1 - relx)**2) - 2 * np.log(
            (dk1 + dk - 2 * s) * relx * (1 - relx) + s))
    return np.where(d['oob'], dt.type(0), grad).astype(dt)


def inverse(y, bw, bh, ks, range_min):
    """TFP `_inverse` -> x."""
    dt = bw.dtype
    y = np.asarray(y, dtype=dt)
    d = compute_shared(bw, bh, ks, range_min, y=y)
    rely = np.where(d['oob'], dt.type(0), y - d['y_k'])
    term2 = rely * (d['d_k1'] + d['d_k'] - 2 * d['s_k'])
    a = d['h_k'] * (d['s_k'] - d['d_k']) + term2
    b = d['h_k'] * d['d_k'] - term2
    c = -d['s_k'] * rely
    with np.errstate(invalid='ignore', divide='ignore'):
        relx = np.where(rely == 0, dt.type(0), (2 * c) / (-b - np.sqrt(b**2 - 4 * a * c)))
    return np.where(d['oob'], y, relx * d['w_k'] + d['x_k']).astype(dt)


def inverse_log_det_jacobian(y, bw, bh, ks, range_min):
    """TFP default: ildj(y) = -fldj(inverse(y))."""
    return -forward_log_det_jacobian(inverse(y, bw, bh, ks, range_min), bw, bh, ks, range_min)


# --------------------------------------------------------------------------------------
# Raw-logit entry points: the op boundary of the CUDA kernel (activations fused).
# --------------------------------------------------------------------------------------
def rqs_from_raw(raw_w, raw_h, raw_s, bin_min, bin_max):
    return bin_positions(raw_w, bin_min, bin_max), bin_positions(raw_h, bin_min, bin_max), slopes(raw_s)


def rqs_forward_raw(x, raw_w, raw_h, raw_s, bin_min, bin_max):
    """x [...], raw_w/raw_h [..., K], raw_s [..., K-1] -> (y [...], fldj [...])."""
    bw, bh, ks = rqs_from_raw(raw_w, raw_h, raw_s, bin_min, bin_max)
    return forward(x, bw, bh, ks, bin_min), forward_log_det_jacobian(x, bw, bh, ks, bin_min)


def rqs_inverse_raw(y, raw_w, raw_h, raw_s, bin_min, bin_max):
    """y [...] -> (x [...], ildj [...])."""
    bw, bh, ks = rqs_from_raw(raw_w, raw_h, raw_s, bin_min, bin_max)
    x = inverse(y, bw, bh, ks, bin_min)
    return x, -forward_log_det_jacobian(x, bw, bh, ks, bin_min)


def rqs_backward_raw(v_in, raw_w, raw_h, raw_s, bin_min, bin_max, g_out, g_ldj, inverse_dir=False):
    """Analytic reverse-mode gradient of the raw-logit RQS op.

    forward direction : (y, L)  = (F(x; theta), fldj(x; theta));  loss = <g_out, y> + <g_ldj, L>
    inverse direction : (x, I)  = (F^-1(y; theta), -fldj(x; theta)); loss = <g_out, x> + <g_ldj, I>
    Returns (g_in, g_raw_w, g_raw_h, g_raw_s) with the shapes of the inputs.
    TF `tf.where` semantics out of bounds: d out / d in = 1, all parameter gradients 0.
    """
    raw_w = np.asarray(raw_w)
    dt = raw_w.dtype
    K = raw_w.shape[-1]
    bw, bh, ks = rqs_from_raw(raw_w, raw_h, raw_s, bin_min, bin_max)
    v_in = np.asarray(v_in, dtype=dt)
    g_out = np.asarray(g_out, dtype=dt)
    g_ldj = np.asarray(g_ldj, dtype=dt)
    if inverse_dir:
        d = compute_shared(bw, bh, ks, bin_min, y=v_in)
        x = inverse(v_in, bw, bh, ks, bin_min)
    else:
        d = compute_shared(bw, bh, ks, bin_min, x=v_in)
        x = v_in
    oob, idx = d['oob'], d['idx']
    w, h, s, dk, dk1, xk = d['w_k'], d['h_k'], d['s_k'], d['d_k'], d['d_k1'], d['x_k']
    r = np.where(oob, dt.type(0.5), (x - xk) / w)
    u = r * (1 - r)
    N = s * r * r + dk * u
    Q = s + (dk1 + dk - 2 * s) * u
    P = dk1 * r * r + 2 * s * u + dk * (1 - r) * (1 - r)
    N_r = 2 * s * r + dk * (1 - 2 * r)
    Q_r = (dk1 + dk - 2 * s) * (1 - 2 * r)
    Q_s = 1 - 2 * u
    P_r = 2 * dk1 * r + 2 * s * (1 - 2 * r) - 2 * dk * (1 - r)
    Q2 = Q * Q
    y_r = h * (N_r * Q - N * Q_r) / Q2
    y_s = h * (r * r * Q - N * Q_s) / Q2
    y_dk = h * u * (Q - N) / Q2
    y_dk1 = -h * N * u / Q2
    y_h = N / Q
    L_r = P_r / P - 2 * Q_r / Q
    L_s = 2 / s + 2 * u / P - 2 * Q_s / Q
    L_dk = (1 - r) * (1 - r) / P - 2 * u / Q
    L_dk1 = r * r / P - 2 * u / Q
    F_x = y_r / w
    L_x = L_r / w
    if inverse_dir:
        G = g_out - g_ldj * L_x
        gy = -G / F_x
        gL = -g_ldj
        g_in = G / F_x
    else:
        gy, gL = g_out, g_ldj
        g_in = gy * F_x + gL * L_x
    g_r = gy * y_r + gL * L_r
    g_s = gy * y_s + gL * L_s
    g_dk = gy * y_dk + gL * L_dk
    g_dk1 = gy * y_dk1 + gL * L_dk1
    g_h = gy * y_h + g_s / w
    g_yk = np.array(gy, dtype=dt, copy=True)  # never alias the caller's g_out (masked in place below)
    g_w = -g_s * s / w - g_r * r / w
    g_xk = -g_r / w
    zero = dt.type(0)
    g_in = np.where(oob, g_out, g_in)
    for a in (g_h, g_yk, g_w, g_xk, g_dk, g_dk1):
        a[...] = np.where(oob, zero, a)
    j = np.arange(K)
    lt = j < idx[..., None]
    eq = j == idx[..., None]
    g_bw = lt * g_xk[..., None] + eq * g_w[..., None]
    g_bh = lt * g_yk[..., None] + eq * g_h[..., None]
    cw = dt.type(bin_max - bin_min - K * 1e-2)
    pw, ph = softmax_tf(raw_w), softmax_tf(raw_h)
    g_raw_w = cw * pw * (g_bw - np.sum(pw * g_bw, axis=-1, keepdims=True))
    g_raw_h = cw * ph * (g_bh - np.sum(ph * g_bh, axis=-1, keepdims=True))
    # slopes: kd = [1, ks_0 .. ks_{K-2}, 1];  d_k = kd[idx], d_k1 = kd[idx+1]
    js = np.arange(K - 1)
    g_ks = (js == (idx[..., None] - 1)) * g_dk[..., None] + (js == idx[..., None]) * g_dk1[..., None]
    g_raw_s = g_ks * sigmoid(np.asarray(raw_s, dtype=dt))
    return g_in.astype(dt), g_raw_w.astype(dt), g_raw_h.astype(dt), g_raw_s.astype(dt)
