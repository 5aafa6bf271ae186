"""VAE ELBO oracle: forward, analytic backward, Adam -- NumPy, test infrastructure only.

Follows
  * `vaemolsim/models.py:289-322` (`VAE.call`: encoder -> sample -> prior -> regulariser -> decoder)
  * `vaemolsim/models.py:206-229` (`MappingToDistribution.call`: FCDeepNN then distribution layer)
  * `vaemolsim/losses.py:253`     (`KLDivergenceEstimate.call`: mean(log q(z|x) - log p(z)))
  * `vaemolsim/losses.py:174-194` (`InfoRegularizer.__call__`: weight * call(...))
  * `vaemolsim/losses.py:58`      (`LogProbLoss.call`: -log p(x|z); Keras mean reduction over the batch)
  * `vaemolsim/dists.py:414-439`  (`FlowedDistribution.call`) + TFP `TransformedDistribution.log_prob`
    = base.log_prob(chain.inverse(z)) + chain.inverse_log_det_jacobian(z)
with the model of `vaemolsim/tests/test_models.py:161-185` (IndependentNormal encoder/decoder, N(0,I) prior; SURVEY
config C1) and `tests/test_models.py:190-228` / `Using_Normalizing_Flows.ipynb` cell 10 (RealNVP-RQS prior; C2).
The optimiser is Keras Adam (`tests/test_models.py:181`: lr 1e-3, beta 0.9/0.999, eps 1e-7).
"""
import numpy as np

from . import dists, flows, nets, rqs
from .rqs import sigmoid, softplus_tf


# ----------------------------------------------------------------------------- construction
def init_vae(seed, dx=6, dz=2, hidden=200, prior='normal', num_blocks=4, num_bins=32, flow_hidden=100,
             bin_range=(-10.0, 10.0), dtype=np.float32):
    """Deterministic weights shared by oracle and kernels (Keras initialiser RNG is not reproducible)."""
    rng = np.random.default_rng(seed)
    P = dict(dx=dx, dz=dz, hidden=hidden, prior=prior, num_bins=num_bins, bin_range=tuple(bin_range))
    P['enc'] = nets.fcdeepnn_init(rng, dx, [hidden], (2 * dz,), dtype=dtype)
    P['dec'] = nets.fcdeepnn_init(rng, dz, [hidden], (2 * dx,), dtype=dtype)
    if prior == 'realnvp':
        P['flow'] = flows.realnvp_init(rng, dz, num_blocks, num_bins, flow_hidden, dtype)
    elif prior != 'normal':
        raise ValueError(prior)
    return P


def param_list(P):
    """Flat, ordered list of (name, array) -- the order of the flat parameter buffer."""
    out = []
    for net in ('enc', 'dec'):
        for li, (W, b) in enumerate(P[net]):
            out += [('%s.%d.W' % (net, li), W), ('%s.%d.b' % (net, li), b)]
    for bi, blk in enumerate(P.get('flow', [])):
        for nm in ('d1', 'w', 'h', 's'):
            out += [('flow.%d.%s.W' % (bi, nm), blk[nm][0]), ('flow.%d.%s.b' % (bi, nm), blk[nm][1])]
    return out


def param_count(P):
    return int(sum(a.size for _, a in param_list(P)))


def cast_params(P, dtype):
    Q = {k: v for k, v in P.items() if k not in ('enc', 'dec', 'flow')}
    Q['enc'] = [(W.astype(dtype), b.astype(dtype)) for W, b in P['enc']]
    Q['dec'] = [(W.astype(dtype), b.astype(dtype)) for W, b in P['dec']]
    if 'flow' in P:
        Q['flow'] = [{k: (v[0].astype(dtype), v[1].astype(dtype)) for k, v in blk.items()} for blk in P['flow']]
    return Q


# ----------------------------------------------------------------------------- pieces
def _mlp_fwd(x, layers):
    h = nets.dense(x, layers[0][0], layers[0][1], 'relu')
    return nets.dense(h, layers[1][0], layers[1][1], None), h


def _mlp_bwd(x, h, layers, g_out):
    (W1, b1), (W2, b2) = layers
    gW2 = h.T @ g_out
    gb2 = g_out.sum(0)
    gh = (g_out @ W2.T) * (h > 0)
    gW1 = x.T @ gh
    gb1 = gh.sum(0)
    gx = gh @ W1.T
    return gx, [(gW1, gb1), (gW2, gb2)]


def _normal_lp_bwd(x, loc, scale, g):
    """d/d(x, loc, scale) of sum_d normal_log_prob given per-row upstream g [B]."""
    u = x / scale - loc / scale
    g = g[:, None]
    return g * (-u / scale), g * (u / scale), g * ((u * u - 1) / scale)


def encoder_dist(P, x):
    p, h = _mlp_fwd(x, P['enc'])
    loc, scale = dists.independent_normal_params(p, P['dz'])
    return loc, scale, p, h


def decoder_dist(P, z):
    p, h = _mlp_fwd(z, P['dec'])
    loc, scale = dists.independent_normal_params(p, P['dx'])
    return loc, scale, p, h


def prior_log_prob(P, z, keep=None):
    """log p(z): N(0, I) or RealNVP-RQS transformed N(0, I) (density direction = chain inverse)."""
    dt = z.dtype
    if P['prior'] == 'normal':
        return dists.normal_log_prob(z, dt.type(0), dt.type(1)).sum(-1).astype(dt)
    v = z
    ildj = np.zeros(z.shape[0], dt)
    n = len(P['flow'])
    for i in reversed(range(n)):
        if keep is not None:
            keep[i] = v
        v, l = flows.realnvp_block(v, P['flow'][i], i, P['num_bins'], P['bin_range'], inverse=True)
        ildj = ildj + l
    if keep is not None:
        keep['base'] = v
    return (dists.normal_log_prob(v, dt.type(0), dt.type(1)).sum(-1) + ildj).astype(dt)


def prior_sample_and_log_prob(P, eps):
    """TransformedDistribution.experimental_sample_and_log_prob: y = fwd(x0), lp = base_lp(x0) - fldj(x0)."""
    dt = eps.dtype
    lp0 = dists.normal_log_prob(eps, dt.type(0), dt.type(1)).sum(-1).astype(dt)
    if P['prior'] == 'normal':
        return eps, lp0
    y, fldj = flows.realnvp_forward(eps, P['flow'], P['num_bins'], P['bin_range'])
    return y, (lp0 - fldj).astype(dt)


# ----------------------------------------------------------------------------- ELBO
def elbo_forward(P, x, eps, weight=1.0, keep=None):
    """Returns dict(z, logq, logpz, logpx, kl, nll, loss).  loss = mean(-logpx) + weight * mean(logq - logpz)."""
    dt = x.dtype
    loc, scale, pe, he = encoder_dist(P, x)
    z = dists.normal_sample(loc, scale, eps)
    logq = dists.normal_log_prob(z, loc, scale).sum(-1).astype(dt)
    fk = {} if keep is not None else None
    logpz = prior_log_prob(P, z, fk)
    locx, scalex, pd, hd = decoder_dist(P, z)
    logpx = dists.normal_log_prob(x, locx, scalex).sum(-1).astype(dt)
    kl = np.mean(logq - logpz, dtype=dt)
    nll = np.mean(-logpx, dtype=dt)
    out = dict(z=z, logq=logq, logpz=logpz, logpx=logpx, kl=kl, nll=nll, loss=dt.type(nll + dt.type(weight) * kl))
    if keep is not None:
        keep.update(loc=loc, scale=scale, pe=pe, he=he, locx=locx, scalex=scalex, pd=pd, hd=hd, flow=fk)
    return out


def _flow_block_bwd(P, i, v_in, g_out, g_ldj):
    """Backward through one RealNVP block applied in the inverse (density) direction."""
    blk = P['flow'][i]
    K = P['num_bins']
    D = v_in.shape[-1]
    cs, ts = flows.realnvp_split(i, D)
    c = v_in[:, cs]
    B = v_in.shape[0]
    cin = np.ones((B, 1), v_in.dtype) if c.shape[-1] == 0 else c
    rw, rh, rs, hid = flows.spline_net_raw(c, blk, K)
    Dt = ts.stop - ts.start
    g_t, g_rw, g_rh, g_rs = rqs.rqs_backward_raw(v_in[:, ts], rw, rh, rs, P['bin_range'][0], P['bin_range'][1],
                                                 g_out[:, ts], np.broadcast_to(g_ldj[:, None], (B, Dt)),
                                                 inverse_dir=True)
    g_in = g_out.copy()
    g_in[:, ts] = g_t
    grads = {}
    g_hid = np.zeros_like(hid)
    for nm, g in (('w', g_rw), ('h', g_rh), ('s', g_rs)):
        g2 = g.reshape(B, -1)
        grads[nm] = (hid.T @ g2, g2.sum(0))
        g_hid = g_hid + g2 @ blk[nm][0].T
    g_pre = g_hid * (1 - hid * hid)
    grads['d1'] = (cin.T @ g_pre, g_pre.sum(0))
    if c.shape[-1] > 0:
        g_in[:, cs] = g_in[:, cs] + g_pre @ blk['d1'][0].T
    return g_in, grads


def elbo_backward(P, x, eps, weight=1.0):
    """Analytic reverse mode of `elbo_forward`'s loss.  Returns (forward dict, grads in the layout of P)."""
    dt = x.dtype
    keep = {}
    out = elbo_forward(P, x, eps, weight, keep)
    B = x.shape[0]
    z = out['z']
    w = dt.type(weight)
    g_logpx = np.full(B, -1.0 / B, dt)
    g_logq = np.full(B, w / B, dt)
    g_logpz = np.full(B, -w / B, dt)
    G = {}
    # decoder
    _, g_locx, g_scalex = _normal_lp_bwd(x, keep['locx'], keep['scalex'], g_logpx)
    raw_x = keep['pd'][:, P['dx']:]
    g_pd = np.concatenate([g_locx, g_scalex * sigmoid(raw_x)], axis=-1)
    g_z, G['dec'] = _mlp_bwd(z, keep['hd'], P['dec'], g_pd)
    # prior
    if P['prior'] == 'normal':
        g_z = g_z + g_logpz[:, None] * (-z)
    else:
        fk = keep['flow']
        g_v = g_logpz[:, None] * (-fk['base'])
        G['flow'] = [None] * len(P['flow'])
        for i in range(len(P['flow'])):
            g_v, G['flow'][i] = _flow_block_bwd(P, i, fk[i], g_v, g_logpz)
        g_z = g_z + g_v
    # encoder log q(z|x) (explicit z path + parameter path) and reparameterisation
    gq_z, gq_loc, gq_scale = _normal_lp_bwd(z, keep['loc'], keep['scale'], g_logq)
    g_z = g_z + gq_z
    g_loc = gq_loc + g_z
    g_scale = gq_scale + g_z * eps
    raw = keep['pe'][:, P['dz']:]
    g_pe = np.concatenate([g_loc, g_scale * sigmoid(raw)], axis=-1)
    _, G['enc'] = _mlp_bwd(x, keep['he'], P['enc'], g_pe)
    return out, G


def grad_list(P, G):
    """Gradients in `param_list` order."""
    out = []
    for net in ('enc', 'dec'):
        for li, (gW, gb) in enumerate(G[net]):
            out += [('%s.%d.W' % (net, li), gW), ('%s.%d.b' % (net, li), gb)]
    for bi, blk in enumerate(G.get('flow', [])):
        for nm in ('d1', 'w', 'h', 's'):
            out += [('flow.%d.%s.W' % (bi, nm), blk[nm][0]), ('flow.%d.%s.b' % (bi, nm), blk[nm][1])]
    return out


def flatten(named):
    return np.concatenate([a.reshape(-1) for _, a in named])


def adam_step(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """Keras Adam: lr_t = lr sqrt(1-b2^t)/(1-b1^t); m,v updates; theta -= lr_t m / (sqrt(v) + eps)."""
    dt = theta.dtype
    m[:] = m + (g - m) * dt.type(1 - b1)
    v[:] = v + (g * g - v) * dt.type(1 - b2)
    lr_t = dt.type(lr * np.sqrt(1 - b2**t) / (1 - b1**t))
    theta[:] = theta - lr_t * m / (np.sqrt(v) + dt.type(eps))
    return theta, m, v
