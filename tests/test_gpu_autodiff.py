"""GPU tests of the reverse mode of the op-by-op path (csrc/autodiff.cu + vaemolsim_b200/_autodiff.py): kernel-level
gradients against the float64 oracle, model-level gradients against finite differences, and ports of the reference's
training tests for compositions outside the fused ELBO family (tests/test_models.py:189-228, models.py:85-139)."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_close
from oracle import dists as odists
from oracle import flows as oflows
from oracle import nets as onets

pytestmark = pytest.mark.gpu


def _i32(a):
    return (C.c_int32 * len(a))(*[int(v) for v in a])


# ------------------------------------------------------------------------------------------------ kernels
def test_blockwise_log_prob_backward_matches_float64_finite_differences(vms):
    """d/dx and d/dparams of sum_d log_prob_d for Normal and von Mises dofs with the transforms of dists.py:56-78 (atan2
    location, softplus + eps concentration / scale): central differences of the float64 oracle."""
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(0)
    B = 257
    kinds = ['normal', 'vonmises', 'normal', 'vonmises']
    nums = [2, 3, 2, 3]
    offs = np.concatenate([[0], np.cumsum(nums)])
    P = int(offs[-1])
    x = rng.uniform(-3, 3, (B, 4))
    params = rng.normal(0, 1.0, (B, P))
    g_lp = rng.normal(size=B)
    kind = [0 if k == 'normal' else 1 for k in kinds]
    loc = [offs[i] for i in range(4)]
    loc2 = [offs[i] + 1 if kinds[i] == 'vonmises' else -1 for i in range(4)]
    sc = [offs[i] + (2 if kinds[i] == 'vonmises' else 1) for i in range(4)]
    xd, pd, gd = v.Tensor.from_numpy(x.astype(np.float32)), v.Tensor.from_numpy(params.astype(np.float32)), \
        v.Tensor.from_numpy(g_lp.astype(np.float32))
    gx, gp = v.Tensor.zeros((B, 4)), v.Tensor.zeros((B, P))
    c.lib.vms_blockwise_log_prob_backward(xd.ptr, 4, pd.ptr, P, B, 4, _i32(kind), _i32(loc), _i32(loc2), _i32(sc), 2, gd.ptr,
                                          gx.ptr, 4, gp.ptr, P, c.stream)
    x32, p32 = x.astype(np.float32).astype(np.float64), params.astype(np.float32).astype(np.float64)
    f = lambda xx, pp: odists.independent_blockwise_log_prob(xx, pp, kinds)
    h = 1e-6
    want_x = np.zeros((B, 4))
    for d in range(4):
        e = np.zeros(4); e[d] = h
        want_x[:, d] = (f(x32 + e, p32) - f(x32 - e, p32)) / (2 * h) * g_lp.astype(np.float32)
    want_p = np.zeros((B, P))
    for j in range(P):
        e = np.zeros(P); e[j] = h
        want_p[:, j] = (f(x32, p32 + e) - f(x32, p32 - e)) / (2 * h) * g_lp.astype(np.float32)
    assert_close(gx.numpy(), want_x, rtol=2e-5, atol=2e-5, what='d log_prob / d x')
    assert_close(gp.numpy(), want_p, rtol=3e-5, atol=3e-5, what='d log_prob / d params')
    # accumulation: a second call adds
    c.lib.vms_blockwise_log_prob_backward(xd.ptr, 4, pd.ptr, P, B, 4, _i32(kind), _i32(loc), _i32(loc2), _i32(sc), 2, gd.ptr,
                                          gx.ptr, 4, gp.ptr, P, c.stream)
    assert_close(gx.numpy(), 2 * want_x, rtol=2e-5, atol=4e-5, what='accumulated d log_prob / d x')


def test_blockwise_sample_backward_normal_pathwise_and_von_mises_implicit(vms):
    """d z / d params: Normal = pathwise (eps recovered from the sample); von Mises = location pathwise + tfp's implicit
    reparameterisation in the concentration, -dF/dk / p(z), against the float64 oracle restatement of tfp's
    `von_mises_cdf` gradient (itself checked against finite differences of scipy's CDF in the CPU suite), both branches
    (series below k = 10.5, corrected Normal approximation above)."""
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(1)
    B = 4000
    # dof 0 Normal (loc, raw scale), dof 1 von Mises (sin, cos, raw concentration)
    params = np.empty((B, 5), np.float32)
    params[:, 0] = rng.normal(size=B)
    params[:, 1] = rng.normal(size=B)
    params[:, 2:4] = rng.normal(size=(B, 2))
    params[:, 4] = rng.uniform(-3, 25, B)  # softplus -> concentrations 0.05 .. 25: both CDF branches
    pd = v.Tensor.from_numpy(params)
    z = v.Tensor((B, 2))
    c.lib.vms_blockwise_sample(pd.ptr, 5, B, 2, _i32([0, 1]), _i32([0, 2]), _i32([-1, 3]), _i32([1, 4]), 2, None, 0, 1234, z.ptr,
                               2, c.stream)
    zn = z.numpy().astype(np.float64)
    g_z = rng.normal(size=(B, 2)).astype(np.float32)
    gp = v.Tensor.zeros((B, 5))
    c.lib.vms_blockwise_sample_backward(pd.ptr, 5, B, 2, _i32([0, 1]), _i32([0, 2]), _i32([-1, 3]), _i32([1, 4]), 2, z.ptr, 2,
                                        v.Tensor.from_numpy(g_z).ptr, 2, gp.ptr, 5, c.stream)
    got = gp.numpy().astype(np.float64)
    p64 = params.astype(np.float64)
    sp = lambda r: np.log1p(np.exp(-np.abs(r))) + np.maximum(r, 0) + 1.1920929e-07
    sig = lambda r: 1.0 / (1.0 + np.exp(-r))
    # Normal
    scale = sp(p64[:, 1])
    assert_close(got[:, 0], g_z[:, 0], rtol=1e-6, atol=1e-6, what='d z / d loc (Normal)')
    assert_close(got[:, 1], g_z[:, 0] * (zn[:, 0] - p64[:, 0]) / scale * sig(p64[:, 1]), rtol=2e-5, atol=2e-5,
                 what='d z / d raw scale (Normal)')
    # von Mises
    s_, c_ = p64[:, 2], p64[:, 3]
    r2 = s_**2 + c_**2
    assert_close(got[:, 2], g_z[:, 1] * c_ / r2, rtol=2e-5, atol=2e-5, what='d z / d sine parameter')
    assert_close(got[:, 3], -g_z[:, 1] * s_ / r2, rtol=2e-5, atol=2e-5, what='d z / d cosine parameter')
    kap = sp(p64[:, 4])
    centred = np.mod(zn[:, 1] - np.arctan2(s_, c_) + np.pi, 2 * np.pi) - np.pi
    want = g_z[:, 1] * odists.vonmises_sample_dconcentration(centred, kap) * sig(p64[:, 4])
    # float32 series / Bessel evaluation: 1e-4 of the gradient's scale
    assert_close(got[:, 4], want, rtol=2e-4, atol=2e-4 * np.abs(want).max(), what='implicit d z / d raw concentration')
    assert (kap < 10.5).sum() > 500 and (kap > 10.5).sum() > 500


def test_periodic_featurise_backward(vms):
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(2)
    B, mask = 300, np.array([False, True, True, False, True])
    x = rng.uniform(-6, 6, (B, 5)).astype(np.float32)
    g_out = rng.normal(size=(B, 8)).astype(np.float32)
    gx = v.Tensor.zeros((B, 5))
    xd, md, gd = v.Tensor.from_numpy(x), v.Tensor.from_numpy(mask.astype(np.uint8)), v.Tensor.from_numpy(g_out)
    c.lib.vms_periodic_featurise_backward(xd.ptr, B, 5, md.ptr, 3, gd.ptr, gx.ptr, c.stream)
    want = np.zeros((B, 5))
    want[:, ~mask] = g_out[:, :2]
    want[:, mask] = -np.sin(x[:, mask].astype(np.float64)) * g_out[:, 2:5] + np.cos(x[:, mask].astype(np.float64)) * g_out[:, 5:8]
    assert_close(gx.numpy(), want, rtol=1e-5, atol=1e-5, what='periodic featurise backward')


# ------------------------------------------------------------------------------------------------ models
def _tape_gradients(v, model, xb, yb):
    from vaemolsim_b200 import _autodiff
    with _autodiff.Tape() as tp:
        loss = model._loss_tensor(v.as_tensor(xb), v.as_tensor(yb), True)
        tp.backward(loss)
    ws = []
    for w in model.weights:
        if not any(w is u for u in ws):
            ws.append(w)
    grads = []
    ws = [w for w in ws if tp.has(w)]  # (moving statistics of batch-norm layers are weights without a gradient)
    for w in ws:
        g = tp.grad(w).numpy().astype(np.float64) if tp.has(w) else np.zeros(w.shape)
        m = getattr(w, '_grad_mask', None)
        grads.append(g * m.numpy() if m is not None else g)
    return float(loss.numpy()), ws, grads


def _directional_check(v, model, xb, yb, rtol, seed=0, h=2e-3, n_dirs=3, training=False):
    """Tape gradient . direction against a central difference of the DEVICE forward along that direction (float32 forward:
    the difference carries ~1e-4 relative noise, enough to catch a missing term, a wrong sign or a wrong transpose)."""
    c = v._abi.ctx()
    loss0, ws, grads = _tape_gradients(v, model, xb, yb)
    assert np.isfinite(loss0) and all(np.isfinite(g).all() for g in grads)
    assert sum(float(np.abs(g).sum()) for g in grads) > 0
    rng = np.random.default_rng(seed)
    base = [w.numpy().copy() for w in ws]
    for _ in range(n_dirs):
        dirs = []
        for w, b0 in zip(ws, base):
            d = rng.normal(size=b0.shape)
            m = getattr(w, '_grad_mask', None)
            dirs.append(d * m.numpy() if m is not None else d)
        want = sum(float((g * d).sum()) for g, d in zip(grads, dirs))
        vals = []
        for sgn in (+1.0, -1.0):
            for w, b0, d in zip(ws, base, dirs):
                a = np.ascontiguousarray((b0 + sgn * h * d).astype(np.float32))
                if w.contiguous:
                    c.lib.vms_memcpy_h2d(w.ptr, a.ctypes.data, a.nbytes, c.stream)
                else:
                    w.assign_cols(0, v.Tensor.from_numpy(a))
            c.synchronize()
            vals.append(float(model._loss_tensor(v.as_tensor(xb), v.as_tensor(yb), training).numpy()))
        got = (vals[0] - vals[1]) / (2 * h)
        assert abs(got - want) <= rtol * max(abs(want), 1e-2) + 5e-4, 'directional derivative %g (finite difference) vs %g (tape)' % (got, want)
    for w, b0 in zip(ws, base):
        a = np.ascontiguousarray(b0.astype(np.float32))
        c.lib.vms_memcpy_h2d(w.ptr, a.ctypes.data, a.nbytes, c.stream)
    c.synchronize()


def _decoder(v, which):
    """Decoders of tests/test_models.py:189-228 at small widths.  The mapping networks use tanh here: a relu network is not
    differentiable along a random direction (units cross zero inside the finite-difference step), and the relu reverse
    mode of vms_dense_backward has its own oracle test (tests/test_gpu_kernels.py)."""
    d = v.dists
    kinds = [d.Normal] * 2 + [d.VonMises] * 2
    fc = lambda target, **kw: v.mappings.FCDeepNN(target, hidden_dim=24, activation='tanh', **kw)
    if which == 'blockwise':
        dist = d.IndependentBlockwise(4, kinds)
        return v.models.MappingToDistribution(dist, mapping=fc(dist.params_size(), periodic_dofs=[False, True, True]),
                                              name='decoder')
    if which == 'autoregressive':
        dist = d.AutoregressiveBlockwise(4, kinds, auto_net_params={'hidden_units': [16, 16]})
        return v.models.MappingToDistribution(dist, mapping=fc(dist.params_size()), name='decoder')
    if which == 'autoregressive-conditional':
        dist = d.AutoregressiveBlockwise(4, kinds, conditional=True, conditional_event_shape=3,
                                         auto_net_params={'hidden_units': [16], 'activation': 'tanh'})
        return v.models.MappingToDistribution(dist, mapping=fc(dist.params_size()), name='decoder')
    if which == 'maf-flowed':
        flow = v.flows.RQSSplineMAF(num_blocks=2, order_seed=3, rqs_params=dict(num_bins=8, hidden_dim=12, bin_range=[-6.0, 6.0]))
        flow(np.ones((1, 4), np.float32))
        dist = d.FlowedDistribution(flow, d.IndependentBlockwise(4, kinds))
        return v.models.MappingToDistribution(dist, mapping=fc(dist.params_size()), name='decoder')
    raise ValueError(which)


@pytest.mark.parametrize('which', ['blockwise', 'autoregressive', 'autoregressive-conditional', 'maf-flowed'])
def test_decoder_training_gradients_match_finite_differences(vms, which):
    """Decoder-only training (Training_VAEs_and_Decoders notebook, cell 40-47): LogProbLoss through FCDeepNN (periodic
    dofs), MADE, blockwise Normal / von Mises log_prob and the MAF flow -- tape gradients vs finite differences."""
    v = vms
    v.set_seed(5)
    rng = np.random.default_rng(3)
    model = _decoder(v, which)
    model.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
    z = rng.normal(size=(96, 3)).astype(np.float32)
    x = np.concatenate([rng.normal(size=(96, 2)), rng.uniform(-3, 3, (96, 2))], axis=1).astype(np.float32)
    model(z)  # build
    for w in model.weights:  # move biases / masked kernels away from their zero initialisation
        a = w.numpy()
        m = getattr(w, '_grad_mask', None)
        a = a + rng.normal(0, 0.1, a.shape).astype(np.float32) * (m.numpy() if m is not None else 1.0)
        v._abi.ctx().lib.vms_memcpy_h2d(w.ptr, np.ascontiguousarray(a).ctypes.data, a.nbytes, v._abi.ctx().stream)
    v._abi.ctx().synchronize()
    _directional_check(v, model, z, x, rtol=2e-2)
    l0 = model.evaluate(z, x, batch_size=96)
    hist = model.fit(z, x, epochs=30, batch_size=32, shuffle=False)
    assert np.isfinite(hist['loss']).all() and model.evaluate(z, x, batch_size=96) < l0 - 0.05, (l0, hist['loss'][-5:])


def test_flow_model_nll_training_matches_oracle_gradient(vms):
    """FlowModel (models.py:85-139) with an RQS MAF over a static N(0, I): NLL gradient against central differences of the
    float64 ORACLE (maf_inverse + standard normal), then `fit` reduces the NLL (Using_Normalizing_Flows cell 10-15)."""
    v = vms
    import vaemolsim_b200._protocols as PR
    rng = np.random.default_rng(7)
    D, K, H, nb = 3, 8, 10, 2
    flow = v.flows.RQSSplineMAF(num_blocks=nb, order_seed=42, rqs_params=dict(num_bins=K, hidden_dim=H))
    fm = v.models.FlowModel(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], D)))
    fm.compile(optimizer=v.models.Adam(2e-3), loss=v.losses.LogProbLoss())
    x = (rng.normal(size=(200, D)) * np.array([0.5, 1.5, 1.0]) + np.array([1.0, -1.0, 0.0])).astype(np.float32)
    fm(x).log_prob(x)
    orders = onets.maf_block_orders(nb, D, 42)
    nets = []
    for bij, order in zip(flow.chain.bijectors[::-1], orders):
        msb = bij.bijector_fn
        for key, net in (('w', msb.bin_widths), ('h', msb.bin_heights), ('s', msb.knot_slopes)):
            masks = onets.made_masks(net.params, D, [H], order)
            new = []
            for k, lay in enumerate(net.layers):
                new += [(rng.normal(0, 0.2, lay.kernel.shape) * masks[k]).astype(np.float32),
                        rng.normal(0, 0.2, lay.units).astype(np.float32)]
            net.set_weights(new)
            nets.append((net, masks))

    def oracle_loss(arrs, ws):
        by_id = {id(w): a for w, a in zip(ws, arrs)}  # (model.weights lists the chain's blocks last-to-first)
        blocks = []
        for b in range(nb):
            blk = {}
            for j, key in enumerate(('w', 'h', 's')):
                net, masks = nets[3 * b + j]
                blk[key] = [dict(W=by_id[id(lay.kernel)] * masks[k], b=by_id[id(lay.bias)], Wc=None, mask=masks[k])
                            for k, lay in enumerate(net.layers)]
            blocks.append(blk)
        xo, il = oflows.maf_inverse(x.astype(np.float64), blocks, K, (-10.0, 10.0))
        return -np.mean(odists.normal_log_prob(xo, 0.0, 1.0).sum(-1) + il)

    loss, ws, grads = _tape_gradients(v, fm, x, x)
    arrs = [w.numpy().astype(np.float64) for w in ws]
    assert abs(loss - oracle_loss(arrs, ws)) < 1e-5 * max(1.0, abs(loss))
    for trial in range(3):
        dirs = [rng.normal(size=a.shape) * (getattr(w, '_grad_mask').numpy() if getattr(w, '_grad_mask', None) is not None else 1.0)
                for a, w in zip(arrs, ws)]
        h = 1e-5
        fd = (oracle_loss([a + h * d for a, d in zip(arrs, dirs)], ws) - oracle_loss([a - h * d for a, d in zip(arrs, dirs)], ws)) / (2 * h)
        got = sum(float((g * d).sum()) for g, d in zip(grads, dirs))
        assert abs(got - fd) <= 1e-4 * max(abs(fd), 1e-2), (got, fd)
    l0 = fm.evaluate(x, x, batch_size=200)
    fm.fit(x, x, epochs=40, batch_size=50)
    assert fm.evaluate(x, x, batch_size=200) < l0 - 0.1


@pytest.mark.parametrize('which', ['blockwise', 'autoregressive', 'autoregressive-conditional', 'maf-flowed'])
def test_vae_with_von_mises_encoder_and_flow_prior_trains(vms, which):
    """Port of tests/test_models.py:189-228 (`test_prior_flow_vary_decoders`): IndependentVonMises encoder over a periodic
    FCDeepNN, RealNVP-RQS prior over a static von Mises, four decoder families; call / sample / log_prob / compile / fit /
    evaluate, and the loss goes down over a few epochs."""
    v = vms
    import vaemolsim_b200._protocols as PR
    d = v.dists
    v.set_seed(11)
    rng = np.random.default_rng(4)
    zdim = 2
    x = np.concatenate([rng.normal(size=(64, 3)), rng.uniform(-3, 3, (64, 3))], axis=1).astype(np.float32)
    kinds = [d.Normal] * 3 + [d.VonMises] * 3
    dec_dist = {
        'blockwise': lambda: d.IndependentBlockwise(6, kinds),
        'autoregressive': lambda: d.AutoregressiveBlockwise(6, kinds),
        'autoregressive-conditional': lambda: d.AutoregressiveBlockwise(6, kinds, conditional=True, conditional_event_shape=2),
        'maf-flowed': lambda: d.FlowedDistribution(v.flows.RQSSplineMAF(rqs_params=dict(hidden_dim=24, num_bins=8)),
                                                   d.IndependentBlockwise(6, kinds)),
    }[which]()
    enc_dist = d.IndependentVonMises(zdim)
    enc_map = v.mappings.FCDeepNN(enc_dist.params_size(zdim), hidden_dim=32, periodic_dofs=[False] * 3 + [True] * 3)
    encoder = v.models.MappingToDistribution(enc_dist, mapping=enc_map, name='encoder')
    decoder = v.models.MappingToDistribution(dec_dist, name='decoder')
    if hasattr(dec_dist, 'flow'):
        dec_dist.flow(np.ones((1, 6), np.float32))
    prior = d.FlowedDistribution(
        v.flows.RQSSplineRealNVP(rqs_params={'bin_range': [-np.pi, np.pi], 'hidden_dim': 24, 'num_bins': 8}),
        PR.DistributionLambda(lambda t: PR.VonMises(np.zeros(zdim, np.float32), np.ones(zdim, np.float32))), name='prior')
    prior.flow(np.ones((1, zdim), np.float32))
    vae = v.models.VAE(encoder, decoder, prior)
    out = vae(x)
    assert isinstance(vae.regularizer, v.losses.KLDivergenceEstimate)
    assert out.sample().shape == (64, 6) and out.log_prob(x).shape == (64, )
    vae.compile(optimizer=v.models.Adam(learning_rate=2e-3), loss=v.losses.LogProbLoss())
    l0 = np.mean([vae.evaluate(x, batch_size=64) for _ in range(4)])
    hist = vae.fit(x, x, epochs=25, batch_size=32)
    assert hist is not None and np.isfinite(hist['loss']).all()
    l1 = np.mean([vae.evaluate(x, batch_size=64) for _ in range(4)])
    assert l1 < l0 - 0.1, (l0, l1, hist['loss'][-3:])


# ------------------------------------------------------------------------------------------------ batch normalisation
def test_batchnorm_backward_matches_float64_finite_differences(vms):
    """vms_batchnorm_backward (TF autodiff through tf.nn.batch_normalization + tf.nn.moments and the tfp bijector's
    log-det): L = sum(G_out * y) + sum_b g_l[b] * ldj, gradients wrt x, gamma, beta, with batch statistics and with constant
    (moving) statistics, against central differences of the float64 oracle."""
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(21)
    B, D, eps = 200, 5, 1e-3
    x = rng.normal(1.0, 2.0, (B, D))
    gamma, beta = rng.uniform(0.5, 1.5, D), rng.normal(size=D)
    G_out, g_l = rng.normal(size=(B, D)), rng.normal(size=B)
    mov_m, mov_v = rng.normal(size=D), rng.uniform(0.5, 2.0, D)

    def L(x_, ga, be, batch):
        m, va = (x_.mean(0), x_.var(0)) if batch else (mov_m, mov_v)
        y, ldj = onets.batch_norm_normalize(x_, m, va, ga, be, eps)
        return float(np.sum(G_out * y) + np.sum(g_l) * ldj)

    f32 = lambda a: v.Tensor.from_numpy(np.ascontiguousarray(a, np.float32))
    for batch in (True, False):
        m, va = (x.mean(0), x.var(0)) if batch else (mov_m, mov_v)
        gx, gg, gb = v.Tensor.zeros((B, D)), v.Tensor.zeros((D, )), v.Tensor.zeros((D, ))
        Gt = f32(np.array([g_l.astype(np.float32).sum()]))
        ws = v.Tensor((int(c.lib.vms_batchnorm_backward_workspace(D)) // 4 + 1, ))
        xd, md, vd, gad, god = f32(x), f32(m), f32(va), f32(gamma), f32(G_out)
        c.lib.vms_batchnorm_backward(xd.ptr, D, B, D, md.ptr, vd.ptr, gad.ptr, eps, 1 if batch else 0, god.ptr, D, Gt.ptr, gx.ptr,
                                     D, gg.ptr, gb.ptr, ws.ptr, c.stream)
        h = 1e-6
        want_x = np.zeros((B, D))
        for b in range(0, B, 17):
            for d in range(D):
                e = np.zeros((B, D)); e[b, d] = h
                want_x[b, d] = (L(x + e, gamma, beta, batch) - L(x - e, gamma, beta, batch)) / (2 * h)
        rows = np.arange(0, B, 17)
        assert_close(gx.numpy()[rows], want_x[rows], rtol=1e-4, atol=1e-4, what='batch norm d/dx (batch stats %s)' % batch)
        wg = np.array([(L(x, gamma + h * np.eye(D)[d], beta, batch) - L(x, gamma - h * np.eye(D)[d], beta, batch)) / (2 * h)
                       for d in range(D)])
        wb = np.array([(L(x, gamma, beta + h * np.eye(D)[d], batch) - L(x, gamma, beta - h * np.eye(D)[d], batch)) / (2 * h)
                       for d in range(D)])
        assert_close(gg.numpy(), wg, rtol=1e-4, atol=2e-3, what='batch norm d/dgamma')
        assert_close(gb.numpy(), wb, rtol=1e-4, atol=2e-3, what='batch norm d/dbeta')


def test_batch_norm_models_train(vms):
    """Port of tests/test_models.py:230-262 (`test_batch_norm`): batch normalisation inside both FCDeepNNs, between the
    blocks of a conditional MAF decoder flow and of the RealNVP prior flow; call / sample / log_prob / compile / fit /
    evaluate.  Plus a finite-difference check of the tape through a batch-normalised decoder in training mode."""
    v = vms
    import vaemolsim_b200._protocols as PR
    d = v.dists
    v.set_seed(13)
    rng = np.random.default_rng(6)
    zdim = 2
    x = np.concatenate([rng.normal(size=(64, 3)), rng.uniform(-3, 3, (64, 3))], axis=1).astype(np.float32)
    kinds = [d.Normal] * 3 + [d.VonMises] * 3
    enc_dist = PR.IndependentNormal(zdim)
    enc_map = v.mappings.FCDeepNN(enc_dist.params_size(zdim), hidden_dim=32, batch_norm=True, periodic_dofs=[False] * 3 + [True] * 3)
    encoder = v.models.MappingToDistribution(enc_dist, mapping=enc_map, name='encoder')
    dec_dist = d.FlowedDistribution(
        v.flows.RQSSplineMAF(rqs_params={'conditional': True, 'conditional_event_shape': zdim, 'hidden_dim': 24, 'num_bins': 8},
                             batch_norm=True), d.IndependentBlockwise(6, kinds))
    dec_dist.flow(np.ones((1, 6), np.float32), conditional_input=np.ones((1, zdim), np.float32))
    dec_map = v.mappings.FCDeepNN(dec_dist.params_size(), hidden_dim=32, batch_norm=True)
    decoder = v.models.MappingToDistribution(dec_dist, mapping=dec_map, name='decoder')
    prior = d.FlowedDistribution(v.flows.RQSSplineRealNVP(batch_norm=True, rqs_params={'hidden_dim': 24, 'num_bins': 8}),
                                 PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], zdim)), name='prior')
    prior.flow(np.ones((1, zdim), np.float32))
    vae = v.models.VAE(encoder, decoder, prior)
    out = vae(x)
    assert out.sample().shape == (64, 6) and out.log_prob(x).shape == (64, )
    vae.compile(optimizer=v.models.Adam(learning_rate=2e-3), loss=v.losses.LogProbLoss())
    hist = vae.fit(x, x, epochs=15, batch_size=32)
    assert hist is not None and np.isfinite(hist['loss']).all() and hist['loss'][-1] < hist['loss'][0]
    assert np.isfinite(vae.evaluate(x, batch_size=64))
    n_bn = sum(1 for w in vae.weights if w.shape == (32, ) or w.shape == (6, ) or w.shape == (zdim, ))
    assert n_bn >= 8  # gamma / beta / moving statistics of the batch-norm layers are model variables
    # tape vs finite differences through a batch-normalised mapping in TRAINING mode (batch statistics)
    dist = d.IndependentBlockwise(4, [d.Normal] * 2 + [d.VonMises] * 2)
    dec = v.models.MappingToDistribution(
        dist, mapping=v.mappings.FCDeepNN(dist.params_size(), hidden_dim=16, activation='tanh', batch_norm=True), name='dec')
    dec.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
    z = rng.normal(size=(96, 3)).astype(np.float32)
    y = np.concatenate([rng.normal(size=(96, 2)), rng.uniform(-3, 3, (96, 2))], axis=1).astype(np.float32)
    dec(z)
    _directional_check(v, dec, z, y, rtol=2e-2, training=True)


def test_graph_replayed_training_steps_equal_eager_steps(vms):
    """The tape trainer captures a whole step (forward, reverse mode, Adam with the step count in device memory) in a CUDA
    graph after two eager steps and replays it: same weights, bit for bit, as eager training; a model whose step draws host
    noise (a VAE's reparameterised sample) falls back to eager steps."""
    v = vms
    import vaemolsim_b200._protocols as PR

    def make():
        v.set_seed(21)
        flow = v.flows.RQSSplineRealNVP(num_blocks=3, rqs_params=dict(bin_range=[-10.0, 10.0], num_bins=16, hidden_dim=32))
        fm = v.models.FlowModel(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 2)))
        fm.compile(optimizer=v.models.Adam(2e-3), loss=v.losses.LogProbLoss())
        return fm

    rng = np.random.default_rng(8)
    xs = [(rng.normal(size=(64, 2)) * 1.5 + 0.3).astype(np.float32) for _ in range(9)]
    a = make()  # (the spline networks draw their weights at the first log_prob, i.e. inside the first step: train a to the
    la = [a.train_on_batch(x, x) for x in xs]  # end before b is seeded and built)
    n0 = v._abi.launch_count()
    la.append(a.train_on_batch(xs[0], xs[0]))
    per_step = v._abi.launch_count() - n0
    a.train_on_batch(xs[0][:32], xs[0][:32])
    b = make()
    b._trainer()._graph_off = True
    lb = [b.train_on_batch(x, x) for x in xs + [xs[0]]]
    b.train_on_batch(xs[0][:32], xs[0][:32])
    tr = a._trainer()
    assert len(getattr(tr, '_graphs', {})) >= 1, 'the step was not captured'
    assert per_step > 20  # the replay accounts for the kernels it launches
    assert la == lb
    assert len(a.weights) == len(b.weights) > 0
    for wa, wb in zip(a.weights, b.weights):  # (includes one step at another batch size: its own capture)
        assert np.array_equal(wa.numpy(), wb.numpy())
    assert len(tr._graphs) == 2
    # host-drawn noise inside the step: no graph, training still works
    d = v.dists
    enc = v.models.MappingToDistribution(PR.IndependentNormal(2), name='encoder')
    dec = v.models.MappingToDistribution(d.IndependentBlockwise(4, [d.Normal] * 4), name='decoder')
    vae = v.models.VAE(enc, dec, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 2)))
    vae.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
    x4 = rng.normal(size=(32, 4)).astype(np.float32)
    m = v.models.Model.train_on_batch
    losses = [m(vae, x4, x4) for _ in range(5)]
    assert np.isfinite(losses).all()
    assert not getattr(vae._trainer(), '_graphs', {})
