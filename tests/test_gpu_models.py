"""GPU parity tests of the host API (the reference's layer / model / MCMC interface) and of the fused ELBO plan against
the CPU oracle, plus ports of the reference's own behavioural tests (shapes, errors, counters)."""
import os

import numpy as np
import pytest

from helpers import assert_close, flat_from_oracle, flat_grad_from_oracle, vae_from_oracle
from oracle import dists as odists
from oracle import flows as oflows
from oracle import mcmc as omc
from oracle import nets as onets
from oracle import vae as ovae

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _std(v, B, D):
    import vaemolsim_b200._protocols as PR
    return PR.DistributionLambda(lambda t: PR.StandardNormal(B if t is None else t.shape[0], D))


# ------------------------------------------------------------------------------------------------ flows
@pytest.mark.parametrize('D', [1, 2, 3, 5])
def test_realnvp_flow_matches_oracle(vms, D):
    v = vms
    rng = np.random.default_rng(D)
    K, H, nb = 8, 16, 4
    flow = v.flows.RQSSplineRealNVP(num_blocks=nb, rqs_params=dict(num_bins=K, hidden_dim=H))
    x = rng.normal(0, 3, (257, D)).astype(np.float32)
    y = flow(x)  # builds
    assert len(flow.chain.bijectors) == nb
    blocks = []
    for bij in flow.chain.bijectors[::-1]:
        sb = bij.bijector_fn
        W, b = sb.heads.get_weights()
        W = W + rng.normal(0, 0.3, W.shape).astype(np.float32)  # make the splines non-trivial
        sb.heads.assign(W, b)
        d1W, d1b = sb.d1.get_weights()
        nw = sb.data_dim * K
        blocks.append(dict(d1=(d1W, d1b), w=(W[:, :nw], b[:nw]), h=(W[:, nw:2 * nw], b[nw:2 * nw]),
                           s=(W[:, 2 * nw:], b[2 * nw:])))
    f64 = [{k: (a.astype(np.float64), c.astype(np.float64)) for k, (a, c) in blk.items()} for blk in blocks]
    yo, flo = oflows.realnvp_forward(x.astype(np.float64), f64, K, (-10.0, 10.0))
    y = flow(x).numpy()
    assert_close(y, yo, rtol=1e-5, atol=1e-5, what='RealNVP forward')
    assert_close(flow.chain.forward_log_det_jacobian(x).numpy(), flo, rtol=1e-5, atol=2e-5, what='RealNVP fldj')
    xb = flow.chain.inverse(y).numpy()
    assert_close(xb, x, rtol=1e-5, atol=5e-5, what='RealNVP round trip')
    # TransformedDistribution.log_prob = base.log_prob(inverse) + ildj  (flows.py:352-353)
    td = flow(_std(v, 257, D)(v.as_tensor(x)))
    xo, ilo = oflows.realnvp_inverse(y.astype(np.float64), f64, K, (-10.0, 10.0))
    want = odists.normal_log_prob(xo, 0.0, 1.0).sum(-1) + ilo
    assert_close(td.log_prob(y).numpy(), want, rtol=1e-5, atol=2e-5, what='flowed log_prob')
    s, lp = td.experimental_sample_and_log_prob()
    assert s.shape == (257, D)
    assert_close(lp.numpy(), td.log_prob(s).numpy(), rtol=1e-4, atol=2e-4, what='sample_and_log_prob consistency')
    # without batch norm training=True is bit-equal to training=False (tests/test_flows.py:176)
    assert np.array_equal(flow(x, training=True).numpy(), flow(x, training=False).numpy())


@pytest.mark.parametrize('D,cond', [(1, 0), (3, 0), (3, 2), (6, 0)])
def test_maf_flow_matches_oracle(vms, D, cond):
    v = vms
    rng = np.random.default_rng(10 * D + cond)
    K, H, nb = 8, 12, 3
    params = dict(num_bins=K, hidden_dim=H)
    if cond:
        params.update(conditional=True, conditional_event_shape=cond)
    flow = v.flows.RQSSplineMAF(num_blocks=nb, order_seed=42, rqs_params=params)
    assert flow.conditional == bool(cond)
    x = rng.normal(0, 2, (130, D)).astype(np.float32)
    cin = rng.normal(size=(130, cond)).astype(np.float32) if cond else None
    kw = dict(conditional_input=cin) if cond else {}
    flow(x, **kw)  # build
    orders = onets.maf_block_orders(nb, D, 42)
    blocks = []
    for bij, order in zip(flow.chain.bijectors[::-1], orders):
        msb = bij.bijector_fn
        blk = {}
        for key, net in (('w', msb.bin_widths), ('h', msb.bin_heights), ('s', msb.knot_slopes)):
            masks = onets.made_masks(net.params, D, [H], order)
            layers = []
            new = []
            for k, lay in enumerate(net.layers):
                W = (rng.normal(0, 0.15, lay.kernel.shape) * masks[k]).astype(np.float32)
                b = rng.normal(0, 0.2, lay.units).astype(np.float32)
                Wc = rng.normal(0, 0.15, (cond, lay.units)).astype(np.float32) if cond else None
                new += [W, b] + ([Wc] if cond else [])
                assert np.array_equal(net.masks[k], masks[k])  # same MADE masks / input orders as the oracle
                layers.append(dict(W=W.astype(np.float64), b=b.astype(np.float64),
                                   Wc=None if Wc is None else Wc.astype(np.float64), mask=masks[k]))
            net.set_weights(new)
            blk[key] = layers
        blocks.append(blk)
    c64 = None if cin is None else cin.astype(np.float64)
    y = flow(x, **kw).numpy()
    yo, flo = oflows.maf_forward(x.astype(np.float64), blocks, K, (-10.0, 10.0), c64)
    assert_close(y, yo, rtol=2e-5, atol=5e-5, what='MAF forward (D passes)')
    xo, ilo = oflows.maf_inverse(y.astype(np.float64), blocks, K, (-10.0, 10.0), c64)
    td = flow(_std(v, 130, D)(v.as_tensor(x)), **kw)
    want = odists.normal_log_prob(xo, 0.0, 1.0).sum(-1) + ilo
    assert_close(td.log_prob(y).numpy(), want, rtol=1e-5, atol=3e-5, what='MAF flowed log_prob')
    assert td.sample().shape == (130, D)
    if cond:
        with pytest.raises(ValueError, match='conditional_input'):
            flow(x)


def test_domain_transform(vms):
    """tests/test_flows.py:15-31."""
    dom = [(-np.pi, np.pi), (0.0, 10.0), (-5.0, 1.0)]
    to = vms.flows.make_domain_transform(dom, (0.0, 1.0))
    back = vms.flows.make_domain_transform(dom, (0.0, 1.0), from_target=True)
    lo = to.forward(np.array([[a for a, _ in dom]], np.float32)).numpy()
    hi = to.forward(np.array([[b for _, b in dom]], np.float32)).numpy()
    np.testing.assert_allclose(lo, 0.0, atol=1e-6)
    np.testing.assert_allclose(hi, 1.0, atol=1e-6)
    x = np.random.default_rng(0).uniform(-3, 1, (10, 3)).astype(np.float32)
    np.testing.assert_allclose(back.forward(to.forward(x)).numpy(), x, atol=1e-5)
    np.testing.assert_allclose(to.inverse(to.forward(x)).numpy(), x, atol=1e-5)
    prm = oflows.domain_transform_params(dom, (0.0, 1.0))
    np.testing.assert_allclose(to.forward_log_det_jacobian(x).numpy(), oflows.domain_transform_fldj(prm), rtol=1e-6)


# ------------------------------------------------------------------------------------------------ dists / losses
def test_distribution_layers_and_losses(vms):
    v = vms
    d, L = v.dists, v.losses
    rng = np.random.default_rng(1)
    B = 100
    # conftest.py fixtures of the reference
    normal_dist = d.Normal(np.linspace(-2.0, 2.0, 5, dtype='float32'), 1.0)
    sample = normal_dist.sample(10)
    assert sample.shape == (10, 5)
    lp = normal_dist.log_prob(sample).numpy()
    assert_close(lp, odists.normal_log_prob(sample.numpy().astype(np.float64), np.linspace(-2, 2, 5), 1.0).sum(-1),
                 rtol=1e-5, atol=1e-5, what='Normal log_prob')
    loss = L.LogProbLoss()(sample, normal_dist)
    assert loss.shape == tuple()
    assert abs(float(loss.numpy()) + lp.astype(np.float64).mean()) < 1e-5
    assert L.LogProbLoss(reduction='none')(sample, normal_dist).shape == (10, )
    # tests/test_losses.py:16-27: two unit Gaussians offset by 1 in 2-D => KL = 1
    dist_a = d.Normal(np.ones((B, 2), np.float32), np.ones((B, 2), np.float32))
    dist_b = d.Normal(np.zeros((B, 2), np.float32), np.ones((B, 2), np.float32))
    kl = L.KLDivergenceEstimate()
    out = kl(dist_a, dist_b)
    assert out.shape == tuple()
    s = dist_a.sample()
    o1 = float(kl(dist_a, dist_b, samples=s).numpy())
    o100 = float(L.KLDivergenceEstimate(weight=100.0)(dist_a, dist_b, samples=s).numpy())
    assert np.float32(o100) == np.float32(100.0) * np.float32(o1)
    sn = s.numpy().astype(np.float64)
    want = np.mean(odists.normal_log_prob(sn, 1.0, 1.0).sum(-1) - odists.normal_log_prob(sn, 0.0, 1.0).sum(-1))
    assert abs(o1 - want) < 1e-5
    rk = float(L.ReverseKLDivergenceEstimate()(dist_a, dist_b, samples=s).numpy())
    assert np.float32(rk) == np.float32(float(kl(dist_b, dist_a, samples=s).numpy()))
    det = d.IndependentDeterministic(2)(np.ones((B, 2), np.float32))
    assert np.allclose(det.sample().numpy(), 1.0)
    o_det = float(kl(det, dist_b).numpy())
    assert abs(o_det + float(np.mean(dist_b.log_prob(np.ones((B, 2), np.float32)).numpy()))) < 1e-6
    lr = float(L.LogProbRegularizer()(dist_a, dist_b, samples=s).numpy())
    assert abs(lr + np.mean(odists.normal_log_prob(sn, 0.0, 1.0).sum(-1))) < 1e-5
    assert L.NonRegularizer()(dist_a, dist_b) == 0.0
    with pytest.raises(NotImplementedError, match='In any subclass'):
        L.InfoRegularizer()(dist_a, dist_b)
    with pytest.raises(ValueError, match='sample_dist'):
        L.InfoRegularizer(sample_dist='dist_c')
    # params_size contracts (tests/test_dists.py:228-242, tests/test_models.py:102-123)
    assert d.IndependentVonMises.params_size(2) == 6 and d.IndependentDeterministic.params_size(3) == 3
    assert d.IndependentBlockwise(3, [d.Normal, d.VonMises, d.Normal]).params_size() == 7
    assert d.AutoregressiveBlockwise(3, d.VonMises).params_size() == (3, 3)
    with pytest.raises(TypeError):
        d.IndependentBlockwise(3, 'Normal')
    with pytest.raises(ValueError):
        d.IndependentBlockwise(3, [d.Normal, d.Normal])
    # von Mises layer: log_prob against the oracle
    p = rng.normal(size=(B, 6)).astype(np.float32)
    vm = d.IndependentVonMises(2)(p)
    xs = vm.sample()
    assert xs.shape == (B, 2)
    assert_close(vm.log_prob(xs).numpy(), odists.independent_vonmises_log_prob(xs.numpy().astype(np.float64),
                                                                               p.astype(np.float64)),
                 rtol=1e-5, atol=1e-5, what='IndependentVonMises log_prob')


def test_autoregressive_blockwise_matches_oracle(vms):
    v = vms
    d = v.dists
    rng = np.random.default_rng(2)
    B, D = 64, 2
    layer = d.AutoregressiveBlockwise(D, d.Normal, conditional=True, conditional_event_shape=1,
                                      auto_net_params={'hidden_units': [10, 100, 10]})
    inputs = rng.normal(size=(B, D, 2)).astype(np.float32)
    cond = rng.normal(size=(B, 1)).astype(np.float32)
    with pytest.raises(ValueError, match='conditional_input'):
        layer(inputs)
    dist = layer(inputs, conditional_input=cond)
    net = layer.auto_net
    masks = onets.made_masks(2, D, [10, 100, 10], 'left-to-right')
    layers, new = [], []
    for k, lay in enumerate(net.layers):
        W = (rng.normal(0, 0.3, lay.kernel.shape) * masks[k]).astype(np.float32)
        b = rng.normal(0, 0.1, lay.units).astype(np.float32)
        Wc = rng.normal(0, 0.3, (1, lay.units)).astype(np.float32)
        new += [W, b, Wc]
        layers.append(dict(W=W.astype(np.float64), b=b.astype(np.float64), Wc=Wc.astype(np.float64), mask=masks[k]))
    net.set_weights(new)
    x = rng.normal(size=(B, D)).astype(np.float32)
    want = odists.autoregressive_blockwise_log_prob(x.astype(np.float64), inputs.astype(np.float64), layers,
                                                    ['normal'] * D, cond.astype(np.float64))
    assert_close(dist.log_prob(x).numpy(), want, rtol=1e-5, atol=1e-5, what='AutoregressiveBlockwise log_prob')
    assert layer.auto_net.count_params() == 2308  # same structure that gives the notebook's 2,332 for 3 parameters
    s = dist.sample()
    assert s.shape == (B, D)
    with pytest.raises(ValueError):
        d.AutoregressiveBlockwise(3, d.Normal)(inputs)


# ------------------------------------------------------------------------------------------------ fused ELBO
@pytest.mark.parametrize('tag,prior', [('c1', 'normal'), ('c2', 'realnvp')])
def test_fused_elbo_matches_golden(vms, tag, prior):
    v = vms
    g = np.load(os.path.join(GOLD, 'elbo_%s.npz' % tag))
    P = ovae.init_vae(1003, prior=prior, flow_hidden=16, num_bins=8, hidden=32)
    model = vae_from_oracle(v, P)
    f = model.fused(256)
    assert f.n_params == g['theta'].size
    assert np.array_equal(f.theta.numpy(), flat_from_oracle(P))
    x, eps = v.as_tensor(g['x']), v.as_tensor(g['eps'])
    out = f.forward(x, eps)
    for k in ('z', 'logq', 'logpz', 'logpx'):
        assert_close(out[k].numpy(), g[k], rtol=1e-5, atol=2e-5, what='%s %s' % (tag, k))
    assert_close(out['scalars'].numpy()[:3], g['scalars'], rtol=1e-5, atol=1e-5, what='scalars')


@pytest.mark.parametrize('prior,dz,B', [('normal', 2, 300), ('realnvp', 2, 300), ('realnvp', 1, 129), ('realnvp', 3, 64),
                                        ('realnvp', 2, 4096)])
def test_fused_elbo_forward_backward_matches_oracle(vms, prior, dz, B):
    v = vms
    P = ovae.init_vae(21 + dz, dx=6, dz=dz, hidden=40, prior=prior, num_blocks=4, num_bins=32 if B > 1000 else 10,
                      flow_hidden=24)
    rng = np.random.default_rng(B)
    if prior == 'realnvp':  # widen the spline heads so bins / slopes vary
        for blk in P['flow']:
            for k in ('w', 'h', 's'):
                blk[k] = ((blk[k][0] * 4).astype(np.float32), rng.normal(0, 0.5, blk[k][1].shape).astype(np.float32))
    x = rng.normal(size=(B, 6)).astype(np.float32)
    eps = rng.normal(size=(B, dz)).astype(np.float32)
    P64 = ovae.cast_params(P, np.float64)
    out, G = ovae.elbo_backward(P64, x.astype(np.float64), eps.astype(np.float64), weight=0.7)
    model = vae_from_oracle(v, P, weight=0.7)
    f = model.fused(B)
    scal = f.forward_backward(v.as_tensor(x), v.as_tensor(eps)).numpy()
    assert_close(scal[:3], [out['loss'], out['nll'], out['kl']], rtol=1e-5, atol=1e-5, what='loss / nll / kl')
    want = flat_grad_from_oracle(P64, G)
    got = f.grad.numpy()
    scale = np.abs(want).max()
    assert_close(got, want, rtol=1e-4, atol=5e-6 * max(scale, 1e-3), what='flat gradient (%s dz=%d)' % (prior, dz))
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert rel < 1e-5, rel
    # graph replay is deterministic
    f.forward_backward(v.as_tensor(x), v.as_tensor(eps))
    assert np.array_equal(f.grad.numpy(), got)


@pytest.mark.parametrize('prior,dz,B,bins', [('realnvp', 2, 4096, 32), ('realnvp', 2, 10007, 32), ('normal', 2, 5000, 8),
                                             ('realnvp', 4, 333, 8), ('realnvp', 1, 77, 20)])
def test_fused_kernel_agrees_with_unfused_plan(vms, prior, dz, B, bins):
    """The single persistent ELBO kernel (elbo_fused.cu) against the per-layer graph path (elbo.cu) on the device:
    same per-row outputs, scalars and flat gradient, including tiles per CTA > 1 and a ragged last tile."""
    v = vms
    P = ovae.init_vae(77 + dz, dx=6, dz=dz, hidden=200 if B > 1000 else 48, prior=prior, num_blocks=4, num_bins=bins,
                      flow_hidden=100 if B > 1000 else 20)
    rng = np.random.default_rng(B)
    if prior == 'realnvp':
        for blk in P['flow']:
            for k in ('w', 'h', 's'):
                blk[k] = ((blk[k][0] * 4).astype(np.float32), rng.normal(0, 0.5, blk[k][1].shape).astype(np.float32))
    x = v.as_tensor(rng.normal(size=(B, 6)).astype(np.float32))
    eps = v.as_tensor(rng.normal(size=(B, dz)).astype(np.float32))
    model = vae_from_oracle(v, P, weight=0.3)
    f = model.fused(B)
    f.set_tc_auto_batch(1 << 40)  # never the large-batch plan in this test
    f.set_mode(4)                 # the single FFMA fused kernel (auto mode would train on the tensor-core fused kernel)
    assert f.is_fused and f.path(B) == 'fused'
    out_f = {k: t.numpy() for k, t in f.forward(x, eps).items()}
    sc_f = f.forward_backward(x, eps).numpy().copy()
    g_f = f.grad.numpy().copy()
    f.forward_backward(x, eps)
    assert np.array_equal(f.grad.numpy(), g_f), 'fused kernel is not deterministic'
    f.set_mode(1)
    assert not f.is_fused
    out_u = {k: t.numpy() for k, t in f.forward(x, eps).items()}
    sc_u = f.forward_backward(x, eps).numpy().copy()
    g_u = f.grad.numpy().copy()
    f.set_mode(0)
    for k in ('z', 'logq', 'logpz', 'logpx'):
        assert_close(out_f[k], out_u[k], rtol=1e-5, atol=2e-5, what='fused vs unfused %s' % k)
    assert_close(out_f['scalars'][:3], out_u['scalars'][:3], rtol=1e-5, atol=1e-5, what='forward scalars')
    assert_close(sc_f[:3], sc_u[:3], rtol=1e-5, atol=1e-5, what='fwd+bwd scalars')
    rel = np.linalg.norm(g_f - g_u) / np.linalg.norm(g_u)
    assert rel < 1e-5, rel  # two float32 evaluations with different summation orders
    assert_close(g_f, g_u, rtol=1e-4, atol=5e-6 * max(np.abs(g_u).max(), 1e-3), what='flat gradient fused vs unfused')


@pytest.mark.parametrize('dz,B,bins,fh,hidden', [(2, 10007, 32, 100, 200), (2, 5000, 20, 40, 64), (1, 700, 8, 16, 32),
                                                 (2, 64, 32, 100, 200)])
def test_tensor_core_plan_matches_oracle_and_ffma_plan(vms, dz, B, bins, fh, hidden):
    """Plan mode 2 -- every RealNVP-RQS coupling block as one tcgen05 kernel (flow_tc.cu: 3 x BF16 split, K-major and
    MN-major reads of the same shared-memory tiles, spline in the epilogue) -- against the float64 oracle (loss scalars,
    flat gradient) and against the float32 FFMA per-layer plan (per-row outputs); ragged last tile, dz = 1 (ones
    conditioner input, flows.py:184-185), both raw-parameter widths (K <= 20: 64 columns, else 96)."""
    v = vms
    P = ovae.init_vae(31 + dz, dx=6, dz=dz, hidden=hidden, prior='realnvp', num_blocks=4, num_bins=bins, flow_hidden=fh)
    rng = np.random.default_rng(B)
    for blk in P['flow']:
        for k in ('w', 'h', 's'):
            blk[k] = ((blk[k][0] * 4).astype(np.float32), rng.normal(0, 0.5, blk[k][1].shape).astype(np.float32))
    x = rng.normal(size=(B, 6)).astype(np.float32)
    eps = rng.normal(size=(B, dz)).astype(np.float32)
    model = vae_from_oracle(v, P, weight=0.7)
    f = model.fused(B)
    xt, et = v.as_tensor(x), v.as_tensor(eps)
    f.set_mode(1)
    out_u = {k: t.numpy() for k, t in f.forward(xt, et).items()}
    f.forward_backward(xt, et)
    g_u = f.grad.numpy().copy()
    f.set_mode(2)
    out_t = {k: t.numpy() for k, t in f.forward(xt, et).items()}
    scal = f.forward_backward(xt, et).numpy().copy()
    g_t = f.grad.numpy().copy()
    f.forward_backward(xt, et)
    assert np.array_equal(f.grad.numpy(), g_t), 'tensor-core plan is not deterministic'
    assert not f.tc_status(), 'a tensor-core completion wait timed out'
    assert f.path(B) == 'tensor-core'
    f.set_mode(0)
    # auto: the whole-step tensor-core kernel for up to three waves of its 32-row tiles (where its shape constraints hold:
    # K a multiple of 4), else the FFMA fused kernel within one wave and the large-batch plan above it
    tcf = bins % 4 == 0 and (B + 31) // 32 <= 3 * 148
    assert f.path(B) == ('tensor-core-fused' if tcf else ('tensor-core' if B > 32 * 148 else 'fused'))
    for k in ('z', 'logq', 'logpz', 'logpx'):
        assert_close(out_t[k], out_u[k], rtol=1e-5, atol=2e-5, what='tensor-core vs FFMA plan %s' % k)
    rel = np.linalg.norm(g_t - g_u) / np.linalg.norm(g_u)
    assert rel < 1e-5, rel
    if B <= 5000:  # float64 oracle (seconds at these sizes)
        P64 = ovae.cast_params(P, np.float64)
        out, G = ovae.elbo_backward(P64, x.astype(np.float64), eps.astype(np.float64), weight=0.7)
        assert_close(scal[:3], [out['loss'], out['nll'], out['kl']], rtol=1e-5, atol=1e-5, what='loss / nll / kl')
        want = flat_grad_from_oracle(P64, G)
        assert_close(g_t, want, rtol=1e-4, atol=5e-6 * max(np.abs(want).max(), 1e-3), what='flat gradient (tensor-core plan)')
        rel = np.linalg.norm(g_t - want) / np.linalg.norm(want)
        assert rel < 1e-5, rel


def test_tensor_core_plan_refuses_unsupported_shapes(vms):
    v = vms
    P = ovae.init_vae(3, dx=6, dz=4, hidden=32, prior='realnvp', num_blocks=2, num_bins=8, flow_hidden=16)  # dt = 2
    f = vae_from_oracle(v, P).fused(128)
    with pytest.raises(Exception):
        f.set_mode(2)


def test_generic_model_path_agrees_with_fused(vms):
    """VAE.call (op-by-op, the reference's code path) and the fused plan evaluate the same ELBO."""
    v = vms
    P = ovae.init_vae(5, prior='realnvp', flow_hidden=16, num_bins=8, hidden=32)
    model = vae_from_oracle(v, P)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(200, 6)).astype(np.float32)
    eps = rng.normal(size=(200, 2)).astype(np.float32)
    f = model.fused(200)
    out = f.forward(v.as_tensor(x), v.as_tensor(eps))
    enc = model.encoder(v.as_tensor(x))
    z, lq = enc.sample_with_noise(v.as_tensor(eps))
    assert_close(z.numpy(), out['z'].numpy(), rtol=1e-6, atol=1e-6, what='z')
    assert_close(lq.numpy(), out['logq'].numpy(), rtol=1e-5, atol=1e-5, what='logq')
    assert_close(model.prior(z).log_prob(z).numpy(), out['logpz'].numpy(), rtol=1e-5, atol=2e-5, what='logpz')
    assert_close(model.decoder(z).log_prob(x).numpy(), out['logpx'].numpy(), rtol=1e-5, atol=2e-5, what='logpx')
    dec = model(x)  # reference-style call: registers the KL loss / metrics (models.py:315-318)
    assert dec.log_prob(x).shape == (200, ) and dec.sample().shape == (200, 6)
    assert set(model.metrics) == {'kl_div', 'regularizer_loss'} and len(model.losses) == 1


def test_training_reduces_loss_and_matches_oracle_adam(vms):
    v = vms
    P = ovae.init_vae(9, prior='realnvp', flow_hidden=16, num_bins=8, hidden=32)
    model = vae_from_oracle(v, P)
    model.compile(optimizer=v.models.Adam(learning_rate=1e-3), loss=v.losses.LogProbLoss())
    rng = np.random.default_rng(1)
    x = rng.normal(size=(256, 6)).astype(np.float32)
    eps = rng.normal(size=(256, 2)).astype(np.float32)
    # three oracle steps
    theta = flat_from_oracle(P).astype(np.float64)
    P64 = ovae.cast_params(P, np.float64)
    first = model.train_step(x, eps)
    out, G = ovae.elbo_backward(P64, x.astype(np.float64), eps.astype(np.float64))
    m, vv = np.zeros_like(theta), np.zeros_like(theta)
    ovae.adam_step(theta, flat_grad_from_oracle(P64, G).astype(np.float64), m, vv, 1)
    assert abs(first['loss'] - out['loss']) < 1e-4 * abs(out['loss'])
    assert_close(model.fused().theta.numpy(), theta, rtol=1e-5, atol=2e-6, what='theta after one Adam step')
    hist = model.fit(x, x, epochs=3, batch_size=64)
    assert hist['loss'][-1] < first['loss']
    assert np.isfinite(model.evaluate(x, x, batch_size=128))
    assert model.predict(x[:10]).shape == (10, 6)


def test_pipelined_train_loop_equals_step_by_step(vms):
    """FusedELBO.train_loop (the loop behind VAE.fit: copy stream, double-buffered inputs, scalar ring read back every 8
    steps) must be the SAME training run as one synchronous train_step per batch: identical per-step scalars and
    bit-identical parameters; 11 steps (ring wrap-around), ragged last batch, with and without a row permutation."""
    v = vms
    rng = np.random.default_rng(4)
    N, bs = 11 * 96 - 40, 96
    x = rng.normal(size=(N, 6)).astype(np.float32)
    eps = rng.normal(size=(N, 2)).astype(np.float32)
    opt = v.models.Adam(learning_rate=1e-3)

    def fresh():
        P = ovae.init_vae(12, prior='realnvp', flow_hidden=16, num_bins=8, hidden=32)
        return vae_from_oracle(v, P).fused(bs)

    f1 = fresh()
    want = []
    for s0 in range(0, N, bs):
        sc = f1.train_step(v.as_tensor(x[s0:s0 + bs]), v.as_tensor(eps[s0:s0 + bs]), opt).numpy()
        want.append(sc[:3].copy())
    f2 = fresh()
    got = f2.train_loop(x, opt, bs, eps_host=eps)
    assert got.shape == (11, 3)
    assert np.array_equal(got, np.array(want)), 'per-step scalars differ'
    assert np.array_equal(f1.theta.numpy(), f2.theta.numpy()), 'parameters differ'
    assert np.array_equal(f2.scalars.numpy()[:3], want[-1])
    # shuffled epoch with device-drawn noise: runs, finite, reproducible for a seed
    order = rng.permutation(N)
    f3, f4 = fresh(), fresh()
    a = f3.train_loop(x, opt, bs, order=order, seed=99)
    b = f4.train_loop(x, opt, bs, order=order, seed=99)
    assert np.all(np.isfinite(a)) and np.array_equal(a, b) and np.array_equal(f3.theta.numpy(), f4.theta.numpy())


def test_standard_normal_stream_continues_across_calls(vms):
    v = vms
    c = v._abi.ctx()
    whole, parts = v.Tensor((1001,)), v.Tensor((1001,))
    c.lib.vms_standard_normal(7, 0, 1001, whole.ptr, c.stream)
    c.lib.vms_standard_normal(7, 0, 333, parts.ptr, c.stream)
    c.lib.vms_standard_normal(7, 333, 668, parts.ptr + 4 * 333, c.stream)
    w = whole.numpy()
    assert np.array_equal(w, parts.numpy())
    assert abs(w.mean()) < 0.15 and abs(w.std() - 1.0) < 0.1


# ------------------------------------------------------------------------------------------------ MCMC
def test_mcmc_counters_and_shapes(vms):
    """tests/test_mcmc.py:34-59."""
    v = vms
    P = ovae.init_vae(3, prior='normal', hidden=32)
    model = vae_from_oracle(v, P)
    x = np.random.default_rng(0).normal(size=(10, 6)).astype(np.float32)
    mc = v.mcmc.MCMC(model, omc.quadratic_energy)
    assert mc._num_trials == 0 and mc._num_acc == 0
    out_x, out_e = mc.single_step(x)
    assert out_x.shape == x.shape and out_e.shape[0] == 10 and mc._num_trials == 10
    out_x, out_e = mc.single_step(out_x, energies=out_e)
    assert mc._num_trials == 20
    mc.reset()
    out_x, out_e = mc.run(x, n_steps=10)
    assert out_x.shape == x.shape and mc._num_trials == 100
    assert mc.acceptance_rate == mc._num_acc / mc._num_trials <= 1.0


@pytest.mark.parametrize('prior', ['normal', 'realnvp'])
def test_mcmc_log_probs_and_decisions_match_oracle(vms, prior):
    """One MC step with injected noise: the six log-probabilities agree with the oracle to 1e-5 and the device
    decisions are bit-exact given those log-probabilities; the end-to-end flip rate is reported."""
    v = vms
    import vaemolsim_b200._protocols as PR
    P = ovae.init_vae(1003, prior=prior, flow_hidden=16, num_bins=8, hidden=32)
    model = vae_from_oracle(v, P)
    B = 2048
    rng = np.random.default_rng(4001)
    x1 = rng.normal(size=(B, 6)).astype(np.float32)
    e1, e2, e3 = (rng.standard_normal((B, d), dtype=np.float32) for d in (2, 2, 6))
    # oracle
    loc, sc, _, _ = ovae.encoder_dist(P, x1)
    z1 = odists.normal_sample(loc, sc, e1)
    lq1 = odists.normal_log_prob(z1, loc, sc).sum(-1)
    z2, lz2 = ovae.prior_sample_and_log_prob(P, e2)
    locx, scx, _, _ = ovae.decoder_dist(P, z2)
    x2 = odists.normal_sample(locx, scx, e3)
    lx2 = odists.normal_log_prob(x2, locx, scx).sum(-1)
    fwd_o = (lq1 + lz2 + lx2).astype(np.float32)
    l2, s2, _, _ = ovae.encoder_dist(P, x2)
    lo1, so1, _, _ = ovae.decoder_dist(P, z1)
    rev_o = (odists.normal_log_prob(z2, l2, s2).sum(-1) + ovae.prior_log_prob(P, z1) +
             odists.normal_log_prob(x1, lo1, so1).sum(-1)).astype(np.float32)
    # device, same noise
    T = v.as_tensor
    dz1, dlq1 = model.encoder(T(x1)).sample_with_noise(T(e1))
    pr = model.prior(dz1)
    if prior == 'normal':
        dz2, dlz2 = T(e2), pr.log_prob(T(e2))
    else:
        y, fldj = pr.bijector._fwd(T(e2))
        dz2, dlz2 = y, pr.distribution.log_prob(T(e2)) - fldj
    dx2, dlx2 = model.decoder(dz2).sample_with_noise(T(e3))
    fwd = dlq1 + dlz2 + dlx2
    rev = model.encoder(dx2).log_prob(dz2) + model.prior(dz2).log_prob(dz1) + model.decoder(dz1).log_prob(T(x1))
    assert_close(dz2.numpy(), z2, rtol=1e-5, atol=2e-5, what='z2')
    assert_close(dx2.numpy(), x2, rtol=1e-5, atol=5e-5, what='x2')
    assert_close(fwd.numpy(), fwd_o, rtol=1e-5, atol=5e-5, what='forward log p')
    assert_close(rev.numpy(), rev_o, rtol=1e-5, atol=5e-5, what='reverse log p')
    e_old, e_new = omc.quadratic_energy(x1), omc.quadratic_energy(x2)
    log_u = np.log(np.random.default_rng(4002).random(B))
    want = omc.accept(e_new, e_old, fwd_o, rev_o, log_u)
    end_to_end = omc.accept(omc.quadratic_energy(dx2.numpy()), e_old, fwd.numpy(), rev.numpy(), log_u)
    flips = int((want != end_to_end).sum())
    print('end-to-end decision flips: %d / %d' % (flips, B))
    assert flips <= max(2, B // 500)


# ------------------------------------------------------------------------------------------------ fused MC kernel
def _mc_noise(seed, n_steps, B, dz, dx):
    """Sampling noise in the order the reference's step draws it (mcmc.py:100-102): encoder, prior, decoder."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_steps, B, 2 * dz + dx), np.float32)
    for s in range(n_steps):
        out[s, :, :dz] = rng.standard_normal((B, dz), dtype=np.float32)
        out[s, :, dz:2 * dz] = rng.standard_normal((B, dz), dtype=np.float32)
        out[s, :, 2 * dz:] = rng.standard_normal((B, dx), dtype=np.float32)
    return out


@pytest.mark.parametrize('lanes', ['1', '8'])
@pytest.mark.parametrize('device_rng', [False, True])
def test_fused_mc_matches_reference_mcmc_py_goldens(vms, device_rng, lanes, monkeypatch):
    """`vms_mc_run` / `vms_mc_run_pcg64` against the decisions of the REFERENCE's own vaemolsim/mcmc.py
    (tests/golden/make_goldens.py): every step restarted from the golden state, same sampling noise, same PCG64 uniform
    stream -- drawn by NumPy on the host (device_rng False) or regenerated on the device by LCG jump-ahead (True).
    Both summation orders of the chain kernel: the 1 / 2 / 4-lane kernel (`lanes` 1) and the 8-lane kernel that shards
    of a multi-GPU job run (`lanes` 8)."""
    v = vms
    monkeypatch.setenv('VMS_MC_TPC', lanes)
    g = np.load(os.path.join(GOLD, 'mcmc_reference_c4a.npz'))
    P = ovae.init_vae(1003, prior='normal', hidden=32)
    model = vae_from_oracle(v, P)
    noise = _mc_noise(777, 5, 64, 2, 6)
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002)
    mc.device_rng = device_rng
    assert mc._fused_plan()['device_rng']  # the C4a shape runs the chain kernel, which carries the device stream
    safe_total = 0
    for s in range(5):
        x_old, e_old = g['x_old_%d' % s], g['e_old_%d' % s]
        x_new, e_new = mc.run_fused(x_old, energies=e_old if s else None, n_steps=1, noise=noise[s:s + 1], trace=True)
        tr = mc._last_trace
        if device_rng:  # bit-identical uniforms; CUDA's double log against np.log: <= 2 ulp
            np.testing.assert_array_max_ulp(tr['log_u'][0], g['log_rand_%d' % s], maxulp=2)
        else:
            assert np.array_equal(tr['log_u'][0], g['log_rand_%d' % s])  # same PCG64 stream as the reference driver
        assert_close(tr['fwd'][0], g['fwd_%d' % s], rtol=1e-5, atol=2e-5, what='forward_log_p step %d' % s)
        assert_close(tr['rev'][0], g['rev_%d' % s], rtol=1e-5, atol=2e-5, what='reverse_log_p step %d' % s)
        assert_close(tr['e_new'][0], g['e_new_%d' % s], rtol=1e-5, atol=1e-5, what='proposal energy step %d' % s)
        # the kernel's acceptance arithmetic is bit-exact given ITS log-probabilities (mcmc.py:116-120, float64)
        want = omc.accept(tr['e_new'][0], e_old, tr['fwd'][0], tr['rev'][0], tr['log_u'][0])
        assert np.array_equal(tr['acc'][0].astype(bool), want)
        # and equals the reference's decision wherever the margin exceeds the float32 log-prob noise
        margin = np.abs(g['e_new_%d' % s] + g['rev_%d' % s] - g['e_old_%d' % s] - g['fwd_%d' % s] - g['log_rand_%d' % s])
        safe = margin > 1e-3
        safe_total += int(safe.sum())
        acc = tr['acc'][0].astype(bool)
        assert np.array_equal(acc[safe], g['acc_%d' % s][safe])
        same = acc == g['acc_%d' % s]
        assert_close(x_new[same], g['configs_%d' % s][same], rtol=1e-5, atol=2e-5, what='configs step %d' % s)
        assert_close(e_new[same], g['energies_%d' % s][same], rtol=1e-5, atol=2e-5, what='energies step %d' % s)
    assert safe_total > 300 and mc._num_trials == 5 * 64
    assert mc.host_stream_reruns == 0
    # the host generator sits where the reference's would after 5 x 64 draws
    want_next = np.random.default_rng(4002)
    want_next.random(size=5 * 64)
    assert mc._rng.random() == want_next.random()


def test_mc_c4b_notebook_model_matches_reference_mcmc_py_goldens(vms):
    """C4b (MC_Moves_with_VAEs.ipynb cell 39-41): MAF prior + conditional autoregressive decoder + Gaussian-mixture energy,
    op-by-op kernels with the chain state on the device, against the decisions of the REFERENCE's own mcmc.py
    (tests/golden/make_goldens.py): every step restarted from the golden state, same sampling noise (host generator, same
    seed and draw order as the oracle twin), same PCG64 uniforms."""
    v = vms
    import vaemolsim_b200._protocols as PR
    from helpers import vae_b_from_oracle
    g = np.load(os.path.join(GOLD, 'mcmc_reference_c4b.npz'))
    P = omc.init_vae_b(2003, hidden=64)
    model = vae_b_from_oracle(v, P)
    energy = v.mcmc.GaussianMixtureEnergy()
    mc = v.mcmc.MCMC(model, energy, random_seed=5002)
    assert mc._fused_plan() is None
    v.set_seed(888)
    safe_total = flips = 0
    for s in range(5):
        x_old, e_old = g['x_old_%d' % s], g['e_old_%d' % s]
        x1 = v.as_tensor(x_old)
        x2, e_out, acc, n_acc = mc._device_step(x1, v.Tensor.from_numpy(e_old))
        assert e_out.dtype == np.float32  # the callback's type, like the reference's returned energies
        acc = acc.numpy().astype(bool)
        margin = np.abs(g['e_new_%d' % s] + g['rev_%d' % s] - g['e_old_%d' % s] - g['fwd_%d' % s] - g['log_rand_%d' % s])
        safe = margin > 2e-3
        safe_total += int(safe.sum())
        flips += int((acc != g['acc_%d' % s]).sum())
        assert np.array_equal(acc[safe], g['acc_%d' % s][safe])
        same = acc == g['acc_%d' % s]
        assert_close(x2.numpy()[same], g['configs_%d' % s][same], rtol=1e-5, atol=5e-5, what='configs step %d' % s)
        assert_close(e_out.numpy()[same], g['energies_%d' % s][same], rtol=1e-5, atol=2e-4, what='energies step %d' % s)
        assert int(n_acc.numpy()[0]) == int(acc.sum())
    assert safe_total > 1200 and flips <= 2
    # whole-loop entry point of the op-by-op path: the state stays on the device and equals n single_steps under the same seeds
    x0 = g['x0']
    a, b = v.mcmc.MCMC(model, energy, random_seed=3), v.mcmc.MCMC(model, energy, random_seed=3)
    a.fuse_notebook = False
    v.set_seed(21)
    xa, ea = a.run(x0, n_steps=4)
    v.set_seed(21)
    xb, eb = x0, None
    for _ in range(4):
        xb, eb = b.single_step(xb, energies=eb)
    assert np.array_equal(xa, xb) and np.array_equal(ea, eb) and a._num_acc == b._num_acc and a._num_trials == 4 * 256
    assert ea.dtype == np.float32 and 0.0 < a.acceptance_rate < 1.0
    # a host callback (the reference's protocol) gives the same chain as the device energy, up to float32 log/exp rounding
    c = v.mcmc.MCMC(model, omc.gmm_energy, random_seed=3)
    v.set_seed(21)
    xc, ec = c.run(x0, n_steps=4)
    assert np.mean(np.all(xc == xa, axis=1)) > 0.99


def _nb_noise(seed, n_steps, B):
    """Sampling noise of the notebook family in the reference's draw order (mcmc.py:100-102): z1 [1], z2 [1], x2 [2]."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_steps, B, 4), np.float32)
    for s in range(n_steps):
        out[s, :, 0:1] = rng.standard_normal((B, 1), dtype=np.float32)
        out[s, :, 1:2] = rng.standard_normal((B, 1), dtype=np.float32)
        out[s, :, 2:4] = rng.standard_normal((B, 2), dtype=np.float32)
    return out


@pytest.mark.parametrize('device_rng', [False, True])
def test_fused_mc_notebook_kernel_matches_reference_mcmc_py_goldens(vms, device_rng):
    """`vms_mc_nb_run` (the fused kernel of the MC notebook's model family, C4b) against the decisions of the REFERENCE's own
    mcmc.py: every step restarted from the golden state, injected sampling noise, PCG64 uniforms from the host or
    regenerated on the device."""
    v = vms
    from helpers import vae_b_from_oracle
    g = np.load(os.path.join(GOLD, 'mcmc_reference_c4b.npz'))
    P = omc.init_vae_b(2003, hidden=64)
    model = vae_b_from_oracle(v, P)
    mc = v.mcmc.MCMC(model, v.mcmc.GaussianMixtureEnergy(), random_seed=5002)
    mc.device_rng = device_rng
    assert mc._fused_plan() is None and mc._nb_plan() is not None
    noise = _nb_noise(888, 5, 256)
    safe_total = flips = 0
    for s in range(5):
        x_old, e_old = g['x_old_%d' % s], g['e_old_%d' % s]
        x_new, e_new = mc.run_nb(x_old, energies=e_old if s else None, n_steps=1, noise=noise[s:s + 1], trace=True)
        tr = mc._last_trace
        if device_rng:
            np.testing.assert_array_max_ulp(tr['log_u'][0], g['log_rand_%d' % s], maxulp=2)
        else:
            assert np.array_equal(tr['log_u'][0], g['log_rand_%d' % s])
        assert_close(tr['fwd'][0], g['fwd_%d' % s], rtol=1e-5, atol=5e-5, what='forward_log_p step %d' % s)
        assert_close(tr['rev'][0], g['rev_%d' % s], rtol=1e-5, atol=2e-4, what='reverse_log_p step %d' % s)
        assert_close(tr['e_new'][0], g['e_new_%d' % s], rtol=1e-5, atol=2e-4, what='proposal energy step %d' % s)
        # the kernel's acceptance arithmetic is NumPy's float32 evaluation given ITS log-probabilities (mcmc.py:116-120)
        la = (tr['e_new'][0] + tr['rev'][0] - e_old - tr['fwd'][0])
        assert la.dtype == np.float32 and np.array_equal(tr['acc'][0].astype(bool), la >= tr['log_u'][0])
        margin = np.abs(g['e_new_%d' % s] + g['rev_%d' % s] - g['e_old_%d' % s] - g['fwd_%d' % s] - g['log_rand_%d' % s])
        safe = margin > 2e-3
        safe_total += int(safe.sum())
        acc = tr['acc'][0].astype(bool)
        flips += int((acc != g['acc_%d' % s]).sum())
        assert np.array_equal(acc[safe], g['acc_%d' % s][safe])
        same = acc == g['acc_%d' % s]
        assert_close(x_new[same], g['configs_%d' % s][same], rtol=1e-5, atol=5e-5, what='configs step %d' % s)
        assert_close(e_new[same], g['energies_%d' % s][same], rtol=1e-5, atol=2e-4, what='energies step %d' % s)
    assert safe_total > 1200 and flips <= 2 and mc._num_trials == 5 * 256 and mc.host_stream_reruns == 0
    assert mc._num_acc == sum(int(g['acc_%d' % s].sum()) for s in range(5)) or flips > 0
    want_next = np.random.default_rng(5002)
    want_next.random(size=5 * 256)
    assert mc._rng.random() == want_next.random()


def test_fused_mc_notebook_kernel_equals_op_by_op_path_over_many_steps(vms):
    """The fused notebook-family kernel against the op-by-op path on 3000 chains x 12 steps with the same sampling noise and
    uniforms (a tanh MADE this time): same decision trace except where a decision sits within float32 rounding of its
    threshold, multi-step launches equal repeated single-step launches bit for bit, shards equal the whole."""
    v = vms
    import vaemolsim_b200._protocols as PR
    from helpers import vae_b_from_oracle
    P = omc.init_vae_b(77, hidden=200)
    model = vae_b_from_oracle(v, P, made_activation='tanh')
    energy = v.mcmc.GaussianMixtureEnergy()
    B, n_steps = 3000, 12
    rng = np.random.default_rng(5)
    k = rng.choice(3, size=B, p=[0.7, 0.2, 0.1])
    x0 = (omc.GMM_LOCS[k] + omc.GMM_SCALES[k] * rng.standard_normal((B, 2))).astype(np.float32)
    noise = _nb_noise(31, n_steps, B)
    a = v.mcmc.MCMC(model, energy, random_seed=9)
    xa, ea = a.run_nb(x0, n_steps=n_steps, noise=noise, trace=True)
    tra = a._last_trace
    # op-by-op: host generator replays the same noise
    b = v.mcmc.MCMC(model, energy, random_seed=9)
    b.fuse_notebook = False
    v.set_seed(31)
    xb, eb = x0, None
    acc_b = np.empty((n_steps, B), bool)
    for s in range(n_steps):
        xb, eb = b.single_step(xb, energies=eb)
        acc_b[s] = b._last_acc.numpy().astype(bool)
    diff_chains = (tra['acc'].astype(bool) != acc_b).any(axis=0)
    assert diff_chains.sum() <= 3, diff_chains.sum()
    ok = ~diff_chains
    assert_close(xa[ok], xb[ok], rtol=1e-5, atol=5e-5, what='final configs')
    assert_close(ea[ok], eb[ok], rtol=1e-5, atol=2e-4, what='final energies')
    assert 0.02 < a.acceptance_rate < 0.9
    # one launch of 12 steps == 12 launches of one step (device rng: the stream position carries over)
    c = v.mcmc.MCMC(model, energy, random_seed=9)
    xc, ec = x0, None
    for s in range(n_steps):
        xc, ec = c.run_nb(xc, energies=ec, n_steps=1, noise=noise[s:s + 1])
    assert np.array_equal(xc, xa) and np.array_equal(ec, ea) and c._num_acc == a._num_acc
    # host-stream uniforms give the same chain as the device stream
    d = v.mcmc.MCMC(model, energy, random_seed=9)
    d.device_rng = False
    xd, ed_ = d.run_nb(x0, n_steps=n_steps, noise=noise)
    assert np.array_equal(xd, xa) and np.array_equal(ed_, ea) and d._rng.random() == a._rng.random()
    # a shard (chain0, n_global) of the chain set reproduces its rows of the whole run, with the device noise stream too
    w = v.mcmc.MCMC(model, energy, random_seed=9)
    xw, ew = w.run(x0, n_steps=n_steps)
    lo, hi = 1000, 2200
    sh = v.mcmc.MCMC(model, energy, random_seed=9, stream_layout=(lo, B))
    xs, es = sh.run(x0[lo:hi], n_steps=n_steps)
    assert np.array_equal(xs, xw[lo:hi]) and np.array_equal(es, ew[lo:hi])


def test_fused_mc_chain_kernel_lane_counts_are_bitwise_equal(vms, monkeypatch):
    """C4a kernel (mc_chain.cu): 1, 2 or 4 lanes per chain walk the same four hidden-unit streams -- identical states,
    energies, log-probability and decision traces (hidden = 50: zero-row padding to a multiple of 4)."""
    v = vms
    P = ovae.init_vae(13, prior='normal', hidden=50)
    model = vae_from_oracle(v, P)
    B, n_steps = 1500, 10
    x0 = np.random.default_rng(6).normal(size=(B, 6)).astype(np.float32)
    res = {}
    for tpc in ('1', '2', '4'):
        monkeypatch.setenv('VMS_MC_TPC', tpc)
        mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=12)
        x, e = mc.run_fused(x0, n_steps=n_steps, trace=True)
        res[tpc] = (x, e, mc._last_trace, mc._num_acc)
    for tpc in ('2', '4'):
        assert np.array_equal(res[tpc][0], res['1'][0]) and np.array_equal(res[tpc][1], res['1'][1])
        for key in ('acc', 'fwd', 'rev', 'e_new', 'log_u'):
            assert np.array_equal(res[tpc][2][key], res['1'][2][key]), (tpc, key)
        assert res[tpc][3] == res['1'][3]
    assert 0 < res['1'][3] < B * n_steps
    # two chains per lane group (what a full GPU runs: one lane x two chains; a half-full one: two lanes x two chains) share
    # every weight load and nothing else: bit-identical again (1,500 chains: the last group is half empty)
    for tpc in ('1', '2', '4'):
        monkeypatch.setenv('VMS_MC_TPC', tpc)
        monkeypatch.setenv('VMS_MC_CPL', '2')
        mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=12)
        x, e = mc.run_fused(x0[:B - 1], n_steps=n_steps, trace=True)   # odd chain count
        monkeypatch.setenv('VMS_MC_CPL', '1')
        mc1 = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=12)
        x1c, e1c = mc1.run_fused(x0[:B - 1], n_steps=n_steps, trace=True)
        assert np.array_equal(x, x1c) and np.array_equal(e, e1c) and mc._num_acc == mc1._num_acc, tpc
        for key in ('acc', 'fwd', 'rev', 'e_new', 'log_u'):
            assert np.array_equal(mc._last_trace[key], mc1._last_trace[key]), (tpc, key)
    monkeypatch.delenv('VMS_MC_CPL')
    # the 8-lane kernel sums the hidden units in its own order: log-probabilities to float32 rounding, the same decision
    # wherever the margin exceeds that rounding, then identical chains
    monkeypatch.setenv('VMS_MC_TPC', '8')
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=12)
    x8, e8 = mc.run_fused(x0, n_steps=1, trace=True)
    t8 = mc._last_trace
    monkeypatch.setenv('VMS_MC_TPC', '1')
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=12)
    x1_, e1_ = mc.run_fused(x0, n_steps=1, trace=True)
    t1 = mc._last_trace
    assert np.array_equal(t8['log_u'], t1['log_u'])
    assert_close(t8['fwd'], t1['fwd'], rtol=1e-5, atol=2e-5, what='8-lane forward_log_p')
    assert_close(t8['rev'], t1['rev'], rtol=1e-5, atol=2e-5, what='8-lane reverse_log_p')
    assert_close(t8['e_new'], t1['e_new'], rtol=1e-5, atol=1e-5, what='8-lane proposal energy')
    e_old = np.zeros(B)
    for d in range(6):  # the kernel's float64 order (tests/test_mcmc.py:28-32 on six coordinates)
        e_old = e_old + (x0[:, d].astype(np.float64) - np.linspace(-2, 2, 6)[d]) ** 2
    want = omc.accept(t8['e_new'][0], e_old, t8['fwd'][0], t8['rev'][0], t8['log_u'][0])
    assert np.array_equal(t8['acc'][0].astype(bool), want)  # bit-exact acceptance arithmetic given ITS log-probabilities
    same = t8['acc'][0] == t1['acc'][0]
    assert same.sum() >= B - 2
    assert_close(x8[same], x1_[same], rtol=1e-5, atol=2e-5, what='8-lane configs')


def test_fused_mc_notebook_kernel_lane_counts_are_bitwise_equal(vms, monkeypatch):
    """The launcher picks 1 or 4 lanes per chain by the number of chains; all lane counts walk the same four unit streams,
    so states, energies and decision traces are bit-identical (a shard of a multi-GPU job equals its rows of the whole)."""
    v = vms
    from helpers import vae_b_from_oracle
    P = omc.init_vae_b(78, hidden=50)   # 50 hidden units: not a multiple of 4 (zero-row padding)
    model = vae_b_from_oracle(v, P, made_activation='relu')
    energy = v.mcmc.GaussianMixtureEnergy()
    B, n_steps = 1500, 10
    rng = np.random.default_rng(6)
    k = rng.choice(3, size=B, p=[0.7, 0.2, 0.1])
    x0 = (omc.GMM_LOCS[k] + omc.GMM_SCALES[k] * rng.standard_normal((B, 2))).astype(np.float32)
    res = {}
    for tpc in ('1', '2', '4'):
        monkeypatch.setenv('VMS_NB_TPC', tpc)
        mc = v.mcmc.MCMC(model, energy, random_seed=12)
        x, e = mc.run_nb(x0, n_steps=n_steps, trace=True)
        res[tpc] = (x, e, mc._last_trace, mc._num_acc)
    for tpc in ('2', '4'):
        assert np.array_equal(res[tpc][0], res['1'][0]) and np.array_equal(res[tpc][1], res['1'][1])
        for key in ('acc', 'fwd', 'rev', 'e_new', 'log_u'):
            assert np.array_equal(res[tpc][2][key], res['1'][2][key]), (tpc, key)
        assert res[tpc][3] == res['1'][3]
    assert 0 < res['1'][3] < B * n_steps


def test_device_pcg64_stream_equals_host_stream(vms):
    """The device-drawn accept uniforms (`vms_mc_run_pcg64`) reproduce NumPy's PCG64 stream: same decisions / final state
    as the host-stream path over many steps, for a whole chain set and for a shard (chain0, n_global) of it; the
    uncertainty fallback re-runs a call on the host stream."""
    v = vms
    P = ovae.init_vae(11, prior='normal', hidden=200)
    model = vae_from_oracle(v, P)
    B, n_steps = 3000, 40
    x0 = np.random.default_rng(5).normal(size=(B, 6)).astype(np.float32)
    energy = v.mcmc.QuadraticEnergy(6)
    a, b = v.mcmc.MCMC(model, energy, random_seed=9), v.mcmc.MCMC(model, energy, random_seed=9)
    b.device_rng = False
    xa, ea = a.run_fused(x0, n_steps=n_steps, trace=True)
    tra = a._last_trace
    xb, eb = b.run_fused(x0, n_steps=n_steps, trace=True)
    trb = b._last_trace
    u = np.random.default_rng(9).random(size=(n_steps, B))
    np.testing.assert_array_max_ulp(tra['log_u'], np.log(u), maxulp=2)
    assert np.array_equal(trb['log_u'], np.log(u))
    assert np.array_equal(tra['acc'], trb['acc']) and np.array_equal(xa, xb) and np.array_equal(ea, eb)
    assert a._num_acc == b._num_acc and a._rng.random() == b._rng.random()
    # a shard of a global chain set draws ITS columns of the global stream
    lo, hi = 1000, 2200
    c1 = v.mcmc.MCMC(model, energy, random_seed=9, stream_layout=(lo, B))
    xc, ec = c1.run(x0[lo:hi], n_steps=n_steps)
    assert np.array_equal(xc, xa[lo:hi]) and np.array_equal(ec, ea[lo:hi])
    c2 = v.mcmc.MCMC(model, energy, random_seed=9, stream_layout=(lo, B))
    c2.device_rng = False
    xd, ed = c2.run(x0[lo:hi], n_steps=n_steps)
    assert np.array_equal(xd, xc) and c2._rng.random() == c1._rng.random()
    # two calls continue the stream exactly like one
    d = v.mcmc.MCMC(model, energy, random_seed=9)
    x1, e1 = d.run(x0, n_steps=15)
    x2, e2 = d.run(x1, energies=e1, n_steps=n_steps - 15)
    assert np.array_equal(x2, xa) and np.array_equal(e2, ea) and d._num_acc == a._num_acc
    # the fallback: pretend the device flagged an uncertain decision -> the call is repeated on the host stream
    f = v.mcmc.MCMC(model, energy, random_seed=9)
    fp = f._fused_plan()
    one = np.array([1], np.uint64)

    class Poisoned(v.Tensor):
        def fill_zero(self):
            c = v._abi.ctx()
            c.lib.vms_memcpy_h2d(self.ptr, one.ctypes.data, 8, c.stream)
            c.synchronize()
            return self

    pois = Poisoned((1, ), np.uint64)
    fp['n_unc'] = pois
    xf, ef = f.run_fused(x0, n_steps=n_steps)
    assert f.host_stream_reruns == 1
    assert np.array_equal(xf, xa) and np.array_equal(ef, ea) and f._num_acc == a._num_acc
    assert f._rng.random() == np.random.default_rng(9).random(size=n_steps * B + 1)[-1]


def test_fused_mc_multi_step_equals_single_steps_and_op_by_op_path(vms):
    """100 steps in one launch == 100 launches of one step (same noise / uniforms), ragged last tile included; the
    op-by-op MCMC path (per-layer kernels, vms_mc_accept) reproduces the same log-probabilities."""
    v = vms
    P = ovae.init_vae(11, prior='normal', hidden=200)
    model = vae_from_oracle(v, P)
    B, n_steps = 1000, 20
    x0 = np.random.default_rng(5).normal(size=(B, 6)).astype(np.float32)
    noise = _mc_noise(99, n_steps, B, 2, 6)
    energy = v.mcmc.QuadraticEnergy(6)
    a = v.mcmc.MCMC(model, energy, random_seed=1)
    xa, ea = a.run_fused(x0, n_steps=n_steps, noise=noise, trace=True)
    tra = a._last_trace
    b = v.mcmc.MCMC(model, energy, random_seed=1)
    xb, eb = x0, None
    for s in range(n_steps):
        xb, eb = b.run_fused(xb, energies=eb, n_steps=1, noise=noise[s:s + 1], trace=True)
        assert np.array_equal(b._last_trace['acc'][0], tra['acc'][s])
    assert np.array_equal(xa, xb) and np.array_equal(ea, eb)
    assert a._num_acc == b._num_acc == float(tra['acc'].sum()) and a._num_trials == B * n_steps
    assert np.array_equal(ea, energy(xa))  # carried energies are the energies of the carried configurations
    # device RNG path: deterministic for a seed, independent of how the chains are split into calls
    c1 = v.mcmc.MCMC(model, energy, random_seed=3)
    c2 = v.mcmc.MCMC(model, energy, random_seed=3)
    x1, e1 = c1.run(x0, n_steps=5)
    assert c1._fused_plan() is not None
    lo = np.log(np.random.default_rng(3).random(size=(5, B)))
    x2a, _ = c2.run_fused(x0[:600], n_steps=5, log_u_dev=v.Tensor.from_numpy(np.ascontiguousarray(lo[:, :600])))
    assert np.array_equal(x1[:600], x2a)
    assert 0.0 < c1.acceptance_rate <= 1.0
    # long runs through the public API are pipelined in chunks of 10 steps (host PCG64 || device): same result as one launch
    c3, c4 = v.mcmc.MCMC(model, energy, random_seed=3), v.mcmc.MCMC(model, energy, random_seed=3)
    c3.device_rng = False
    x3, e3 = c3.run(x0, n_steps=25)
    lo25 = np.log(np.random.default_rng(3).random(size=(25, B)))
    x4, e4 = c4.run_fused(x0, n_steps=25, log_u_dev=v.Tensor.from_numpy(lo25))
    assert np.array_equal(x3, x4) and np.array_equal(e3, e4) and c3._num_acc == c4._num_acc
    # op-by-op path on the same first step: same six log-probabilities
    z1, lq1 = model.encoder(v.as_tensor(x0)).sample_with_noise(v.as_tensor(noise[0, :, :2]))
    assert_close(lq1.numpy() + model.prior(z1).log_prob(v.as_tensor(noise[0, :, 2:4])).numpy() +
                 model.decoder(v.as_tensor(noise[0, :, 2:4])).sample_with_noise(v.as_tensor(noise[0, :, 4:]))[1].numpy(),
                 tra['fwd'][0], rtol=1e-5, atol=3e-5, what='forward_log_p vs op-by-op path')


# ------------------------------------------------------------------------------------------------ batch normalisation
def test_batch_norm_layer_and_bijector_match_restatement(vms):
    """tf.keras.layers.BatchNormalization (mappings.py:113-114) and tfp.bijectors.BatchNormalization (flows.py:308-309)
    on the device (csrc/batchnorm.cu) against the NumPy restatement: inference and training statistics, moving-average
    update, both directions of the bijector and its log-dets."""
    from oracle import nets as onets
    v = vms
    PR = v._protocols
    rng = np.random.default_rng(3)
    B, D = 3001, 5
    x = (rng.normal(size=(B, D)) * [1.0, 2.0, 0.5, 3.0, 1.5] + [0.5, -1.0, 2.0, 0.0, 4.0]).astype(np.float32)
    gamma, beta = rng.uniform(0.5, 2.0, D).astype(np.float32), rng.normal(size=D).astype(np.float32)
    lay = PR.KerasBatchNormalization()
    out0 = lay(v.as_tensor(x)).numpy()  # inference with the initial moving statistics (0, 1)
    assert_close(out0, x / np.sqrt(1.0 + 1e-3), rtol=1e-6, atol=1e-6, what='BN inference, initial statistics')
    lay.state.gamma, lay.state.beta = v.Tensor.from_numpy(gamma), v.Tensor.from_numpy(beta)
    mean, var = onets.batch_norm_moments(x.astype(np.float64))
    out1 = lay(v.as_tensor(x), training=True).numpy()
    want, _ = onets.batch_norm_normalize(x.astype(np.float64), mean, var, gamma, beta)
    assert_close(out1, want, rtol=1e-5, atol=1e-5, what='BN training output')
    assert_close(lay.state.moving_mean.numpy(), 0.01 * mean, rtol=1e-5, atol=1e-6, what='moving mean')
    assert_close(lay.state.moving_variance.numpy(), 0.99 + 0.01 * var, rtol=1e-5, atol=1e-6, what='moving variance')
    assert lay.count_params() == 4 * D
    # bijector: inverse = normalisation (batch statistics when training), forward = de-normalisation (moving statistics)
    bij = PR.BatchNormalization(training=True)
    y, ildj = bij._inv(v.as_tensor(x))
    want, want_ldj = onets.batch_norm_normalize(x.astype(np.float64), mean, var, np.ones(D), np.zeros(D))
    assert_close(y.numpy(), want, rtol=1e-5, atol=1e-5, what='bijector inverse (training)')
    assert_close(ildj.numpy(), np.full(B, want_ldj), rtol=1e-5, atol=1e-5, what='bijector ildj')
    bij.training = False
    mm, mv = bij.state.moving_mean.numpy().astype(np.float64), bij.state.moving_variance.numpy().astype(np.float64)
    z, fldj = bij._fwd(v.as_tensor(x))
    want, want_ldj = onets.batch_norm_denormalize(x.astype(np.float64), mm, mv, np.ones(D), np.zeros(D))
    assert_close(z.numpy(), want, rtol=1e-5, atol=1e-5, what='bijector forward')
    assert_close(fldj.numpy(), np.full(B, want_ldj), rtol=1e-5, atol=1e-5, what='bijector fldj')
    back = bij.inverse(z).numpy()
    assert_close(back, x, rtol=1e-5, atol=1e-5, what='bijector round trip')


def test_batch_norm_variants_of_the_reference_layers(vms):
    """Structure checks of the reference's own tests (tests/test_mappings.py:24-27, tests/test_flows.py:178-196,
    tests/test_dists.py:172-189): layer counts, `training` plumbing, shapes."""
    v = vms
    PR = v._protocols
    rng = np.random.default_rng(5)
    x = rng.normal(size=(200, 6)).astype(np.float32)
    nn = v.mappings.FCDeepNN(3, batch_norm=True)
    out = nn(x)
    assert len(nn.layer_list) == len(nn.hidden_dim) * 2 + 2 and out.shape == (200, 3)
    assert not np.array_equal(nn(x, training=True).numpy(), out.numpy())
    plain = v.mappings.FCDeepNN(3)
    plain(x)
    assert len(plain.layer_list) == len(plain.hidden_dim) + 2
    for f_class in (v.flows.RQSSplineRealNVP, v.flows.RQSSplineMAF):
        f = f_class(num_blocks=4, batch_norm=True)
        y0 = f(x)
        bn = [b for b in f.chain.bijectors if isinstance(b, PR.BatchNormalization)]
        assert len(f.chain.bijectors) == 2 * f.num_blocks - 1 and len(bn) == 3 and not any(b.training for b in bn)
        f(x, training=True)
        assert all(b.training for b in bn)
        assert y0.shape == x.shape and not np.array_equal(y0.numpy(), x)
        f(x, training=False)
        base = PR.StandardNormal(200, 6)
        lp = f(base).log_prob(y0).numpy()
        assert lp.shape == (200,) and np.all(np.isfinite(lp))
        assert_close(f.chain.inverse(f.chain.forward(v.as_tensor(x))).numpy(), x, rtol=1e-4, atol=1e-4,
                     what='flow with batch norm: round trip')


def test_keras_style_weight_export_import_round_trip(vms, tmp_path):
    """`get_weights` / `set_weights` / `save_weights` / `load_weights` (the Keras variable-order protocol the reference's
    users move weights with): a second, independently initialised model reproduces the first one's outputs after the
    transfer, MADE kernels stay masked, and the variable count is the notebook's (Training_VAEs_and_Decoders cell 43: 3,938
    parameters for FCDeepNN 1 -> 200 -> (2, 3) + conditional MADE [10, 100, 10])."""
    v = vms
    d = v.dists

    def make(seed):
        v.set_seed(seed)
        dist = d.AutoregressiveBlockwise(2, [d.Normal, d.VonMises], conditional=True, conditional_event_shape=1,
                                         auto_net_params={'hidden_units': [10, 100, 10]})
        m = v.models.MappingToDistribution(dist, name='decoder')
        m(np.zeros((2, 1), np.float32))
        return m

    a, b = make(1), make(2)
    assert a.count_params() == 3938
    z = np.random.default_rng(0).normal(size=(50, 1)).astype(np.float32)
    x = np.random.default_rng(1).uniform(-2, 2, (50, 2)).astype(np.float32)
    la = a(z).log_prob(x).numpy()
    assert not np.allclose(la, b(z).log_prob(x).numpy())
    path = str(tmp_path / 'decoder_weights.npz')
    a.save_weights(path)
    b.load_weights(path)
    assert np.array_equal(b(z).log_prob(x).numpy(), la)
    ws = b.get_weights()
    assert [w.shape for w in ws] == [w.shape for w in a.get_weights()] and sum(w.size for w in ws) == 3938
    for w in b._unique_weights():
        m = getattr(w, '_grad_mask', None)
        if m is not None:
            assert np.all(w.numpy()[m.numpy() == 0] == 0)
    with pytest.raises(ValueError):
        b.set_weights(ws[:-1])


@pytest.mark.parametrize('made_hidden,order,act,bins,blocks,hidden', [
    ([10, 100, 10], 'left-to-right', None, 20, 4, 200),
    ([10, 100, 10], 'right-to-left', 'tanh', 20, 2, 64),
    ([7, 33, 12], 'left-to-right', 'relu', 8, 1, 37),
    ([16, 64, 16], 'right-to-left', None, 12, 3, 50),
    ([3, 5, 2], 'left-to-right', 'tanh', 4, 2, 8),
])
def test_fused_mc_notebook_kernel_shapes_orders_activations(vms, monkeypatch, made_hidden, order, act, bins, blocks, hidden):
    """The fused notebook-family kernel over the shapes it accepts -- padded MADE widths (12 / 16), both input orders (which
    dof the masks put first), the three activations, 1-4 flow blocks, hidden sizes that are no multiple of 4 -- against the
    op-by-op kernels on the same noise and uniforms; and the generic D + 1-pass sampling path (no mask knowledge) against
    the one-pass path, bit for bit."""
    v = vms
    import vaemolsim_b200._protocols as PR
    d = v.dists
    v.set_seed(17)
    enc = v.models.MappingToDistribution(PR.IndependentNormal(1), name='encoder')
    dec_dist = d.AutoregressiveBlockwise(2, [d.Normal] * 2, conditional=True, conditional_event_shape=(1, ),
                                         auto_net_params={'hidden_units': made_hidden, 'activation': act, 'input_order': order})
    dec = v.models.MappingToDistribution(dec_dist, name='decoder')
    enc.mapping.hidden_dim = [hidden]
    dec.mapping.hidden_dim = [hidden + 3]
    flow = v.flows.RQSSplineMAF(num_blocks=blocks, rqs_params={'bin_range': [-6.0, 6.0], 'num_bins': bins, 'hidden_dim': 9})
    flow(np.zeros((2, 1), np.float32))
    prior = d.FlowedDistribution(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 1)), name='prior')
    model = v.models.VAE(enc, dec, prior)
    model(np.zeros((2, 2), np.float32))
    rng = np.random.default_rng(2)
    # biases away from zero (a zero-bias 1-D MAF is the identity), the MADE's shift kept moderate
    net = dec_dist.auto_net
    arrs = []
    for k, lay in enumerate(net.layers):
        sc = np.float32(0.3 if k == len(net.layers) - 1 else 1.0)
        arrs += [lay.kernel.numpy() * sc, rng.normal(0, 0.3, lay.units).astype(np.float32) * sc, net.cond_kernels[k].numpy() * sc]
    net.set_weights(arrs)
    for bij in flow.chain.bijectors:
        msb = bij.bijector_fn
        for sub in (msb.bin_widths, msb.bin_heights, msb.knot_slopes):
            sub.set_weights([a for lay in sub.layers for a in (lay.kernel.numpy(), rng.normal(0, 0.5, lay.units).astype(np.float32))])
    energy = v.mcmc.GaussianMixtureEnergy(probs=(0.5, 0.5), locs=((0.0, 0.0), (1.0, -1.0)), scales=((1.0, 1.0), (0.7, 1.3)))
    B, n_steps = 777, 6
    x0 = rng.normal(size=(B, 2)).astype(np.float32)
    noise = _nb_noise(5, n_steps, B)
    a = v.mcmc.MCMC(model, energy, random_seed=4)
    assert a._nb_plan() is not None and a._nb_plan()['model'].made_first_dof == (0 if order == 'left-to-right' else 1)
    xa, ea = a.run_nb(x0, n_steps=n_steps, noise=noise, trace=True)
    tra = a._last_trace
    b = v.mcmc.MCMC(model, energy, random_seed=4)
    b.fuse_notebook = False
    v.set_seed(5)
    xb, eb = x0, None
    acc_b = np.empty((n_steps, B), bool)
    for s in range(n_steps):
        xb, eb = b.single_step(xb, energies=eb)
        acc_b[s] = b._last_acc.numpy().astype(bool)
    diff = (tra['acc'].astype(bool) != acc_b).any(axis=0)
    assert diff.sum() <= 2, diff.sum()
    assert_close(xa[~diff], xb[~diff], rtol=1e-5, atol=1e-4, what='final configs')
    assert_close(ea[~diff], eb[~diff], rtol=1e-5, atol=3e-4, what='final energies')
    assert 0.01 < a.acceptance_rate < 0.99
    # tfp's full procedure (D + 1 sampling passes + log_prob pass) gives the same bits as the one-pass path
    monkeypatch.setenv('VMS_NB_GENERIC', '1')
    g = v.mcmc.MCMC(model, energy, random_seed=4)
    assert g._nb_plan()['model'].made_first_dof == -1
    xg, eg = g.run_nb(x0, n_steps=n_steps, noise=noise, trace=True)
    assert np.array_equal(xg, xa) and np.array_equal(eg, ea)
    for key in ('acc', 'fwd', 'rev', 'e_new'):
        assert np.array_equal(g._last_trace[key], tra[key]), key


def test_mc_notebook_workflow_train_then_sample(vms):
    """The workflow of examples/MC_Moves_with_VAEs.ipynb end to end on the device: build the notebook's VAE (cells 11-20),
    train it on samples of the Gaussian mixture (cell 22-25; the generic tape path: MAF prior, autoregressive decoder,
    KL estimate), then run VAE-proposal MC from data samples (cells 39-43; the fused notebook-family kernel reads the
    TRAINED weights).  Training lowers the loss and raises the acceptance rate."""
    v = vms
    import bench
    model = bench.build_c4b_model(v, seed=5)
    # (undo the bench's hand-aimed decoder: start from a plain random initialisation of the output layers)
    rng = np.random.default_rng(0)
    x = bench.gmm_start(4096, seed=9)
    energy = v.mcmc.GaussianMixtureEnergy()
    before = v.mcmc.MCMC(model, energy, random_seed=1)
    assert before._nb_plan() is not None
    before.run(x[:2048], n_steps=20)
    model.compile(optimizer=v.models.Adam(learning_rate=2e-3), loss=v.losses.LogProbLoss())
    l0 = np.mean([model.evaluate(x, batch_size=4096) for _ in range(3)])
    hist = model.fit(x, x, epochs=12, batch_size=256)
    l1 = np.mean([model.evaluate(x, batch_size=4096) for _ in range(3)])
    assert np.isfinite(hist['loss']).all() and l1 < l0 - 0.3, (l0, l1)
    after = v.mcmc.MCMC(model, energy, random_seed=1)
    xs, es = after.run(x[:2048], n_steps=20)
    assert np.isfinite(xs).all() and np.isfinite(es).all()
    assert after.acceptance_rate > before.acceptance_rate + 0.02, (before.acceptance_rate, after.acceptance_rate)
    # the chains stay in the mixture: their energies (log-densities) are those of data samples, not of outliers
    assert np.median(es) > np.median(energy(x[:2048])) - 1.0
