"""GPU parity tests AT the BASELINE.json configurations (C2 / C3 / C4a / C5 shapes) and for the entry points
round 1 left without a device test: every comparison here is CUDA path vs the CPU oracle (float64 where it is cheap),
never CUDA vs CUDA."""
import os

import numpy as np
import pytest

from helpers import assert_close, flat_grad_from_oracle, vae_from_oracle
from oracle import dists as odists
from oracle import flows as oflows
from oracle import mappings as omap
from oracle import mcmc as omc
from oracle import nets as onets
from oracle import vae as ovae

pytestmark = pytest.mark.gpu


def _c2_params(seed=2003, widen=True):
    """The bench's C2 model: enc 6-200-4, 4 RealNVP-RQS blocks K = 32 H = 100 over N(0, I_2), dec 2-200-12 (44,396 params)."""
    P = ovae.init_vae(seed, dx=6, dz=2, hidden=200, prior='realnvp', num_blocks=4, num_bins=32, flow_hidden=100)
    assert ovae.param_count(P) == 44396
    if widen:  # spline heads away from the near-identity initialisation, so bins / slopes vary along the batch
        rng = np.random.default_rng(seed + 1)
        for blk in P['flow']:
            for k in ('w', 'h', 's'):
                blk[k] = ((blk[k][0] * 4).astype(np.float32), rng.normal(0, 0.5, blk[k][1].shape).astype(np.float32))
    return P


def _check_against_float64_oracle(f, v, P, x, eps, weight, what):
    """north_star: log-probs and gradients within 1e-5 relative in fp32.  Loss scalars (batch means of the log-probs) are
    held to 1e-5.  The flat gradient is held to 1e-5 (norm) OR to twice the error the float32 NumPy oracle -- the reference's
    own arithmetic -- makes on the same inputs against float64, whichever is larger: on the widened (ill-conditioned) splines
    of these tests float32 evaluation itself is 0.7-2.4e-5 away from the float64 gradient (scripts/diag_c2_precision.py)."""
    P64 = ovae.cast_params(P, np.float64)
    out, G = ovae.elbo_backward(P64, x.astype(np.float64), eps.astype(np.float64), weight=weight)
    want = flat_grad_from_oracle(P64, G).astype(np.float64)
    _, G32 = ovae.elbo_backward(P, x, eps, weight=weight)
    rel32 = np.linalg.norm(flat_grad_from_oracle(P, G32).astype(np.float64) - want) / np.linalg.norm(want)
    scal = f.forward_backward(v.as_tensor(x), v.as_tensor(eps)).numpy()
    assert_close(scal[:3], [out['loss'], out['nll'], out['kl']], rtol=1e-5, atol=1e-5, what='%s: loss / nll / kl' % what)
    got = f.grad.numpy().astype(np.float64)
    rel = np.linalg.norm(got - want) / np.linalg.norm(want)
    print('%s: flat gradient norm-relative error %.2e (float32 oracle: %.2e)' % (what, rel, rel32))
    assert rel < max(1e-5, 2.0 * rel32), '%s: flat gradient norm-relative error %.2e (float32 oracle %.2e)' % (what, rel, rel32)
    assert_close(got, want, rtol=1e-4, atol=max(5e-6, 8.0 * rel32) * max(np.abs(want).max(), 1e-3), what='%s: flat gradient' % what)
    return out


@pytest.mark.parametrize('mode,name', [(4, 'fused'), (1, 'ffma'), (2, 'tensor-core'), (3, 'tensor-core-fused'),
                                       (0, 'tensor-core-fused')])
def test_c2_exact_shape_every_plan_vs_float64_oracle(vms, mode, name):
    """BASELINE configs[1] exactly -- H = 200, FH = 100, K = 32, B = 4096, 44,396 parameters -- through every plan of
    `vms_elbo_forward_backward`, against the float64 oracle on ALL rows (0.4 s of CPU)."""
    v = vms
    P = _c2_params()
    B = 4096
    rng = np.random.default_rng(1001)
    x = rng.standard_normal((B, 6), dtype=np.float32)
    eps = rng.standard_normal((B, 2), dtype=np.float32)
    f = vae_from_oracle(v, P, max_batch=B, weight=1.0).fused(B)
    f.set_tc_auto_batch(1 << 40)
    f.set_mode(mode)
    assert f.path(B) == name
    out = _check_against_float64_oracle(f, v, P, x, eps, 1.0, 'C2 B=4096 plan %s' % name)
    assert not f.tc_status()
    if mode in (4, 1, 2):  # per-row outputs of the forward entry point as well
        fw = f.forward(v.as_tensor(x), v.as_tensor(eps))
        for k in ('z', 'logq', 'logpz', 'logpx'):
            assert_close(fw[k].numpy(), out[k], rtol=1e-5, atol=3e-5, what='C2 %s %s' % (name, k))


@pytest.mark.parametrize('B,mode,name', [(10007, 2, 'tensor-core'), (10007, 0, 'tensor-core-fused'),
                                         (14239, 0, 'tensor-core')])
def test_c5_shard_large_batch_plans_vs_float64_oracle(vms, B, mode, name):
    """The plans that serve batches beyond one wave of tiles, at the C2 widths with a ragged last tile, float64 oracle on
    all rows: the per-block tensor-core plan (forced at B = 10,007, auto-selected above three waves of 32-row tiles) and
    the whole-step tensor-core kernel running three tiles per SM (auto at B = 10,007 = 313 tiles on 148 SMs)."""
    v = vms
    P = _c2_params(seed=2011)
    rng = np.random.default_rng(B)
    x = rng.standard_normal((B, 6), dtype=np.float32)
    eps = rng.standard_normal((B, 2), dtype=np.float32)
    f = vae_from_oracle(v, P, max_batch=B, weight=0.7).fused(B)
    f.set_mode(mode)
    assert f.path(B) == name
    _check_against_float64_oracle(f, v, P, x, eps, 0.7, 'C5-shard B=%d (%s)' % (B, name))
    assert not f.tc_status()


@pytest.mark.parametrize('k', [50, 10])
def test_dist_select_c3_shape_bit_exact(vms, k):
    """BASELINE configs[2] / SURVEY C3: one frame of N = 10,000 particles in a periodic box L = 46.416 (density 0.1)
    replicated per reference row, cutoff 3.0, max_included 50 (and 10), one-hot particle_info P = 2, int32 indices on --
    values AND indices bit-exact against the oracle, through the layer API (`DistanceSelection.__call__`)."""
    v = vms
    rng = np.random.default_rng(3001)
    N, L, B = 10000, np.float32(46.416), 64
    frame = rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32)
    coords = np.ascontiguousarray(np.broadcast_to(frame, (B, N, 3)))
    ref = np.random.default_rng(3002).uniform(-L / 2, L / 2, (B, 1, 3)).astype(np.float32)
    kinds = rng.integers(0, 2, N)
    info = np.ascontiguousarray(np.broadcast_to(np.eye(2, dtype=np.float32)[kinds], (B, N, 2)))
    box = np.array([L, L, L], np.float32)
    layer = v.mappings.DistanceSelection(3.0, max_included=k, box_lengths=box)
    sel, sinfo, idx = layer(coords, ref, particle_info=info, return_indices=True)
    want = omap.distance_selection(coords, ref.reshape(B, 3), 3.0, k, box_lengths=box, particle_info=info,
                                   return_indices=True)
    assert np.array_equal(idx.numpy(), want[2]), 'C3 neighbour indices differ'
    assert np.array_equal(sel.numpy(), want[0]) and np.array_equal(sinfo.numpy(), want[1])
    # ~11 neighbours within the cutoff at this density: both the "fewer than k inside" and (k = 10) "more than k" branches
    inside = (np.abs(want[0]).sum(-1) > 0).sum(1)
    assert inside.min() >= 1 and (inside.max() == k if k == 10 else inside.max() < k)
    # the values-only call (no indices: the path the bench times) returns the same selection
    assert np.array_equal(layer(coords, ref).numpy(), want[0])


def test_mc_c4a_4096_of_65536_chains_100_steps_vs_oracle(vms):
    """BASELINE configs[3] / C4a: chains [8192, 12288) of the 65,536-chain job, 100 MC steps in ONE launch with the uniform
    stream drawn on the device, against the oracle restatement of mcmc.py:68-130 (itself pinned to the reference's mcmc.py)
    driven with the same sampling noise and the same PCG64 columns.  Chains are compared decision by decision; a chain whose
    float32 log-probabilities put a decision within rounding of its threshold may flip and then diverges -- the flip count
    is reported and bounded."""
    v = vms
    n_global, lo, B, n_steps = 65536, 8192, 4096, 100
    P = ovae.init_vae(2003, dx=6, dz=2, hidden=200, prior='normal')
    model = vae_from_oracle(v, P)
    x0 = np.random.default_rng(4001).standard_normal((n_global, 6), dtype=np.float32)[lo:lo + B]
    rng = np.random.default_rng(777)
    noise = np.empty((n_steps, B, 10), np.float32)
    for s in range(n_steps):  # the order the reference's step draws in (mcmc.py:100-102): encoder, prior, decoder
        noise[s, :, :2] = rng.standard_normal((B, 2), dtype=np.float32)
        noise[s, :, 2:4] = rng.standard_normal((B, 2), dtype=np.float32)
        noise[s, :, 4:] = rng.standard_normal((B, 6), dtype=np.float32)
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002, stream_layout=(lo, n_global))
    x_dev, e_dev = mc.run_fused(x0, n_steps=n_steps, noise=noise, trace=True)
    tr = mc._last_trace
    assert mc.host_stream_reruns == 0
    # oracle: same noise (OracleVAE draws from default_rng(777) in the same order), same columns of the uniform stream
    ovm = omc.OracleVAE(P, noise_seed=777)
    u = np.random.default_rng(4002).random(size=(n_steps, n_global))[:, lo:lo + B]
    np.testing.assert_array_max_ulp(tr['log_u'], np.log(u), maxulp=2)

    class _Cols(object):  # hands the oracle step the shard's columns of the global stream
        def __init__(self):
            self.s = 0

        def random(self, size):
            self.s += 1
            return u[self.s - 1]

    cols = _Cols()
    xo, eo = x0.copy(), None
    acc_o = np.empty((n_steps, B), bool)
    for s in range(n_steps):
        xo, eo, acc_o[s] = omc.single_step(ovm, omc.quadratic_energy, cols, xo, eo)
    acc_d = tr['acc'].astype(bool)
    differs = (acc_d != acc_o).any(axis=0)
    flips = int(differs.sum())
    print('C4a parity: %d of %d chains flipped a decision in %d steps (accepted: device %d, oracle %d)' %
          (flips, B, n_steps, acc_d.sum(), acc_o.sum()))
    assert flips <= 4, flips
    same = ~differs
    assert_close(x_dev[same], xo[same], rtol=1e-5, atol=2e-5, what='C4a final configurations')
    assert_close(e_dev[same], eo[same], rtol=1e-5, atol=2e-5, what='C4a final energies')
    assert mc._num_trials == B * n_steps and mc._num_acc == float(acc_d.sum())


# ------------------------------------------------------------------------------------------------ untested entry points
@pytest.mark.parametrize('B,mask', [(1, [True, False, True]), (257, [False, True, True, False, False, True]),
                                    (4096, [True] * 6), (33, [False] * 4)])
def test_periodic_featurise_matches_oracle(vms, B, mask):
    """`vms_periodic_featurise` (mappings.py:144-149): [x[~periodic] | cos(x[periodic]) | sin(x[periodic])], directly and
    through `FCDeepNN(periodic_dofs=...)` against the oracle network."""
    v = vms
    c = v._abi.ctx()
    rng = np.random.default_rng(B)
    D = len(mask)
    m = np.asarray(mask, bool)
    x = rng.uniform(-4 * np.pi, 4 * np.pi, (B, D)).astype(np.float32)
    if m.any():
        out = v.Tensor((B, D + int(m.sum())))
        xd, md = v.as_tensor(x), v.Tensor.from_numpy(m.astype(np.uint8))  # keep the device buffers alive across the launch
        c.lib.vms_periodic_featurise(xd.ptr, B, D, md.ptr, out.ptr, c.stream)
        want = onets.periodic_featurise(x.astype(np.float64), m)
        assert out.numpy().shape == want.shape
        assert_close(out.numpy(), want, rtol=1e-5, atol=2e-6, what='periodic featurise')  # |x| <= 4 pi: cos / sin abs 1e-6
    net = v.mappings.FCDeepNN((3, 2), hidden_dim=[20, 12], periodic_dofs=list(mask), activation='tanh')
    y = net(x)
    assert y.shape == (B, 3, 2)
    dense = [l for l in net.layer_list if hasattr(l, 'kernel')]
    layers = [tuple(a.astype(np.float64) for a in l.get_weights()) for l in dense]
    assert dense[0].kernel.shape[0] == D + int(m.sum())
    want = onets.fcdeepnn_forward(x.astype(np.float64), layers, (3, 2), periodic_mask=m if m.any() else None,
                                  activation='tanh')
    assert_close(y.numpy(), want, rtol=1e-5, atol=1e-5, what='FCDeepNN with periodic dofs')


def test_fcdeepnn_periodic_shape_error(vms):
    net = vms.mappings.FCDeepNN(4, periodic_dofs=[True, False])
    with pytest.raises(ValueError, match='periodic_dofs'):
        net(np.zeros((3, 5), np.float32))  # mappings.py:99-101


def _flow_blocks_from_layer(flow, K):
    blocks = []
    for bij in flow.chain.bijectors[::-1]:
        sb = bij.bijector_fn
        W, b = sb.heads.get_weights()
        d1W, d1b = sb.d1.get_weights()
        nw = sb.data_dim * K
        blocks.append({'d1': (d1W.astype(np.float64), d1b.astype(np.float64)),
                       'w': (W[:, :nw].astype(np.float64), b[:nw].astype(np.float64)),
                       'h': (W[:, nw:2 * nw].astype(np.float64), b[nw:2 * nw].astype(np.float64)),
                       's': (W[:, 2 * nw:].astype(np.float64), b[2 * nw:].astype(np.float64))})
    return blocks


def _widen(flow, rng):
    for bij in flow.chain.bijectors:
        sb = bij.bijector_fn
        W, b = sb.heads.get_weights()
        sb.heads.assign(W + rng.normal(0, 0.3, W.shape).astype(np.float32), b + rng.normal(0, 0.3, b.shape).astype(np.float32))


def test_static_flowed_distribution_matches_oracle(vms):
    """`StaticFlowedDistribution` (dists.py:515-530): a static distribution OBJECT pushed through a flow; inputs are
    ignored; log_prob / sample against the oracle chain."""
    v = vms
    import vaemolsim_b200._protocols as PR
    rng = np.random.default_rng(12)
    D, K, B = 2, 8, 300
    flow = v.flows.RQSSplineRealNVP(num_blocks=3, rqs_params=dict(num_bins=K, hidden_dim=16, bin_range=[-6.0, 6.0]))
    flow(np.zeros((2, D), np.float32))  # tensor input: builds the chain and its spline networks
    _widen(flow, rng)
    layer = v.dists.StaticFlowedDistribution(flow, PR.StandardNormal(None, D))
    d1 = layer(None)
    d2 = layer(np.ones((7, 5), np.float32), training=True)  # inputs ignored, training forwarded (no batch norm: no effect)
    y = rng.normal(0, 2, (B, D)).astype(np.float32)
    blocks = _flow_blocks_from_layer(flow, K)
    xo, ildj = oflows.realnvp_inverse(y.astype(np.float64), blocks, K, (-6.0, 6.0))
    want = odists.normal_log_prob(xo, 0.0, 1.0).sum(-1) + ildj
    assert_close(d1.log_prob(y).numpy(), want, rtol=1e-5, atol=2e-5, what='StaticFlowedDistribution.log_prob')
    assert np.array_equal(d1.log_prob(y).numpy(), d2.log_prob(y).numpy())
    s = d1.sample(64).numpy()
    assert s.shape == (64, D) and np.isfinite(s).all()
    xs, _ = oflows.realnvp_inverse(s.astype(np.float64), blocks, K, (-6.0, 6.0))
    assert np.abs(xs).max() < 6.0  # samples invert to plausible N(0, 1) draws


def test_potential_energy_log_prob_loss_matches_oracle(vms):
    """`PotentialEnergyLogProbLoss` (losses.py:94-113): potential(samples) - decoder.log_prob(samples) per sample, batch
    mean through `__call__`, samples drawn from the decoder when None."""
    v = vms
    import vaemolsim_b200._protocols as PR
    rng = np.random.default_rng(5)
    B, D = 500, 3
    params = rng.normal(size=(B, 2 * D)).astype(np.float32)
    dec = PR.IndependentNormal(D)(v.as_tensor(params))
    pot = lambda s: np.sum(np.asarray(s.numpy() if hasattr(s, 'numpy') else s, np.float64)**2, axis=-1).astype(np.float32)
    x = rng.normal(size=(B, D)).astype(np.float32)
    want = pot(v.as_tensor(x)).astype(np.float64) - odists.independent_normal_log_prob(x.astype(np.float64),
                                                                                      params.astype(np.float64))
    loss = v.losses.PotentialEnergyLogProbLoss(pot, reduction='none')
    assert_close(loss(v.as_tensor(x), dec).numpy(), want, rtol=1e-5, atol=1e-5, what='PotentialEnergyLogProbLoss per sample')
    mean = v.losses.PotentialEnergyLogProbLoss(pot)(v.as_tensor(x), dec)
    assert_close(float(mean.numpy()), want.mean(), rtol=1e-5, atol=1e-5, what='PotentialEnergyLogProbLoss mean')
    v.set_seed(3)
    drawn = v.losses.PotentialEnergyLogProbLoss(pot, reduction='none')(None, dec).numpy()
    assert drawn.shape == (B, ) and np.isfinite(drawn).all()
    assert v.losses.PotentialEnergyLogProbLoss(pot).get_config()['potential'] is pot


def test_flow_model_static_and_mapped_latent_match_oracle(vms):
    """`FlowModel` (models.py:85-139).  (1) Using_Normalizing_Flows cell 10: a 1-D RQS RealNVP over a static N(0, 1) --
    no mapping, log_prob of data = flow NLL.  (2) a non-static latent: FCDeepNN -> IndependentNormal -> flow."""
    v = vms
    import vaemolsim_b200._protocols as PR
    rng = np.random.default_rng(8)
    K = 32
    flow = v.flows.RQSSplineRealNVP(num_blocks=4, rqs_params=dict(bin_range=[-10.0, 10.0], num_bins=K, hidden_dim=100))
    latent = PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 1))
    fm = v.models.FlowModel(flow, latent)
    assert fm.mapping is None
    x = rng.normal(0, 3, (1000, 1)).astype(np.float32)
    fm(x).log_prob(x)  # the spline networks are built at their first evaluation
    _widen(flow, rng)
    dist = fm(x)
    blocks = _flow_blocks_from_layer(flow, K)
    xo, ildj = oflows.realnvp_inverse(x.astype(np.float64), blocks, K, (-10.0, 10.0))
    want = odists.normal_log_prob(xo, 0.0, 1.0).sum(-1) + ildj
    # four chained K = 32 splines on N(0, 3) data: a float32 knot position is uncertain by ulp(10) ~ 1e-6, which narrow bins
    # amplify in the log-det; bound the error by the float32 oracle's own error on the same inputs
    f32 = [{k: (a.astype(np.float32), b.astype(np.float32)) for k, (a, b) in blk.items()} for blk in blocks]
    xo32, ildj32 = oflows.realnvp_inverse(x, f32, K, (-10.0, 10.0))
    err32 = np.abs((odists.normal_log_prob(xo32, 0.0, 1.0).sum(-1) + ildj32).astype(np.float64) - want).max()
    assert_close(dist.log_prob(x).numpy(), want, rtol=1e-5, atol=max(2e-5, 2.0 * err32),
                 what='FlowModel (static latent) log_prob')
    nll = v.losses.LogProbLoss()(v.as_tensor(x), dist)
    assert_close(float(nll.numpy()), -want.mean(), rtol=1e-5, atol=1e-5, what='FlowModel NLL')
    assert fm.predict(x[:40], batch_size=16).shape == (40, 1)
    # (2) a latent layer that is NOT a tfp.layers.DistributionLambda gets an inferred FCDeepNN mapping (models.py:74-78)
    D = 2
    flow2 = v.flows.RQSSplineRealNVP(num_blocks=2, rqs_params=dict(num_bins=8, hidden_dim=16))
    fm2 = v.models.FlowModel(flow2, v.dists.IndependentBlockwise(D, v.dists.Normal))
    assert isinstance(fm2.mapping, v.mappings.FCDeepNN) and fm2.mapping.target_shape == (2 * D, )
    cond = rng.normal(size=(200, 5)).astype(np.float32)
    y = rng.normal(0, 1.5, (200, D)).astype(np.float32)
    fm2(cond).log_prob(y)
    _widen(flow2, rng)
    d2 = fm2(cond)
    dense = [l for l in fm2.mapping.layer_list if hasattr(l, 'kernel')]
    layers = [tuple(a.astype(np.float64) for a in l.get_weights()) for l in dense]
    params = onets.fcdeepnn_forward(cond.astype(np.float64), layers, (2 * D, ))
    blocks2 = _flow_blocks_from_layer(flow2, 8)
    xo2, ildj2 = oflows.realnvp_inverse(y.astype(np.float64), blocks2, 8, (-10.0, 10.0))
    want2 = odists.independent_blockwise_log_prob(xo2, params, ['normal'] * D) + ildj2
    assert_close(d2.log_prob(y).numpy(), want2, rtol=1e-5, atol=3e-5, what='FlowModel (mapped latent) log_prob')
    # a tfp.layers-style latent (IndependentNormal IS a DistributionLambda) gets no mapping, exactly as in the reference
    assert v.models.FlowModel(flow2, PR.IndependentNormal(D)).mapping is None


def test_mc_c4b_1024_of_65536_chains_100_steps_vs_oracle(vms):
    """BASELINE configs[3] / C4b (the MC notebook's model at its full widths: hidden 200, 4 MAF blocks, K = 20, H = 40, MADE
    [10, 100, 10]): chains [20480, 21504) of the 65,536-chain job, 100 MC steps in ONE launch of the fused kernel with the
    uniform stream drawn on the device, against the oracle restatement of mcmc.py:68-130 on `OracleVAEb` with the same
    sampling noise and the same PCG64 columns.  Chains are compared decision by decision; the flip count is bounded."""
    v = vms
    from helpers import vae_b_from_oracle
    n_global, lo, B, n_steps = 65536, 20480, 1024, 100
    P = omc.init_vae_b(7, hidden=200)  # (a seed whose untrained proposal is accepted ~20 % of the time)
    model = vae_b_from_oracle(v, P)
    rng0 = np.random.default_rng(5001)
    k = rng0.choice(3, size=n_global, p=[0.7, 0.2, 0.1])
    x0 = (omc.GMM_LOCS[k] + omc.GMM_SCALES[k] * rng0.standard_normal((n_global, 2))).astype(np.float32)[lo:lo + B]
    rng = np.random.default_rng(888)
    noise = np.empty((n_steps, B, 4), np.float32)
    for s in range(n_steps):  # the order the reference's step draws in (mcmc.py:100-102): encoder, prior, decoder
        noise[s, :, 0:1] = rng.standard_normal((B, 1), dtype=np.float32)
        noise[s, :, 1:2] = rng.standard_normal((B, 1), dtype=np.float32)
        noise[s, :, 2:4] = rng.standard_normal((B, 2), dtype=np.float32)
    mc = v.mcmc.MCMC(model, v.mcmc.GaussianMixtureEnergy(), random_seed=5002, stream_layout=(lo, n_global))
    assert mc._nb_plan() is not None
    x_dev, e_dev = mc.run_nb(x0, n_steps=n_steps, noise=noise, trace=True)
    tr = mc._last_trace
    assert mc.host_stream_reruns == 0
    u = np.random.default_rng(5002).random(size=(n_steps, n_global))[:, lo:lo + B]
    np.testing.assert_array_max_ulp(tr['log_u'], np.log(u), maxulp=2)

    class _Cols(object):  # hands the oracle step the shard's columns of the global stream
        s = 0

        def random(self, size):
            self.s += 1
            return u[self.s - 1]

    ovm, cols = omc.OracleVAEb(P, noise_seed=888), _Cols()
    xo, eo = x0.copy(), None
    acc_o = np.empty((n_steps, B), bool)
    for s in range(n_steps):
        xo, eo, acc_o[s] = omc.single_step(ovm, omc.gmm_energy, cols, xo, eo)
    acc_d = tr['acc'].astype(bool)
    differs = (acc_d != acc_o).any(axis=0)
    flips = int(differs.sum())
    print('C4b parity: %d of %d chains flipped a decision in %d steps (accepted: device %d, oracle %d)' %
          (flips, B, n_steps, acc_d.sum(), acc_o.sum()))
    assert flips <= 4, flips
    same = ~differs
    assert_close(x_dev[same], xo[same], rtol=1e-5, atol=1e-4, what='C4b final configurations')
    # (the mixture's narrow component, sigma = 0.05, makes the energy steep: |dE/dx| = |x - mu| / sigma^2 reaches ~40, so the
    #  1e-5 agreement of the configurations after 100 float32 steps is a few 1e-4 in the energy)
    assert_close(e_dev[same], eo[same], rtol=1e-5, atol=2e-3, what='C4b final energies')
    assert mc._num_trials == B * n_steps and mc._num_acc == float(acc_d.sum()) and acc_d.sum() > 1000


def test_dist_select_shared_frame_equals_tiled_call(vms):
    """`DistanceSelection.select_from_frame` (vms_dist_select_frame): B sites around ONE frame give exactly the tiled
    reference-shaped call -- coordinates, particle info and top_k indices -- with a stored box, a per-site box and no box;
    and the oracle on the first rows."""
    v = vms
    rng = np.random.default_rng(8)
    N, B, k, L = 3001, 97, 24, np.float32(21.5)
    frame = rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32)
    info = np.eye(3, dtype=np.float32)[rng.integers(0, 3, N)]
    ref = rng.uniform(-L / 2, L / 2, (B, 3)).astype(np.float32)
    tiled, tinfo = np.ascontiguousarray(np.broadcast_to(frame, (B, N, 3))), np.ascontiguousarray(np.broadcast_to(info, (B, N, 3)))
    for box, per_site in ((np.array([L, L, L], np.float32), None), (None, rng.uniform(18, 25, (B, 3)).astype(np.float32)),
                          (None, None)):
        layer = v.mappings.DistanceSelection(5.0, max_included=k, box_lengths=box)
        a = [t.numpy() for t in layer.select_from_frame(frame, ref, box_lengths=per_site, particle_info=info, return_indices=True)]
        b = [t.numpy() for t in layer(tiled, ref, box_lengths=per_site, particle_info=tinfo, return_indices=True)]
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        want = omap.distance_selection(tiled[:8], ref[:8], 5.0, k, box_lengths=box if per_site is None else per_site[:8],
                                       particle_info=tinfo[:8], return_indices=True)
        assert np.array_equal(a[0][:8], want[0]) and np.array_equal(a[1][:8], want[1]) and np.array_equal(a[2][:8], want[2])
    only = v.mappings.DistanceSelection(5.0, max_included=k).select_from_frame(frame, ref)
    assert only.shape == (B, k, 3)

