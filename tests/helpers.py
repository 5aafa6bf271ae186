"""Shared helpers for the parity tests: conversions between the oracle's per-layer parameter dicts and the C ABI's flat
layout (include/vms_b200.h: heads of a flow block are column-concatenated [w | h | s])."""
import numpy as np

from oracle import vae as ovae


def flat_from_oracle(P):
    parts = []
    for net in ('enc', 'dec'):
        for W, b in P[net]:
            parts += [W.reshape(-1), b.reshape(-1)]
    for blk in P.get('flow', []):
        parts += [blk['d1'][0].reshape(-1), blk['d1'][1].reshape(-1)]
        parts += [np.concatenate([blk[k][0] for k in ('w', 'h', 's')], axis=1).reshape(-1)]
        parts += [np.concatenate([blk[k][1] for k in ('w', 'h', 's')]).reshape(-1)]
    return np.concatenate(parts).astype(np.float32)


def flat_grad_from_oracle(P, G):
    parts = []
    for net in ('enc', 'dec'):
        for gW, gb in G[net]:
            parts += [gW.reshape(-1), gb.reshape(-1)]
    for blk in G.get('flow', []):
        parts += [blk['d1'][0].reshape(-1), blk['d1'][1].reshape(-1)]
        parts += [np.concatenate([blk[k][0] for k in ('w', 'h', 's')], axis=1).reshape(-1)]
        parts += [np.concatenate([blk[k][1] for k in ('w', 'h', 's')]).reshape(-1)]
    return np.concatenate(parts).astype(np.float32)


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=''):
    """north_star tolerance: 1e-5 relative in fp32 (atol covers values near zero)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, '%s shape %s != %s' % (what, a.shape, b.shape)
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    bad = err > tol
    assert not bad.any(), '%s: %d / %d elements differ, max err %.3e (tol %.3e) at %s: %r vs %r' % (
        what, bad.sum(), bad.size, err.max(), tol.reshape(-1)[np.argmax(err)], np.unravel_index(np.argmax(err), err.shape),
        a.reshape(-1)[np.argmax(err)], b.reshape(-1)[np.argmax(err)])


def vae_from_oracle(v, P, max_batch=4096, weight=1.0):
    """Builds the product VAE (host API) with the oracle's weights."""
    from vaemolsim_b200 import dists, flows, losses, models
    import vaemolsim_b200._protocols as PR
    dx, dz = P['dx'], P['dz']
    enc = models.MappingToDistribution(PR.IndependentNormal(dz), mapping=None, name='encoder')
    dec = models.MappingToDistribution(PR.IndependentNormal(dx), mapping=None, name='decoder')
    enc.mapping.hidden_dim = [P['hidden']]
    dec.mapping.hidden_dim = [P['hidden']]
    if P['prior'] == 'normal':
        prior = PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], dz))
    else:
        nb = len(P['flow'])
        fh = P['flow'][0]['d1'][0].shape[1]
        flow = flows.RQSSplineRealNVP(num_blocks=nb, rqs_params=dict(bin_range=list(P['bin_range']),
                                                                      num_bins=P['num_bins'], hidden_dim=fh))
        prior = dists.FlowedDistribution(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], dz)))
    model = models.VAE(enc, dec, prior, regularizer=losses.KLDivergenceEstimate(weight=weight))
    x0 = np.zeros((2, dx), np.float32)
    model(x0)  # build
    dense = lambda m: [l for l in m.layer_list if hasattr(l, 'kernel')]
    for lay, (W, b) in zip(dense(enc.mapping), P['enc']):
        lay.assign(W, b)
    for lay, (W, b) in zip(dense(dec.mapping), P['dec']):
        lay.assign(W, b)
    if P['prior'] != 'normal':
        blocks = [b.bijector_fn for b in prior.flow.chain.bijectors[::-1]]
        for sb, blk in zip(blocks, P['flow']):
            sb.d1.assign(*blk['d1'])
            sb.heads.assign(np.concatenate([blk[k][0] for k in ('w', 'h', 's')], axis=1),
                            np.concatenate([blk[k][1] for k in ('w', 'h', 's')]))
    return model


def vae_b_from_oracle(v, P, made_activation=None):
    """The MC notebook's model family (examples/MC_Moves_with_VAEs.ipynb cells 11-20) through the product's host API, with
    the weights of `oracle.mcmc.init_vae_b`."""
    from vaemolsim_b200 import dists, flows, models
    import vaemolsim_b200._protocols as PR
    enc = models.MappingToDistribution(PR.IndependentNormal(1), name='encoder')
    dec_dist = dists.AutoregressiveBlockwise(2, [dists.Normal] * 2, conditional=True, conditional_event_shape=(1, ),
                                             auto_net_params={'hidden_units': [L['W'].shape[1] for L in P['made'][:-1]],
                                                              'activation': made_activation})
    dec = models.MappingToDistribution(dec_dist, name='decoder')
    enc.mapping.hidden_dim = [P['hidden']]
    dec.mapping.hidden_dim = [P['hidden']]
    nb = len(P['maf'])
    H = P['maf'][0]['w'][0]['W'].shape[1]
    flow = flows.RQSSplineMAF(num_blocks=nb, rqs_params=dict(bin_range=list(P['bin_range']), num_bins=P['num_bins'],
                                                             hidden_dim=H))
    flow(np.zeros((2, 1), np.float32))
    prior = dists.FlowedDistribution(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 1)), name='prior')
    model = models.VAE(enc, dec, prior)
    model(np.zeros((2, 2), np.float32))  # build
    dense = lambda m: [l for l in m.layer_list if hasattr(l, 'kernel')]
    for lay, (W, b) in zip(dense(enc.mapping), P['enc']):
        lay.assign(W, b)
    for lay, (W, b) in zip(dense(dec.mapping), P['dec']):
        lay.assign(W, b)
    arrs = []
    for L in P['made']:
        arrs += [L['W'], L['b'], L['Wc']]
    dec_dist.auto_net.set_weights(arrs)
    for bij, blk in zip(flow.chain.bijectors[::-1], P['maf']):
        msb = bij.bijector_fn
        for key, net in (('w', msb.bin_widths), ('h', msb.bin_heights), ('s', msb.knot_slopes)):
            net.set_weights([a for L in blk[key] for a in (L['W'], L['b'])])
    return model
