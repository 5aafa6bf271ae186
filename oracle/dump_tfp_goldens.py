"""Dump golden vectors from the REAL reference (TensorFlow <= 2.15, TensorFlow-Probability <= 0.23, vaemolsim) -- test
infrastructure only.

    python oracle/dump_tfp_goldens.py [--reference /path/to/vae-mol-sim] [--out tests/golden]

WHY.  The arithmetic of the hot path lives in TFP; this image has Python 3.12 and no network, so TF / TFP cannot
be installed here and everything except `mcmc.py` is "parity unpinned": the NumPy oracle restates the published TFP v0.23
algorithms and is validated by identities, not by TFP's own output.  This script closes that gap wherever TF exists: run it
ONCE on any box with the reference's environment (devtools/conda-envs/test_env.yaml: python <= 3.11, tensorflow <= 2.15,
tensorflow-probability <= 0.23), commit the `tfp_*.npz` files it writes, and `tests/test_oracle.py::test_oracle_matches_tfp_
goldens` (skipped while the files are absent) pins the oracle -- and through it every CUDA parity test -- to real TFP output.

WHAT.  Fixed inputs and weights (NumPy default_rng seeds, never Keras initialisers, which cannot be reproduced outside TF)
are pushed through the reference's own layers; every output the oracle restates is stored next to its inputs:
  tfp_rqs.npz        flows.SplineBijector activations (flows.py:86-101) + tfp.bijectors.RationalQuadraticSpline forward /
                     inverse / forward_log_det_jacobian on raw logits, K in {8, 20, 32}, inside and outside the range,
                     and d(y, ldj)/d(x, raw logits) by tf.GradientTape
  tfp_realnvp.npz    flows.RQSSplineRealNVP (flows.py:281-355) D in {1, 2, 3}: forward, inverse, both log-dets, log_prob of
                     the flowed N(0, I)
  tfp_maf.npz        flows.RQSSplineMAF (flows.py:597-690) with order_seed 42, unconditional and conditional
  tfp_dists.npz      dists.make_param_transform / IndependentBlockwise / AutoregressiveBlockwise / IndependentVonMises:
                     constrained parameters and log_prob
  tfp_vae.npz        models.VAE of tests/test_models.py:161-185 (C1) and :190-228 (flow prior, C2 widths reduced): z, the
                     three log-probabilities per row, loss terms and the flat gradient by GradientTape
  tfp_distsel.npz    mappings.DistanceSelection on dense / ragged inputs with and without particle_info: values and top_k
                     indices
  tfp_gaa.npz        mappings.AttentionBlock / ParticleEmbedding (mappings.py:480-688) over the REAL geometric_algebra_attention
                     package (un-vendored, unpinned): inputs with zero-padded particles, every Keras variable of the layer by
                     NAME (so the oracle's weight mapping can be checked without knowing Keras' creation order), outputs with
                     and without mask_zero.  This is what pins oracle/gaa.py.
The weight layouts are the oracle's (`oracle/vae.py::param_list`, `oracle/flows.py`), assigned into the Keras layers with
`set_weights`, so the oracle consumes the same arrays unchanged.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def need_tf(reference):
    try:
        import tensorflow as tf  # noqa: F401
        import tensorflow_probability as tfp  # noqa: F401
    except ImportError as e:
        sys.exit('dump_tfp_goldens: TensorFlow / TensorFlow-Probability are required (%s).  Run this on a box with the '
                 "reference's environment (python <= 3.11, tensorflow <= 2.15, tensorflow-probability <= 0.23)." % e)
    sys.path.insert(0, reference)
    try:
        import vaemolsim  # noqa: F401
    except ImportError as e:
        sys.exit('dump_tfp_goldens: cannot import the reference package from %s (%s)' % (reference, e))
    import tensorflow as tf
    import tensorflow_probability as tfp
    import vaemolsim
    return tf, tfp, vaemolsim


def dump_rqs(tf, tfp, vms, out):
    res = {}
    for K in (8, 20, 32):
        rng = np.random.default_rng(100 + K)
        n = 512
        raw_w = rng.normal(0, 1.0, (n, 1, K)).astype(np.float32)
        raw_h = rng.normal(0, 1.0, (n, 1, K)).astype(np.float32)
        raw_s = rng.normal(0, 1.0, (n, 1, K - 1)).astype(np.float32)
        x = rng.uniform(-12, 12, (n, 1)).astype(np.float32)
        sb = vms.flows.SplineBijector(1, bin_range=[-10.0, 10.0], num_bins=K)
        xt, rw, rh, rs = (tf.constant(a) for a in (x, raw_w, raw_h, raw_s))
        with tf.GradientTape(persistent=True) as tape:
            tape.watch([xt, rw, rh, rs])
            # flows.py:86-101: the activations the layer applies to the Dense outputs
            bw = sb._bin_positions(rw)
            bh = sb._bin_positions(rh)
            ks = sb._slopes(rs)
            bij = tfp.bijectors.RationalQuadraticSpline(bin_widths=bw, bin_heights=bh, knot_slopes=ks, range_min=-10.0)
            y = bij.forward(xt)
            fldj = bij.forward_log_det_jacobian(xt, event_ndims=0)
            xi = bij.inverse(xt)
            ildj = bij.inverse_log_det_jacobian(xt, event_ndims=0)
            sy, sl = tf.reduce_sum(y), tf.reduce_sum(fldj)
        tag = 'K%d_' % K
        res.update({tag + 'x': x, tag + 'raw_w': raw_w, tag + 'raw_h': raw_h, tag + 'raw_s': raw_s,
                    tag + 'bin_widths': bw.numpy(), tag + 'bin_heights': bh.numpy(), tag + 'knot_slopes': ks.numpy(),
                    tag + 'forward': y.numpy(), tag + 'fldj': fldj.numpy(), tag + 'inverse': xi.numpy(),
                    tag + 'ildj': ildj.numpy()})
        for name, target in (('dy', sy), ('dldj', sl)):
            g = tape.gradient(target, [xt, rw, rh, rs])
            for gname, gv in zip(('x', 'raw_w', 'raw_h', 'raw_s'), g):
                res[tag + name + '_d' + gname] = gv.numpy()
    np.savez_compressed(os.path.join(out, 'tfp_rqs.npz'), **res)


def _assign_spline_net(sb, blk):
    """oracle/flows.py block dict {'d1': (W, b), 'w': (W, b), 'h': (W, b), 's': (W, b)} -> SplineBijector sub-layers."""
    sb.d1.set_weights([blk['d1'][0], blk['d1'][1]])
    sb.bin_widths.set_weights([blk['w'][0], blk['w'][1]])
    sb.bin_heights.set_weights([blk['h'][0], blk['h'][1]])
    sb.knot_slopes.set_weights([blk['s'][0], blk['s'][1]])


def dump_realnvp(tf, tfp, vms, out):
    from oracle import flows as oflows
    res = {}
    for D in (1, 2, 3):
        rng = np.random.default_rng(200 + D)
        K, H, nb = 8, 16, 4
        blocks = oflows.realnvp_init(rng, D, num_blocks=nb, num_bins=K, hidden_dim=H)
        for blk in blocks:  # move the splines away from the identity
            for k in ('w', 'h', 's'):
                blk[k] = (blk[k][0] + rng.normal(0, 0.3, blk[k][0].shape).astype(np.float32), blk[k][1])
        flow = vms.flows.RQSSplineRealNVP(num_blocks=nb, rqs_params=dict(num_bins=K, hidden_dim=H))
        x = rng.normal(0, 3, (257, D)).astype(np.float32)
        flow(tf.constant(x))  # build
        # chain.bijectors is applied right to left (flows.py:323): block i of the oracle is bijectors[nb - 1 - i]
        for i, blk in enumerate(blocks):
            _assign_spline_net(flow.chain.bijectors[nb - 1 - i].bijector_fn if hasattr(
                flow.chain.bijectors[nb - 1 - i], 'bijector_fn') else flow.chain.bijectors[nb - 1 - i]._bijector_fn, blk)
        y = flow(tf.constant(x)).numpy()
        base = tfp.distributions.Independent(tfp.distributions.Normal(tf.zeros(D), tf.ones(D)), 1)
        td = flow(base)
        tag = 'D%d_' % D
        res.update({tag + 'x': x, tag + 'y': y,
                    tag + 'fldj': flow.chain.forward_log_det_jacobian(tf.constant(x), event_ndims=1).numpy(),
                    tag + 'x_back': flow.chain.inverse(tf.constant(y)).numpy(),
                    tag + 'ildj': flow.chain.inverse_log_det_jacobian(tf.constant(y), event_ndims=1).numpy(),
                    tag + 'log_prob': td.log_prob(tf.constant(y)).numpy()})
        for i, blk in enumerate(blocks):
            for k, (W, b) in blk.items():
                res['%sblk%d_%s_W' % (tag, i, k)] = W
                res['%sblk%d_%s_b' % (tag, i, k)] = b
    np.savez_compressed(os.path.join(out, 'tfp_realnvp.npz'), **res)


def dump_maf(tf, tfp, vms, out):
    res = {}
    rng = np.random.default_rng(300)
    D, K, H = 3, 8, 12
    for cond in (False, True):
        kw = dict(num_bins=K, hidden_dim=H)
        if cond:
            kw.update(conditional=True, conditional_event_shape=(2, ))
        flow = vms.flows.RQSSplineMAF(num_blocks=3, order_seed=42, rqs_params=kw)
        x = rng.normal(0, 2, (129, D)).astype(np.float32)
        c = rng.normal(size=(129, 2)).astype(np.float32)
        args = dict(conditional_input=tf.constant(c)) if cond else {}
        y = flow(tf.constant(x), **args)
        tag = 'cond_' if cond else 'plain_'
        res.update({tag + 'x': x, tag + 'c': c, tag + 'y': y.numpy()})
        # the weights TFP's AutoregressiveNetwork initialised (masks already applied by its constraint), in layer order
        for j, w in enumerate(flow.weights):
            res['%sw%d_%s' % (tag, j, w.name.replace('/', '.'))] = w.numpy()
        kwargs = {b.name: {'conditional_input': tf.constant(c)} for b in flow.chain.bijectors} if cond else {}
        res[tag + 'x_back'] = flow.chain.inverse(y, **kwargs).numpy()
        res[tag + 'fldj'] = flow.chain.forward_log_det_jacobian(tf.constant(x), event_ndims=1, **kwargs).numpy()
    np.savez_compressed(os.path.join(out, 'tfp_maf.npz'), **res)


def dump_dists(tf, tfp, vms, out):
    res = {}
    rng = np.random.default_rng(400)
    p = rng.normal(size=(64, 3)).astype(np.float32)
    for name, cls in (('normal', tfp.distributions.Normal), ('vonmises', tfp.distributions.VonMises)):
        t = vms.dists.make_param_transform(cls)(tf.constant(p))
        for k, v in t.items():
            res['transform_%s_%s' % (name, k)] = v.numpy()
    res['transform_in'] = p
    layer = vms.dists.IndependentBlockwise(3, [tfp.distributions.Normal, tfp.distributions.VonMises, tfp.distributions.Normal])
    params = rng.normal(size=(64, layer.params_size())).astype(np.float32)
    x = rng.uniform(-3, 3, (64, 3)).astype(np.float32)
    res.update(blockwise_params=params, blockwise_x=x, blockwise_log_prob=layer(tf.constant(params)).log_prob(tf.constant(x)).numpy())
    ivm = vms.dists.IndependentVonMises(4)
    pv = rng.normal(size=(64, int(ivm.params_size(4)))).astype(np.float32)
    xv = rng.uniform(-np.pi, np.pi, (64, 4)).astype(np.float32)
    res.update(vonmises_params=pv, vonmises_x=xv, vonmises_log_prob=ivm(tf.constant(pv)).log_prob(tf.constant(xv)).numpy())
    np.savez_compressed(os.path.join(out, 'tfp_dists.npz'), **res)


def dump_vae(tf, tfp, vms, out):
    from oracle import vae as ovae
    res = {}
    for prior in ('normal', 'realnvp'):
        P = ovae.init_vae(1003, prior=prior, flow_hidden=16, num_bins=8, hidden=32)
        dx, dz = P['dx'], P['dz']
        enc = vms.models.MappingToDistribution(tfp.layers.IndependentNormal(dz), mapping=vms.mappings.FCDeepNN(2 * dz, hidden_dim=32))
        dec = vms.models.MappingToDistribution(tfp.layers.IndependentNormal(dx), mapping=vms.mappings.FCDeepNN(2 * dx, hidden_dim=32))
        latent = tfp.layers.DistributionLambda(
            lambda t: tfp.distributions.Independent(tfp.distributions.Normal(tf.zeros((tf.shape(t)[0], dz)), 1.0), 1))
        if prior == 'normal':
            pr = latent
        else:
            flow = vms.flows.RQSSplineRealNVP(num_blocks=len(P['flow']), rqs_params=dict(num_bins=8, hidden_dim=16))
            pr = vms.dists.FlowedDistribution(flow, latent)
        model = vms.models.VAE(enc, dec, pr)
        x = np.random.default_rng(1001).standard_normal((64, dx)).astype(np.float32)
        eps = np.random.default_rng(1002).standard_normal((64, dz)).astype(np.float32)
        model(tf.constant(x))  # build
        for lay, (W, b) in zip([l for l in enc.mapping.layer_list if l.weights], P['enc']):
            lay.set_weights([W, b])
        for lay, (W, b) in zip([l for l in dec.mapping.layer_list if l.weights], P['dec']):
            lay.set_weights([W, b])
        if prior != 'normal':
            nb = len(P['flow'])
            for i, blk in enumerate(P['flow']):
                _assign_spline_net(flow.chain.bijectors[nb - 1 - i].bijector_fn, blk)
        with tf.GradientTape() as tape:
            qd = enc(tf.constant(x))
            loc, scale = qd.distribution.loc, qd.distribution.scale
            z = loc + scale * tf.constant(eps)  # the reparameterised sample with FIXED noise (models.py:310 draws it)
            logq = qd.log_prob(z)
            logpz = pr(z).log_prob(z)
            logpx = dec(z).log_prob(tf.constant(x))
            kl = tf.reduce_mean(logq - logpz)
            nll = -tf.reduce_mean(logpx)
            loss = nll + kl
        grads = tape.gradient(loss, model.trainable_variables)
        tag = 'c1_' if prior == 'normal' else 'c2_'
        res.update({tag + 'x': x, tag + 'eps': eps, tag + 'z': z.numpy(), tag + 'logq': logq.numpy(), tag + 'logpz': logpz.numpy(),
                    tag + 'logpx': logpx.numpy(), tag + 'scalars': np.array([loss.numpy(), nll.numpy(), kl.numpy()], np.float32)})
        for v_, g in zip(model.trainable_variables, grads):
            res[tag + 'grad_' + v_.name.replace('/', '.')] = g.numpy()
        res[tag + 'theta'] = ovae.flatten(ovae.param_list(P))
    np.savez_compressed(os.path.join(out, 'tfp_vae.npz'), **res)


def dump_distsel(tf, tfp, vms, out):
    rng = np.random.default_rng(500)
    B, N, L = 8, 300, 10.0
    coords = rng.uniform(0, L, (B, N, 3)).astype(np.float32)
    coords[:, 5] = coords[:, 3]
    ref = rng.uniform(0, L, (B, 1, 3)).astype(np.float32)
    info = rng.normal(size=(B, N, 2)).astype(np.float32)
    layer = vms.mappings.DistanceSelection(3.0, max_included=50, box_lengths=np.array([L, L, L], np.float32))
    sel, sinfo = layer(tf.constant(coords), tf.constant(ref), particle_info=tf.constant(info))
    # the indices tf.math.top_k returns inside the layer (mappings.py:433), recomputed with the layer's own ops
    local = coords - ref
    local = local - L * tf.round(local / L)
    d2 = tf.reduce_sum(local * local, axis=-1)
    _, idx = tf.math.top_k(-d2, k=50)
    np.savez_compressed(os.path.join(out, 'tfp_distsel.npz'), coords=coords, ref=ref, info=info, box=np.float32(L),
                        select=sel.numpy(), select_info=sinfo.numpy(), indices=idx.numpy())


def dump_gaa(tf, tfp, vms, out):
    """Needs `geometric_algebra_attention` (pip install geometric-algebra-attention), as `vaemolsim.mappings` itself does."""
    rng = np.random.default_rng(600)
    res = {}
    B, n, P = 5, 12, 9
    coords = rng.uniform(-3, 3, (B, n, 3)).astype(np.float32)
    info = np.round(rng.uniform(0, 1, (B, n, P))).astype(np.float32)
    for b in range(B):  # DistanceSelection-style zero padding; cloud 1 is padding only
        k = n if b == 1 else b
        if k:
            coords[b, n - k:] = 0
            info[b, n - k:] = 0
    res['coords'], res['info'] = coords, info

    def randomise(layer):  # Keras initialisers give LayerNormalization gamma = 1, beta = 0 and zero biases: exercise them
        for var in layer.variables:
            var.assign(var + tf.constant(rng.normal(0, 0.1, var.shape).astype(np.float32)))

    def store(prefix, layer):
        res[prefix + '_names'] = np.array([v_.name for v_ in layer.variables])
        for i, var in enumerate(layer.variables):
            res['%s_var%d' % (prefix, i)] = var.numpy()

    blk = vms.mappings.AttentionBlock(hidden_dim=16)
    blk([tf.constant(coords), tf.constant(info)])
    randomise(blk)
    store('block', blk)
    res['block_out'] = blk([tf.constant(coords), tf.constant(info)]).numpy()
    mask = tf.reduce_any(tf.not_equal(tf.constant(coords), 0.0), axis=-1)
    res['block_out_masked'] = blk([tf.constant(coords), tf.constant(info)], mask=[mask, None]).numpy()
    for mz in (True, False):
        pe = vms.mappings.ParticleEmbedding(10, hidden_dim=16, num_blocks=2, mask_zero=mz)
        pe(tf.constant(coords), tf.constant(info))
        randomise(pe)
        tag = 'embed_mask' if mz else 'embed_nomask'
        store(tag, pe)
        res[tag + '_out'] = pe(tf.constant(coords), tf.constant(info)).numpy()
    np.savez_compressed(os.path.join(out, 'tfp_gaa.npz'), **res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden'))
    args = ap.parse_args()
    tf, tfp, vms = need_tf(args.reference)
    sys.path.insert(0, ROOT)
    os.makedirs(args.out, exist_ok=True)
    for fn in (dump_rqs, dump_realnvp, dump_maf, dump_dists, dump_vae, dump_distsel, dump_gaa):
        try:
            fn(tf, tfp, vms, args.out)
            print('wrote', fn.__name__)
        except ImportError as ex:  # dump_gaa without geometric_algebra_attention installed
            print('skipped %s: %s' % (fn.__name__, ex))


if __name__ == '__main__':
    main()
