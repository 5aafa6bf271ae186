"""MC acceptance oracle -- NumPy restatement of `vaemolsim/mcmc.py:68-130`, test infrastructure only.

PINNED: `tests/golden/make_goldens.py` runs the reference's own `vaemolsim/mcmc.py` (imports only NumPy, mcmc.py:5) on
top of `OracleVAE` below and commits the decisions; `single_step` here is checked against them bit for bit
(tests/test_oracle_mcmc.py).  The oracle VAE takes its sampling noise from a seeded NumPy generator so that the reference
driver, this restatement and the CUDA path can all be fed identical log-probabilities.
"""
import numpy as np

from . import dists, flows, nets, vae


def accept(e_new, e_old, forward_log_p, reverse_log_p, log_rand):
    """mcmc.py:116-120: log_acc = new_energies + reverse_log_p - energies - forward_log_p; acc = log_acc >= log_rand."""
    log_acc = e_new + reverse_log_p - e_old - forward_log_p
    return log_acc >= log_rand


class _T(object):
    """Tensor stand-in with `.numpy()` and `+` (what mcmc.py:103,109,112 use)."""

    def __init__(self, a):
        self.a = np.asarray(a)

    def numpy(self):
        return self.a.copy()

    def __add__(self, o):
        return _T(self.a + (o.a if isinstance(o, _T) else o))

    @property
    def shape(self):
        return self.a.shape


class _NormalDist(object):

    def __init__(self, loc, scale, rng):
        self.loc, self.scale, self.rng = loc, scale, rng

    def log_prob(self, x):
        x = x.a if isinstance(x, _T) else np.asarray(x, np.float32)
        return _T(dists.normal_log_prob(x.astype(np.float32), self.loc, self.scale).sum(-1).astype(np.float32))

    def experimental_sample_and_log_prob(self):
        eps = self.rng.standard_normal(self.loc.shape, dtype=np.float32)
        z = dists.normal_sample(self.loc, self.scale, eps)
        return _T(z), self.log_prob(z)


class _PriorDist(object):

    def __init__(self, P, batch, rng):
        self.P, self.batch, self.rng = P, batch, rng

    def log_prob(self, z):
        z = z.a if isinstance(z, _T) else np.asarray(z, np.float32)
        return _T(vae.prior_log_prob(self.P, z.astype(np.float32)))

    def experimental_sample_and_log_prob(self):
        eps = self.rng.standard_normal((self.batch, self.P['dz']), dtype=np.float32)
        y, lp = vae.prior_sample_and_log_prob(self.P, eps)
        return _T(y), _T(lp)


class OracleVAE(object):
    """Duck-typed `vae` for the reference's MCMC driver: `.encoder(x)`, `.prior(z)`, `.decoder(z)` return distributions
    with `log_prob` and `experimental_sample_and_log_prob` (mcmc.py:100-108)."""

    def __init__(self, P, noise_seed):
        self.P = P
        self.rng = np.random.default_rng(noise_seed)

    def _arr(self, x):
        return (x.a if isinstance(x, _T) else np.asarray(x)).astype(np.float32)

    def encoder(self, x):
        loc, scale, _, _ = vae.encoder_dist(self.P, self._arr(x))
        return _NormalDist(loc, scale, self.rng)

    def decoder(self, z):
        loc, scale, _, _ = vae.decoder_dist(self.P, self._arr(z))
        return _NormalDist(loc, scale, self.rng)

    def prior(self, z):
        return _PriorDist(self.P, self._arr(z).shape[0], self.rng)


# ----------------------------------------------------------------------------- C4b: the MC notebook's model
def init_vae_b(seed, hidden=200, num_blocks=4, num_bins=20, flow_hidden=40, made_hidden=(10, 100, 10),
               bin_range=(-10.0, 10.0), bias_scale=0.3):
    """Weights of the VAE of examples/MC_Moves_with_VAEs.ipynb (cells 11-20): 2-D data, 1-D latent.
    encoder  FCDeepNN(2 -> hidden -> 2) into tfp.layers.IndependentNormal(1)                      (cell 11)
    prior    RQSSplineMAF(4 blocks, 20 bins, hidden 40, range [-10, 10]) over Independent N(0, 1)  (cell 14)
    decoder  FCDeepNN(1 -> hidden -> (2, 2)) into AutoregressiveBlockwise(2, [Normal]*2, conditional on z,
             MADE hidden_units [10, 100, 10], default (linear) activation)                          (cell 17)
    Keras initialises every bias with zeros, which would make the 1-D MAF the identity: the biases get N(0, bias_scale)
    values so that the test model exercises every term."""
    rng = np.random.default_rng(seed)
    enc = nets.fcdeepnn_init(rng, 2, [hidden], (2, ))
    dec = nets.fcdeepnn_init(rng, 1, [hidden], (2, 2))
    made = nets.made_init(rng, 2, 2, list(made_hidden), 'left-to-right', cond_size=1)
    maf = flows.maf_init(rng, 1, num_blocks, None, num_bins, flow_hidden)
    shake = lambda b: (b + rng.normal(0, bias_scale, b.shape)).astype(np.float32)
    enc = [(W, shake(b)) for W, b in enc]
    dec = [(W, shake(b)) for W, b in dec]
    # an untrained decoder proposes far from the mixture and nearly every move is rejected; the output layer is aimed at
    # the mixture's main component (loc (-0.5, 0), scales about (0.1, 0.5)) so that the fixture holds both outcomes
    Wl, bl = dec[-1]
    inv_softplus = lambda y: np.log(np.expm1(y))
    dec[-1] = ((Wl * np.float32(0.2)).astype(np.float32),
               np.array([-0.5, inv_softplus(0.1), 0.0, inv_softplus(0.5)], np.float32))
    for k, L in enumerate(made):
        L['b'] = shake(L['b'])
        if k == len(made) - 1:  # ... and the autoregressive shift of the parameters kept moderate
            L['W'], L['Wc'], L['b'] = (np.float32(0.2) * a for a in (L['W'], L['Wc'], L['b']))
    for blk in maf:
        for key in ('w', 'h', 's'):
            for L in blk[key]:
                L['b'] = shake(L['b'])
    return dict(enc=enc, dec=dec, made=made, maf=maf, num_bins=num_bins, bin_range=tuple(bin_range), hidden=hidden)


class _MafPriorDist(object):
    """FlowedDistribution(RQSSplineMAF, Independent N(0, 1)) (dists.py:343-420): TransformedDistribution of the chain."""

    def __init__(self, P, batch, rng):
        self.P, self.batch, self.rng = P, batch, rng

    def log_prob(self, z):
        z = (z.a if isinstance(z, _T) else np.asarray(z)).astype(np.float32)
        x, ildj = flows.maf_inverse(z, self.P['maf'], self.P['num_bins'], self.P['bin_range'])
        return _T((dists.normal_log_prob(x, np.float32(0), np.float32(1)).sum(-1) + ildj).astype(np.float32))

    def experimental_sample_and_log_prob(self):
        eps = self.rng.standard_normal((self.batch, 1), dtype=np.float32)
        y, fldj = flows.maf_forward(eps, self.P['maf'], self.P['num_bins'], self.P['bin_range'])
        lp = dists.normal_log_prob(eps, np.float32(0), np.float32(1)).sum(-1) - fldj
        return _T(y), _T(lp.astype(np.float32))


class _AutoregressiveNormalDist(object):
    """AutoregressiveBlockwise(2, [Normal]*2, conditional=True)(inputs, conditional_input=z) (dists.py:303-340)."""

    def __init__(self, P, inputs, cond, rng):
        self.P, self.inputs, self.cond, self.rng = P, inputs, cond, rng

    def log_prob(self, x):
        x = (x.a if isinstance(x, _T) else np.asarray(x)).astype(np.float32)
        return _T(dists.autoregressive_blockwise_log_prob(x, self.inputs, self.P['made'], ['normal', 'normal'],
                                                          cond=self.cond).astype(np.float32))

    def experimental_sample_and_log_prob(self):
        eps = self.rng.standard_normal(self.inputs.shape[:2], dtype=np.float32)
        x = dists.autoregressive_blockwise_sample_normal(self.inputs, self.P['made'], eps, cond=self.cond)
        return _T(x), self.log_prob(x)


class OracleVAEb(object):
    """Duck-typed notebook VAE for the reference's MCMC driver (same protocol as `OracleVAE`)."""

    def __init__(self, P, noise_seed):
        self.P = P
        self.rng = np.random.default_rng(noise_seed)

    def _arr(self, x):
        return (x.a if isinstance(x, _T) else np.asarray(x)).astype(np.float32)

    def encoder(self, x):
        prm = nets.fcdeepnn_forward(self._arr(x), self.P['enc'], (2, ))
        loc, scale = dists.independent_normal_params(prm, 1)
        return _NormalDist(loc, scale, self.rng)

    def prior(self, z):
        return _MafPriorDist(self.P, self._arr(z).shape[0], self.rng)

    def decoder(self, z):
        z = self._arr(z)
        return _AutoregressiveNormalDist(self.P, nets.fcdeepnn_forward(z, self.P['dec'], (2, 2)), z, self.rng)


GMM_PROBS = np.array([0.7, 0.2, 0.1], np.float32)
GMM_LOCS = np.array([[-0.5, 0.0], [1.0, 2.0], [-1.5, 0.0]], np.float32)
GMM_SCALES = np.array([[0.05, 0.5], [1.0, 0.5], [0.5, 0.2]], np.float32)


def gmm_energy(configs):
    """MC_Moves_with_VAEs.ipynb cell 5 + 38: `data_dist.log_prob(configs).numpy()` for
    Mixture(Categorical(probs), [Independent(Normal(loc_k, scale_k), 1)]): logsumexp_k(log_prob_k(x) + log p_k), float32."""
    x = np.asarray(configs, np.float32)
    lp = np.stack([dists.normal_log_prob(x, m, s).sum(-1, dtype=np.float32) + np.log(p)
                   for p, m, s in zip(GMM_PROBS, GMM_LOCS, GMM_SCALES)], axis=-1).astype(np.float32)
    mx = lp.max(axis=-1, keepdims=True)
    return (mx[..., 0] + np.log(np.sum(np.exp(lp - mx), axis=-1, dtype=np.float32))).astype(np.float32)


def quadratic_energy(configs):
    """tests/test_mcmc.py:28-32."""
    means = np.linspace(-2, 2, configs.shape[-1])[np.newaxis, :]
    return np.sum((configs - means)**2, axis=-1)


def single_step(vae_obj, energy_func, rng, configs, energies=None, trace=None):
    """Restatement of MCMC.single_step (mcmc.py:68-130).  Returns (new_configs, new_energies, acc_bool)."""
    configs = np.array(configs)
    if energies is None:
        energies = energy_func(configs)
    z1, l1 = vae_obj.encoder(configs).experimental_sample_and_log_prob()
    z2, l2 = vae_obj.prior(z1).experimental_sample_and_log_prob()
    x2, l3 = vae_obj.decoder(z2).experimental_sample_and_log_prob()
    forward_log_p = (l1 + l2 + l3).numpy()
    r1 = vae_obj.encoder(x2).log_prob(z2)
    r2 = vae_obj.prior(z2).log_prob(z1)
    r3 = vae_obj.decoder(z1).log_prob(configs)
    reverse_log_p = (r1 + r2 + r3).numpy()
    new_configs = x2.numpy()
    new_energies = energy_func(new_configs)
    log_rand = np.log(rng.random(size=forward_log_p.shape[0]))
    acc = accept(new_energies, energies, forward_log_p, reverse_log_p, log_rand)
    if trace is not None:
        trace.update(e_new=new_energies.copy(), e_old=np.asarray(energies).copy(), fwd=forward_log_p, rev=reverse_log_p,
                     log_rand=log_rand, x_new=new_configs.copy(), x_old=configs.copy(), z1=z1.numpy(), z2=z2.numpy())
    new_configs[~acc, ...] = configs[~acc, ...]
    new_energies[~acc] = np.asarray(energies)[~acc]
    return new_configs, new_energies, acc
