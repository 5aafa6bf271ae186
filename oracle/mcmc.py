"""MC acceptance oracle -- NumPy restatement of `vaemolsim/mcmc.py:68-130`, test infrastructure only.

PINNED: `tests/golden/make_goldens.py` runs the reference's own `vaemolsim/mcmc.py` (imports only NumPy, mcmc.py:5) on
top of `OracleVAE` below and commits the decisions; `single_step` here is checked against them bit for bit
(tests/test_oracle_mcmc.py).  The oracle VAE takes its sampling noise from a seeded NumPy generator so that the reference
driver, this restatement and the CUDA path can all be fed identical log-probabilities.
"""
import numpy as np

from . import dists, vae


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
