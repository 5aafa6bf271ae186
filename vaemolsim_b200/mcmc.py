"""VAE-based Monte Carlo -- host-side mirror of `vaemolsim/mcmc.py` with the accept / reject step on the device.

Same class, methods and counters as the reference (mcmc.py:12 `MCMC`, :48 `acceptance_rate`, :52 `reset`, :68
`single_step`, :133 `run`).  The six distribution evaluations of a step (mcmc.py:100-108) are the kernels behind
`vae.encoder / prior / decoder`; the acceptance arithmetic of mcmc.py:116-128 is `vms_mc_accept` (float64, same
operation order).  The uniform stream stays NumPy's PCG64 + `np.log` on the host (mcmc.py:119) so decisions are
bit-identical to the reference under the same seed; 8 bytes per chain per step are uploaded.
"""
import numpy as np

from ._abi import Tensor, as_tensor, ctx


class MCMC(object):
    """Markov chain Monte Carlo with a VAE proposal: as many independent chains as input configurations."""

    def __init__(self, vae, energy_func, random_seed=None):
        self.vae = vae
        self.energy_func = energy_func
        self._num_trials = 0.0
        self._num_acc = 0.0
        self._rng = np.random.default_rng(seed=random_seed)

    @property
    def acceptance_rate(self):
        return self._num_acc / self._num_trials

    def reset(self, random_seed=None):
        self._num_trials = 0.0
        self._num_acc = 0.0
        self._rng = np.random.default_rng(seed=random_seed)

    def _energies(self, configs_host, configs_dev):
        """energy_func is the reference's host callable on NumPy arrays; a callable flagged `on_device` instead maps a
        device Tensor [B, D] float32 to a device Tensor [B] float64 (no PCIe round trip)."""
        if getattr(self.energy_func, 'on_device', False):
            return self.energy_func(configs_dev)
        if configs_host is None:
            configs_host = configs_dev.numpy()
        return Tensor.from_numpy(np.asarray(self.energy_func(configs_host), dtype=np.float64))

    def single_step(self, configs, energies=None):
        c = ctx()
        configs = np.array(configs.numpy() if isinstance(configs, Tensor) else configs)
        x1 = Tensor.from_numpy(configs, dtype=np.float32)
        B = x1.shape[0]
        x1 = x1.reshape(B, -1)
        if energies is None:
            e_old = self._energies(configs, x1)
        else:
            e_old = as_tensor(energies, dtype=np.float64)

        # forward proposal: encode, move in latent space, decode (mcmc.py:100-103)
        z1, log_z1_given_x1 = self.vae.encoder(x1).experimental_sample_and_log_prob()
        z2, log_z2 = self.vae.prior(z1).experimental_sample_and_log_prob()
        x2, log_x2_given_z2 = self.vae.decoder(z2).experimental_sample_and_log_prob()
        forward_log_p = log_z1_given_x1 + log_z2 + log_x2_given_z2

        # reverse proposal probability (mcmc.py:106-109)
        log_z2_given_x2 = self.vae.encoder(x2).log_prob(z2)
        log_z1 = self.vae.prior(z2).log_prob(z1)
        log_x1_given_z1 = self.vae.decoder(z1).log_prob(x1)
        reverse_log_p = log_z2_given_x2 + log_z1 + log_x1_given_z1

        e_new = self._energies(None, x2)

        # accept / reject on the device with the host's uniform stream (mcmc.py:116-128)
        log_rand = Tensor.from_numpy(np.log(self._rng.random(size=B)))
        x2 = x2.contig() if x2.contiguous else x2.contig()
        e_out = Tensor((B, ), np.float64)
        acc = Tensor((B, ), np.uint8)
        n_acc = Tensor.zeros((1, ), np.uint64)
        c.lib.vms_mc_accept(e_new.ptr, e_old.ptr, forward_log_p.ptr, reverse_log_p.ptr, log_rand.ptr, B, x2.shape[1],
                            x1.ptr, x2.ptr, e_out.ptr, acc.ptr, n_acc.ptr, c.stream)
        self._num_trials += B
        self._num_acc += float(n_acc.numpy()[0])
        self._last_acc = acc
        return x2.numpy().reshape(configs.shape), e_out.numpy()

    def run(self, configs, energies=None, n_steps=1):
        for n in range(n_steps):
            configs, energies = self.single_step(configs, energies=energies)
        return configs, energies
