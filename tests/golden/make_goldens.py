"""Generates the committed golden fixtures under tests/golden/.  Run HERE (the build container), never on the GPU box:

    python tests/golden/make_goldens.py

* mcmc_reference.npz -- produced by the REFERENCE's own `vaemolsim/mcmc.py` (loaded by file path from /root/reference;
  it imports only NumPy) driving `oracle.mcmc.OracleVAE`: per-step accept decisions, configurations, energies, counters
  and the log-probability traces fed to the acceptance rule.  Pins `oracle.mcmc.single_step` and `vms_mc_accept`.
* mcmc_reference_c4b.npz -- the same driver on `oracle.mcmc.OracleVAEb` (the MC notebook's model family: MAF prior,
  conditional autoregressive decoder) with the notebook's Gaussian-mixture log-density as energy callback.
* maf_orders.json    -- MAF block input orders computed with the reference's recipe (flows.py:606-621) for the seeds
  named in SURVEY 8c item 10.
* elbo_c1.npz / elbo_c2.npz -- oracle ELBO forward / backward outputs on seeded inputs (regression fixtures for the
  oracle itself: the reference's TF/TFP arithmetic is not importable here, SURVEY 8c -- "parity unpinned").
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))

from oracle import mcmc as omc  # noqa: E402
from oracle import nets, vae  # noqa: E402

REF = '/root/reference/vaemolsim/mcmc.py'


def load_reference_mcmc():
    spec = importlib.util.spec_from_file_location('ref_mcmc', REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_mcmc(prior, tag, n_chains=64, n_steps=5):
    ref = load_reference_mcmc()
    P = vae.init_vae(1003, prior=prior, flow_hidden=16, num_bins=8, hidden=32)
    x0 = np.random.default_rng(4001).standard_normal((n_chains, 6)).astype(np.float32)

    class Traced(omc.OracleVAE):
        pass

    model = Traced(P, noise_seed=777)
    mc = ref.MCMC(model, omc.quadratic_energy, random_seed=4002)
    # run the reference driver step by step, re-deriving its internals with an identical twin for the traces
    twin = omc.OracleVAE(P, noise_seed=777)
    twin_rng = np.random.default_rng(4002)
    configs, energies = x0, None
    tconfigs, tenergies = x0, None
    out = {}
    for s in range(n_steps):
        configs, energies = mc.single_step(configs, energies=energies)
        trace = {}
        tconfigs, tenergies, acc = omc.single_step(twin, omc.quadratic_energy, twin_rng, tconfigs, tenergies, trace)
        assert np.array_equal(configs, tconfigs) and np.array_equal(energies, tenergies), 'oracle != reference mcmc.py'
        out['configs_%d' % s] = configs.copy()
        out['energies_%d' % s] = energies.copy()
        out['acc_%d' % s] = acc
        for k, v in trace.items():
            out['%s_%d' % (k, s)] = v
    out['num_trials'] = np.float64(mc._num_trials)
    out['num_acc'] = np.float64(mc._num_acc)
    out['x0'] = x0
    np.savez_compressed(os.path.join(HERE, 'mcmc_reference_%s.npz' % tag), **out)
    print(tag, 'acceptance', mc.acceptance_rate)


def gmm_sample(rng, n):
    """Draws from the notebook's data distribution (cell 5): the chains start in the mixture."""
    k = rng.choice(3, size=n, p=omc.GMM_PROBS.astype(np.float64) / omc.GMM_PROBS.astype(np.float64).sum())
    return (omc.GMM_LOCS[k] + omc.GMM_SCALES[k] * rng.standard_normal((n, 2))).astype(np.float32)


def make_mcmc_c4b(n_chains=256, n_steps=5, hidden=64):
    """C4b (MC_Moves_with_VAEs.ipynb cell 39-41): the reference's mcmc.py driving the notebook's model family -- MAF prior,
    conditional AutoregressiveBlockwise decoder, Gaussian-mixture log-density as the energy callback."""
    ref = load_reference_mcmc()
    P = omc.init_vae_b(2003, hidden=hidden)
    x0 = gmm_sample(np.random.default_rng(5001), n_chains)
    mc = ref.MCMC(omc.OracleVAEb(P, noise_seed=888), omc.gmm_energy, random_seed=5002)
    twin = omc.OracleVAEb(P, noise_seed=888)
    twin_rng = np.random.default_rng(5002)
    configs, energies = x0, None
    tconfigs, tenergies = x0, None
    out = {}
    for s in range(n_steps):
        configs, energies = mc.single_step(configs, energies=energies)
        trace = {}
        tconfigs, tenergies, acc = omc.single_step(twin, omc.gmm_energy, twin_rng, tconfigs, tenergies, trace)
        assert np.array_equal(configs, tconfigs) and np.array_equal(energies, tenergies), 'oracle != reference mcmc.py'
        out['configs_%d' % s] = configs.copy()
        out['energies_%d' % s] = energies.copy()
        out['acc_%d' % s] = acc
        for k, v in trace.items():
            out['%s_%d' % (k, s)] = v
    out['num_trials'] = np.float64(mc._num_trials)
    out['num_acc'] = np.float64(mc._num_acc)
    out['x0'] = x0
    np.savez_compressed(os.path.join(HERE, 'mcmc_reference_c4b.npz'), **out)
    print('c4b acceptance', mc.acceptance_rate)


def make_orders():
    res = {}
    for nb, D, seed in ((3, 3, 42), (4, 6, 42), (4, 2, 7)):
        res['%d_%d_%d' % (nb, D, seed)] = [o if isinstance(o, str) else [int(v) for v in o]
                                          for o in nets.maf_block_orders(nb, D, seed)]
    json.dump(res, open(os.path.join(HERE, 'maf_orders.json'), 'w'), indent=1)


def make_elbo(prior, tag, B=64):
    P = vae.init_vae(1003, prior=prior, flow_hidden=16, num_bins=8, hidden=32)
    x = np.random.default_rng(1001).standard_normal((B, 6)).astype(np.float32)
    eps = np.random.default_rng(1002).standard_normal((B, 2)).astype(np.float32)
    out, G = vae.elbo_backward(P, x, eps)
    np.savez_compressed(os.path.join(HERE, 'elbo_%s.npz' % tag), x=x, eps=eps, theta=vae.flatten(vae.param_list(P)),
                        grad=vae.flatten(vae.grad_list(P, G)), z=out['z'], logq=out['logq'], logpz=out['logpz'],
                        logpx=out['logpx'], scalars=np.array([out['loss'], out['nll'], out['kl']], np.float32))


if __name__ == '__main__':
    make_mcmc('normal', 'c4a')
    make_mcmc('realnvp', 'flow')
    make_mcmc_c4b()
    make_orders()
    make_elbo('normal', 'c1')
    make_elbo('realnvp', 'c2')
