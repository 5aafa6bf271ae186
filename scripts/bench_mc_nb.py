"""Development aid: device-resident timing of the fused notebook-family MC kernel (vms_mc_nb_run), 65,536 chains x 100 steps,
for the lanes-per-chain variants: arguments `tpc[:chains]`, e.g. `1 4 4:8192 auto`."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
model = bench.build_c4b_model(v)
x0 = bench.gmm_start(65536)
for arg in sys.argv[1:] or ['auto']:
    tpc, _, nch = arg.partition(':')
    nch, _, occ_ = nch.partition(':')
    nch = int(nch or 65536)
    os.environ['VMS_NB_OCC'] = occ_ or '2'
    if tpc == 'auto':
        os.environ.pop('VMS_NB_TPC', None)
    else:
        os.environ['VMS_NB_TPC'] = tpc
    occ = 'tpc %s, %d chains, occ %s' % (tpc, nch, occ_ or '2')
    mc = v.mcmc.MCMC(model, v.mcmc.GaussianMixtureEnergy(), random_seed=5002)
    xd = v.Tensor.from_numpy(np.ascontiguousarray(x0[:nch]))
    xd, ed = mc.run_nb(None, n_steps=100, configs_dev=xd)
    c.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        mc.run_nb(None, n_steps=100, configs_dev=xd, energies_dev=ed)
    c.synchronize()
    dt = (time.perf_counter() - t0) / 3
    mc.sync_counters()
    print('%s: %.4f ms / MC step, %.1f M proposals/s, acceptance %.4f' % (occ, dt * 10, nch * 100 / dt / 1e6, mc.acceptance_rate),
          flush=True)
