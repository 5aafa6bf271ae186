"""cProfile of the op-by-op device-resident MC loop (MCMC.run_device) on the notebook model."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = bench.build_c4b_model(v)
mc = v.mcmc.MCMC(model, v.mcmc.GaussianMixtureEnergy(), random_seed=1)
mc.fuse_notebook = False
x = v.Tensor.from_numpy(bench.gmm_start(n))
x, e = mc.run_device(None, n_steps=3, configs_dev=x)
pr = cProfile.Profile()
pr.enable()
x, e = mc.run_device(None, n_steps=20, configs_dev=x, energies_dev=e)
v._abi.ctx().synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
