"""Small invocations of every fused kernel family, for `compute-sanitizer --tool racecheck|memcheck python scripts/sanitize_targets.py`
(the sanitizer slows kernels 10-100x: sizes are tiny, a few tiles / chains each)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
rng = np.random.default_rng(0)
which = sys.argv[1:] or ['elbo', 'mc', 'nb', 'distsel']

if 'elbo' in which:
    w = bench.WORKLOADS['c2']
    for mode, B in ((4, 96), (3, 96), (2, 192), (1, 64)):
        model = bench.build_model(v, w, B)
        f = model.fused(B)
        f.set_mode(mode)
        x = v.Tensor.from_numpy(rng.standard_normal((B, 6)).astype(np.float32))
        eps = v.Tensor.from_numpy(rng.standard_normal((B, 2)).astype(np.float32))
        out = f.forward_backward(x, eps)
        c.synchronize()
        print('elbo mode', mode, f.path(B), out.numpy(), flush=True)
if 'mc' in which:
    model = bench.build_model(v, bench.WORKLOADS['c1'], 256)
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=1)
    x, e = mc.run(rng.standard_normal((200, 6)).astype(np.float32), n_steps=3)
    print('mc_chain', mc.acceptance_rate, flush=True)
if 'nb' in which:
    model = bench.build_c4b_model(v)
    for tpc in ('1', '4'):
        os.environ['VMS_NB_TPC'] = tpc
        mc = v.mcmc.MCMC(model, v.mcmc.GaussianMixtureEnergy(), random_seed=1)
        x, e = mc.run(bench.gmm_start(200), n_steps=3)
        print('mc_nb tpc', tpc, mc.acceptance_rate, flush=True)
if 'distsel' in which:
    m = v.mappings.DistanceSelection(6.0, max_included=10, box_lengths=np.array([20.0, 20.0, 20.0], np.float32))
    coords = rng.uniform(0, 20, (4, 700, 3)).astype(np.float32)
    ref = rng.uniform(0, 20, (4, 3)).astype(np.float32)
    out = m(coords, ref)
    out = out[0] if isinstance(out, (tuple, list)) else out
    print('dist_select', out.numpy().shape, flush=True)
print('ok')
