"""Development aid: device-resident timing of the C4a fused MC kernel (mc_chain.cu) for lane counts / chain counts:
arguments `tpc[:chains[:chains_per_lane]]`, e.g. `2 4 2:8192 4:65536:2 auto:16384`."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
model = bench.build_model(v, bench.WORKLOADS['c1'], 4096)
x0 = np.random.default_rng(4001).standard_normal((65536, 6), dtype=np.float32)
for arg in sys.argv[1:] or ['auto']:
    tpc, _, nch = arg.partition(':')
    nch, _, cpl = nch.partition(':')
    nch = int(nch or 65536)
    if cpl or tpc != 'auto':
        os.environ['VMS_MC_CPL'] = cpl or '1'
    else:
        os.environ.pop('VMS_MC_CPL', None)  # `auto` without a third field: the launcher's own choice
    tpc_label = tpc
    if tpc == 'auto':
        os.environ.pop('VMS_MC_TPC', None)
    else:
        os.environ['VMS_MC_TPC'] = tpc
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002, stream_layout=(0, 65536))
    xd = v.Tensor.from_numpy(np.ascontiguousarray(x0[:nch]))
    xd, ed = mc.run_fused(None, n_steps=100, configs_dev=xd)
    c.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        mc.run_fused(None, n_steps=100, configs_dev=xd, energies_dev=ed)
    c.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print('tpc %s cpl %s, %d chains: %.4f ms / MC step, %.1f M proposals/s' % (tpc, cpl or '1', nch, dt * 10, nch * 100 / dt / 1e6), flush=True)
