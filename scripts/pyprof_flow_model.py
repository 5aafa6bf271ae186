"""cProfile of FlowModel tape-path training steps (host overhead of the op-by-op path)."""
import cProfile
import os
import pstats
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v  # noqa: E402
import vaemolsim_b200._protocols as PR  # noqa: E402

v.set_seed(7)
flow = v.flows.RQSSplineRealNVP(num_blocks=4, rqs_params=dict(bin_range=[-10.0, 10.0], num_bins=32, hidden_dim=100))
fm = v.models.FlowModel(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 1)))
fm.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = (np.random.default_rng(21).normal(size=(B, 1)) * 1.5 + 0.5).astype(np.float32)
for _ in range(5):
    fm.train_on_batch(x, x)
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    fm.train_on_batch(x, x)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(18)
