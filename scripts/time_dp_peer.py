"""Development aid (2+ GPUs, one process per GPU): back-to-back data-parallel C2 steps through `PeerExchange.train_step` and the
finish / exchange kernel's own timeline of the last step.  torchrun --nproc-per-node 2 scripts/time_dp_peer.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402
from vaemolsim_b200 import parallel  # noqa: E402

grp = parallel.Group()
c = v._abi.ctx()
w = bench.WORKLOADS['c2']
B = 4096
model = bench.build_model(v, w, B)
f = model.fused(B)
rng = np.random.default_rng(1 + grp.rank)
x = v.Tensor.from_numpy(rng.standard_normal((B, 6), dtype=np.float32))
e = v.Tensor.from_numpy(rng.standard_normal((B, 2), dtype=np.float32))
opt = model.optimizer
pe = parallel.PeerExchange(grp, f.n_params)
for _ in range(50):
    pe.train_step(f, x, e, B, opt)
c.synchronize()
grp.barrier()
t0 = time.perf_counter()
n = 500
for _ in range(n):
    pe.train_step(f, x, e, B, opt)
c.synchronize()
dt = (time.perf_counter() - t0) / n * 1e6
tr = pe.buf.numpy().view(np.uint64)[pe.n + 56:pe.n + 60].astype(np.int64)
print('rank %d: %.1f us per step; finish kernel: partial sums %.1f us, wait for the peers %.1f us, pull + Adam + images %.1f us'
      % (grp.rank, dt, (tr[1] - tr[0]) / 1e3, (tr[2] - tr[1]) / 1e3, (tr[3] - tr[2]) / 1e3), flush=True)
pe.close()
grp.close()
