"""Development aid: C2 training-step time of plan modes 3 (whole-step tcgen05 kernel) and 2 (per-block tensor-core plan) over
batch sizes, to place the automatic switch.  python scripts/time_modes.py [batch ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
w = bench.WORKLOADS['c2']
for B in [int(a) for a in sys.argv[1:]] or [4096, 6144, 8192, 9472, 12288, 16384, 24576]:
    model = bench.build_model(v, w, B)
    f = model.fused(B)
    rng = np.random.default_rng(1)
    x = v.Tensor.from_numpy(rng.standard_normal((B, 6), dtype=np.float32))
    e = v.Tensor.from_numpy(rng.standard_normal((B, 2), dtype=np.float32))
    res = []
    for mode in (3, 2):
        f.set_tc_auto_batch(1 << 40)
        try:
            f.set_mode(mode)
            for _ in range(10):
                f.train_step(x, e, model.optimizer)
            c.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                f.train_step(x, e, model.optimizer)
            c.synchronize()
            dt = (time.perf_counter() - t0) / 50
            res.append('%s %.3f ms (%.1f M/s)' % (f.path(B), dt * 1e3, B / dt / 1e6))
        except Exception as ex:
            res.append('mode %d: %s' % (mode, ex))
    print('B = %6d: %s' % (B, ' | '.join(res)), flush=True)
