"""Development aid: what the data-parallel form of the C2 step costs on ONE GPU, without the exchange: (a) train_step (tile
kernel + finish with Adam, packed weights written by the finish kernel) against (b) forward_backward + a separate Adam
kernel (pre-pack + tile kernel + finish + Adam: the launch structure of the N > 1 step minus the peer wait)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
w = bench.WORKLOADS['c2']
B = 4096
model = bench.build_model(v, w, B)
f = model.fused(B)
rng = np.random.default_rng(1)
x = v.Tensor.from_numpy(rng.standard_normal((B, 6), dtype=np.float32))
e = v.Tensor.from_numpy(rng.standard_normal((B, 2), dtype=np.float32))
opt = model.optimizer


def timed(fn, n=300):
    for _ in range(20):
        fn()
    c.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    c.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def dp_like():
    f.forward_backward(x, e)
    f.adam_step(opt)


print('path', f.path(B))
print('train_step                      %.1f us' % timed(lambda: f.train_step(x, e, opt)))
print('forward_backward + adam_step    %.1f us' % timed(dp_like))
print('forward_backward only (theta fixed: no pre-pack after the first) %.1f us' % timed(lambda: f.forward_backward(x, e)))
