"""A few shared-frame neighbour selections at the C3 shape (for ncu): python scripts/prof_distsel_frame.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
rng = np.random.default_rng(3001)
B, N, L = 4096, 10000, np.float32(46.416)
frame = v.Tensor.from_numpy(rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32))
info = v.Tensor.from_numpy(np.eye(2, dtype=np.float32)[rng.integers(0, 2, N)])
ref = v.Tensor.from_numpy(np.random.default_rng(3002).uniform(-L / 2, L / 2, (B, 3)).astype(np.float32))
layer = v.mappings.DistanceSelection(8.0, max_included=50, box_lengths=np.array([L, L, L], np.float32))
for _ in range(4):
    out = layer.select_from_frame(frame, ref, particle_info=info, return_indices=True)
c.synchronize()
print('ok', out[0].shape)
import time
for k in (50, 10):
    layer = v.mappings.DistanceSelection(3.0, max_included=k, box_lengths=np.array([L, L, L], np.float32))
    for idx in (True, False):
        for _ in range(3):
            layer.select_from_frame(frame, ref, particle_info=info if idx else None, return_indices=idx)
        c.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            layer.select_from_frame(frame, ref, particle_info=info if idx else None, return_indices=idx)
        c.synchronize()
        print('frame mode k=%d %s: %.4f ms' % (k, 'xyz+info+idx' if idx else 'xyz only', (time.perf_counter() - t0) / 20 * 1e3), flush=True)
