"""A few ParticleEmbedding forward passes over the descriptors of the C3 box (for ncu): python scripts/prof_gaa.py [k]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v  # noqa: E402

c = v._abi.ctx()
k = int(sys.argv[1]) if len(sys.argv) > 1 else 50
rng = np.random.default_rng(3001)
B, N, L = 4096, 10000, np.float32(46.416)
frame = v.Tensor.from_numpy(rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32))
info = v.Tensor.from_numpy(np.eye(2, dtype=np.float32)[rng.integers(0, 2, N)])
ref = v.Tensor.from_numpy(np.random.default_rng(3002).uniform(-L / 2, L / 2, (B, 3)).astype(np.float32))
sel = v.mappings.DistanceSelection(3.0, max_included=k, box_lengths=np.array([L, L, L], np.float32))
v.set_seed(8)
pe = v.mappings.ParticleEmbedding(20)
xyz, inf = sel.select_from_frame(frame, ref, particle_info=info)
for _ in range(2):
    out = pe(xyz, inf)
c.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    out = pe(xyz, inf)
c.synchronize()
print('k = %d: %.3f ms per embedding of %d sites' % (k, (time.perf_counter() - t0) / 5 * 1e3, B), out.shape, flush=True)
