"""Runs a few ELBO training steps of the C2 workload under a given plan mode (for ncu captures).
    python scripts/prof_step.py <mode> [batch] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402


def main():
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    w = bench.WORKLOADS['c2']
    c = v._abi.ctx()
    model = bench.build_model(v, w, B)
    f = model.fused(B)
    if mode != 0:
        f.set_tc_auto_batch(1 << 40)
        f.set_mode(mode)
    rng = np.random.default_rng(11)
    x = v.Tensor.from_numpy(rng.standard_normal((B, w['dx']), dtype=np.float32))
    e = v.Tensor.from_numpy(rng.standard_normal((B, w['dz']), dtype=np.float32))
    for _ in range(steps):
        f.train_step(x, e, model.optimizer)
    c.synchronize()
    print('mode %d path %s loss %.6f' % (mode, f.path(B), float(f.scalars.numpy()[0])))


if __name__ == '__main__':
    main()
