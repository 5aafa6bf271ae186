"""One C3 neighbour-selection call of each flavour (for ncu).  python scripts/prof_distsel.py [rows]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vaemolsim_b200 as v  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    c = v._abi.ctx()
    rng = np.random.default_rng(3001)
    N, L = 10000, np.float32(46.416)
    frame = rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32)
    coords = v.Tensor.from_numpy(np.ascontiguousarray(np.broadcast_to(frame, (B, N, 3))))
    info = v.Tensor.from_numpy(np.ascontiguousarray(np.broadcast_to(np.eye(2, dtype=np.float32)[rng.integers(0, 2, N)], (B, N, 2))))
    ref = v.Tensor.from_numpy(np.random.default_rng(3002).uniform(-L / 2, L / 2, (B, 1, 3)).astype(np.float32))
    box = np.array([L, L, L], np.float32)
    layer = v.mappings.DistanceSelection(3.0, max_included=50, box_lengths=box)
    for _ in range(2):
        layer(coords, ref)
        layer(coords, ref, particle_info=info, return_indices=True)
    c.synchronize()
    print('ok')


if __name__ == '__main__':
    main()
