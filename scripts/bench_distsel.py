"""C3 neighbour-selection timing (device-resident, CUDA events, L2 flushed) for the streaming kernel and the one-CTA-per-row
kernel, with / without particle_info and indices.   python scripts/bench_distsel.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402


def main():
    c = v._abi.ctx()
    rng = np.random.default_rng(3001)
    B, N, L = 4096, 10000, np.float32(46.416)
    frame = rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32)
    coords = v.Tensor.from_numpy(np.ascontiguousarray(np.broadcast_to(frame, (B, N, 3))))
    info = v.Tensor.from_numpy(np.ascontiguousarray(np.broadcast_to(np.eye(2, dtype=np.float32)[rng.integers(0, 2, N)], (B, N, 2))))
    ref = v.Tensor.from_numpy(np.random.default_rng(3002).uniform(-L / 2, L / 2, (B, 1, 3)).astype(np.float32))
    box = np.array([L, L, L], np.float32)
    flush = v.Tensor((64 << 20, ))
    ev = bench.Events(c, 1)
    for k in (50, 10):
        layer = v.mappings.DistanceSelection(3.0, max_included=k, box_lengths=box)
        for tag, kw in (('xyz only', {}), ('xyz + info + idx', dict(particle_info=info, return_indices=True))):
            for _ in range(2):
                layer(coords, ref, **kw)
            ts = []
            for _ in range(10):
                c.lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
                ev.record(0)
                layer(coords, ref, **kw)
                ev.record(1)
                c.synchronize()
                ts.append(ev.elapsed_ms(0, 1))
            ms = float(np.mean(ts[2:]))
            nbytes = B * (12 * N + 12 + k * 12 + (k * (4 * 2 + 4 * 2 + 4) if kw else 0))
            print('VMS_DISTSEL_STREAM=%s k=%d %-18s %.4f ms  %.0f GB/s  %.3f of 6544.7' % (
                os.environ.get('VMS_DISTSEL_STREAM', '0'), k, tag, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / 6544.7))


if __name__ == '__main__':
    main()
