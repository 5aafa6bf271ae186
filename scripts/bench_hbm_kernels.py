"""GPU microbenchmark of the HBM-bound kernels at streaming sizes (development aid; also the command profiled by ncu
for profiles/*rqs*).  python scripts/bench_hbm_kernels.py [n_elements] [reps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v

c = v._abi.ctx()
lib = c.lib
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
K = 32


def ev():
    e = C.c_void_p()
    lib.vms_event_create(C.byref(e))
    return e.value


E0, E1 = ev(), ev()
flush = v.Tensor((64 << 20, ))


def timeit(fn):
    ts = []
    for _ in range(reps):
        lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
        lib.vms_event_record(E0, c.stream)
        fn()
        lib.vms_event_record(E1, c.stream)
        c.synchronize()
        ms = C.c_float()
        lib.vms_event_elapsed_ms(E0, E1, C.byref(ms))
        ts.append(ms.value * 1e3)
    ts = np.array(ts[2:])
    return float(np.median(ts)), float(ts.min())


T = lambda a: v.Tensor.from_numpy(np.ascontiguousarray(a, np.float32))
rw, rh, rs = T(rng.normal(0, 0.5, (n, K))), T(rng.normal(0, 0.5, (n, K))), T(rng.normal(0, 0.5, (n, K - 1)))
x, g = T(rng.uniform(-10, 10, n)), T(rng.normal(size=n))
y, l = v.Tensor((n, )), v.Tensor((n, ))
gi, gw, gh, gs = v.Tensor((n, )), v.Tensor((n, K)), v.Tensor((n, K)), v.Tensor((n, K - 1))
peak = 6544.7
for nm, nbytes, fn in (
    ('rqs_forward', n * (4 * (3 * K - 1) + 12),
     lambda: lib.vms_rqs_forward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, y.ptr, l.ptr, c.stream)),
    ('rqs_inverse', n * (4 * (3 * K - 1) + 12),
     lambda: lib.vms_rqs_inverse(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, y.ptr, l.ptr, c.stream)),
    ('rqs_backward', n * (8 * (3 * K - 1) + 16),
     lambda: lib.vms_rqs_backward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, 1, g.ptr, g.ptr, gi.ptr, gw.ptr, gh.ptr,
                                  gs.ptr, c.stream)),
):
    med, mn = timeit(fn)
    print('%-12s n=%8d K=%d  median %8.2f us  min %8.2f us  %7.1f GB/s algorithmic = %.3f of %.1f GB/s' %
          (nm, n, K, med, mn, nbytes / med / 1e3, nbytes / med / 1e3 / peak, peak))

# decoder-distribution log_prob (K4) and neighbour selection (K6)
i32 = lambda a: (C.c_int32 * len(a))(*a)
nr, D = 1 << 22, 6
xx, pp, lp = T(rng.standard_normal((nr, D))), T(rng.standard_normal((nr, 2 * D))), v.Tensor((nr, ))
kind, loc, loc2, sc = i32([0] * D), i32(list(range(D))), i32([-1] * D), i32(list(range(D, 2 * D)))
med, mn = timeit(lambda: lib.vms_blockwise_log_prob(xx.ptr, D, pp.ptr, 2 * D, nr, D, kind, loc, loc2, sc, 1, lp.ptr, 0, c.stream))
nb = nr * (3 * D * 4 + 4)
print('normal_lp    rows=%8d D=%d  median %8.2f us  min %8.2f us  %7.1f GB/s algorithmic = %.3f of %.1f GB/s' %
      (nr, D, med, mn, nb / med / 1e3, nb / med / 1e3 / peak, peak))
for Bs, N, k in ((1024, 10000, 50), (4096, 10000, 50), (4096, 10000, 10)):
    coords = T(np.broadcast_to(rng.uniform(-23.2, 23.2, (1, N, 3)).astype(np.float32), (Bs, N, 3)))
    ref = T(rng.uniform(-23.2, 23.2, (Bs, 3)))
    box = T(np.full(3, 46.416, np.float32))
    oxyz = v.Tensor((Bs, k, 3))
    med, mn = timeit(lambda: lib.vms_dist_select(coords.ptr, None, Bs, N, ref.ptr, box.ptr, 0, 9.0, k, None, 0, oxyz.ptr,
                                                 None, None, c.stream))
    nb = Bs * (12 * N + 12 + k * 12)
    print('dist_select  rows=%5d N=%d k=%d  median %8.2f us  min %8.2f us  %7.1f GB/s algorithmic = %.3f of %.1f GB/s' %
          (Bs, N, k, med, mn, nb / med / 1e3, nb / med / 1e3 / peak, peak))
    del coords
