"""GPU probe: CUDA-event timings of individual C-ABI calls at the C2 workload shapes (development aid)."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v

c = v._abi.ctx()
lib = c.lib
rng = np.random.default_rng(0)


def ev():
    e = C.c_void_p(); lib.vms_event_create(C.byref(e)); return e.value


E0, E1 = ev(), ev()
flush = v.Tensor((64 << 20, ))


def timeit(fn, reps=30, cold=False):
    ts = []
    for _ in range(reps):
        if cold:
            lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
        lib.vms_event_record(E0, c.stream)
        fn()
        lib.vms_event_record(E1, c.stream)
        c.synchronize()
        ms = C.c_float(); lib.vms_event_elapsed_ms(E0, E1, C.byref(ms)); ts.append(ms.value * 1e3)
    ts = np.array(ts[5:])
    return float(np.median(ts)), float(ts.min())


def T(a):
    return v.Tensor.from_numpy(np.ascontiguousarray(a, np.float32))


B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
print('B =', B)
for (K, N, act, tag) in [(6, 200, 1, 'enc0'), (200, 4, 0, 'enc1'), (2, 200, 1, 'dec0'), (200, 12, 0, 'dec1'),
                         (1, 100, 2, 'flow d1'), (100, 95, 0, 'flow heads')]:
    x, W, b = T(rng.normal(size=(B, K))), T(rng.normal(size=(K, N))), T(rng.normal(size=N))
    out, g = v.Tensor((B, N)), T(rng.normal(size=(B, N)))
    gx, gW, gb = v.Tensor((B, K)), v.Tensor((K, N)), v.Tensor((N, ))
    ws = v.Tensor((max(1, lib.vms_dense_backward_workspace(B, K, N, 0) // 4), ))
    f = lambda: lib.vms_dense_forward(x.ptr, K, W.ptr, b.ptr, B, K, N, act, None, 0, None, 0, out.ptr, N, c.stream)
    bx = lambda: lib.vms_dense_backward(x.ptr, K, W.ptr, B, K, N, act, out.ptr, N, g.ptr, N, None, 0, None, 0, gx.ptr, K,
                                        0, None, None, None, 0, None, 0, ws.ptr, c.stream)
    bw = lambda: lib.vms_dense_backward(x.ptr, K, W.ptr, B, K, N, act, out.ptr, N, g.ptr, N, None, 0, None, 0, None, K,
                                        0, gW.ptr, gb.ptr, None, 0, None, 0, ws.ptr, c.stream)
    fl = 2.0 * B * K * N
    for nm, fn in (('fwd', f), ('gx', bx), ('gW+gb', bw)):
        med, mn = timeit(fn)
        print('%-10s %-6s [%d x %d x %d]  median %7.2f us  min %7.2f us  %6.2f TFLOP/s' % (tag, nm, B, K, N, med, mn,
                                                                                        fl / med / 1e6))
K = 32
for n in (B, 1 << 19, 1 << 21):
    rw, rh, rs = T(rng.normal(size=(n, K))), T(rng.normal(size=(n, K))), T(rng.normal(size=(n, K - 1)))
    x, g = T(rng.uniform(-10, 10, n)), T(rng.normal(size=n))
    y, l = v.Tensor((n, )), v.Tensor((n, ))
    gi, gw, gh, gs = v.Tensor((n, )), v.Tensor((n, K)), v.Tensor((n, K)), v.Tensor((n, K - 1))
    for nm, nbytes, fn in (
        ('rqs_forward', n * (4 * (3 * K - 1) + 12), lambda: lib.vms_rqs_forward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, y.ptr, l.ptr, c.stream)),
        ('rqs_inverse', n * (4 * (3 * K - 1) + 12), lambda: lib.vms_rqs_inverse(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, y.ptr, l.ptr, c.stream)),
        ('rqs_backward', n * (8 * (3 * K - 1) + 16), lambda: lib.vms_rqs_backward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, 1, g.ptr, g.ptr, gi.ptr, gw.ptr, gh.ptr, gs.ptr, c.stream)),
    ):
        med, mn = timeit(fn, cold=True)
        print('%-12s n=%8d  median %8.2f us  min %8.2f us  %7.1f GB/s (algorithmic)' % (nm, n, med, mn, nbytes / med / 1e3))
