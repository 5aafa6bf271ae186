"""GPU check + timing of the tcgen05 (3 x TF32) Dense forward (csrc/gemm_tc.cu) against NumPy float64 and against the FFMA
kernel (VMS_DENSE_TC=0).  python scripts/check_gemm_tc.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v


def main():

    c = v._abi.ctx()
    lib = c.lib
    rng = np.random.default_rng(0)
    T = lambda a: v.Tensor.from_numpy(np.ascontiguousarray(a, np.float32))


    def ev():
        e = C.c_void_p()
        lib.vms_event_create(C.byref(e))
        return e.value


    E0, E1 = ev(), ev()
    flush = v.Tensor((64 << 20, ))
    print('VMS_DENSE_TC =', os.environ.get('VMS_DENSE_TC', '(unset: tensor-core path on)'))
    for (B, K, N, act) in ((4096, 100, 95, 0), (5000, 100, 95, 2), (262144, 100, 95, 0), (65536, 64, 128, 1), (262144, 32, 16, 0)):
        x = rng.standard_normal((B, K)).astype(np.float32)
        W = (rng.standard_normal((K, N)) * 0.3).astype(np.float32)
        b = rng.standard_normal(N).astype(np.float32)
        dx, dW, db, out = T(x), T(W), T(b), v.Tensor((B, N))
        lib.vms_memset(out.ptr, 0xff, out.nbytes, c.stream)
        fn = lambda: lib.vms_dense_forward(dx.ptr, K, dW.ptr, db.ptr, B, K, N, act, None, 0, None, 0, out.ptr, N, c.stream)
        fn()
        c.synchronize()
        got = out.numpy()
        n_chk = min(B, 20000)
        idx = np.concatenate([np.arange(min(B, 300)), rng.integers(0, B, n_chk - min(B, 300)), np.arange(B - 300, B)])
        want = x[idx].astype(np.float64) @ W.astype(np.float64) + b.astype(np.float64)
        if act == 1:
            want = np.maximum(want, 0)
        elif act == 2:
            want = np.tanh(want)
        scale = np.sqrt((x[idx].astype(np.float64)**2) @ (W.astype(np.float64)**2)) + 1e-30   # conditioning of each dot product
        err = np.abs(got[idx] - want)
        ts = []
        for _ in range(8):
            lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
            lib.vms_event_record(E0, c.stream)
            fn()
            lib.vms_event_record(E1, c.stream)
            c.synchronize()
            ms = C.c_float()
            lib.vms_event_elapsed_ms(E0, E1, C.byref(ms))
            ts.append(ms.value * 1e3)
        us = float(np.median(ts[2:]))
        nbytes = 4.0 * (B * K + B * N + K * N)
        print('B=%7d K=%3d N=%3d act=%d  max|err| %.2e  max err/|x||w| %.2e  finite %s  %8.1f us  %6.1f GB/s  %6.2f TFLOP/s' %
              (B, K, N, act, err.max(), (err / scale).max(), bool(np.isfinite(got).all()), us, nbytes / us / 1e3,
               2.0 * B * K * N / us / 1e6))


if __name__ == '__main__':
    main()
