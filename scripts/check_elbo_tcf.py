"""Device cross-check and timing of the EXPERIMENTAL whole-step tensor-core kernel (csrc/elbo_tcf.cu, plan mode 3) against
the float32 FFMA per-layer plan (mode 1) and the production fused kernel (mode 0) at the named batch.
    python scripts/check_elbo_tcf.py [batch ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vaemolsim_b200 as v  # noqa: E402


def main():
    batches = [int(a) for a in sys.argv[1:]] or [4096, 777]
    w = bench.WORKLOADS['c2']
    c = v._abi.ctx()
    for B in batches:
        model = bench.build_model(v, w, B)
        f = model.fused(B)
        f.set_tc_auto_batch(1 << 40)
        rng = np.random.default_rng(11)
        x = v.Tensor.from_numpy(rng.standard_normal((B, w['dx']), dtype=np.float32))
        e = v.Tensor.from_numpy(rng.standard_normal((B, w['dz']), dtype=np.float32))
        res = {}
        for mode in (1, 3, 4):
            f.set_mode(mode)
            f.forward_backward(x, e)
            c.synchronize()
            res[mode] = (f.grad.numpy().copy(), f.scalars.numpy()[:3].copy(), f.path(B))
        print('B = %d   paths: %s   tensor-core wait time-out: %s' % (B, [res[m][2] for m in (1, 3, 4)], f.tc_status()))
        ga, sa = res[3][:2]
        for other in (1, 4):
            gb, sb = res[other][:2]
            print('  mode 3 vs mode %d: scalars %s vs %s | grad max-rel %.2e  norm-rel %.2e' % (
                other, sa, sb, float(np.abs(ga - gb).max() / np.abs(gb).max()),
                float(np.linalg.norm(ga - gb) / np.linalg.norm(gb))))
        gb = res[1][0]
        names = [('enc0W', 1200), ('enc0b', 200), ('enc1W', 800), ('enc1b', 4), ('dec0W', 400), ('dec0b', 200), ('dec1W', 2400),
                 ('dec1b', 12)]
        for blk in range(4):
            names += [('b%d.d1W' % blk, 100), ('b%d.d1b' % blk, 100), ('b%d.hW' % blk, 9500), ('b%d.hb' % blk, 95)]
        o, line = 0, []
        for name, n in names:
            line.append('%s %.1e' % (name, float(np.linalg.norm(ga[o:o + n] - gb[o:o + n]) / (np.linalg.norm(gb[o:o + n]) + 1e-30))))
            o += n
        print('  per layer (norm-rel vs mode 1): ' + '  '.join(line))
        ev = bench.Events(c, 1)
        for mode in (3, 4):
            f.set_mode(mode)
            for _ in range(3):
                f.forward_backward(x, e)
            c.synchronize()
            ev.record(0)
            for _ in range(20):
                f.forward_backward(x, e)
            ev.record(1)
            c.synchronize()
            print('  mode %d (%s): %.4f ms / call  forward_backward (always re-packs the weight images)' % (
                mode, f.path(B), ev.elapsed_ms(0, 1) / 20))
        # training steps: the finish kernel writes the next step's weight images (no pre-pack launch in steady state);
        # parameters after 5 steps against the FFMA fused kernel's
        theta0 = f.theta.numpy().copy()
        thetas = {}
        for mode in (3, 4):
            f.set_mode(mode)
            c.lib.vms_memcpy_h2d(f.theta.ptr, theta0.ctypes.data, theta0.nbytes, c.stream)
            c.lib.vms_memset(f.m.ptr, 0, f.m.nbytes, c.stream)
            c.lib.vms_memset(f.v.ptr, 0, f.v.nbytes, c.stream)
            c.synchronize()
            f.invalidate()
            f.t = 0
            for _ in range(5):
                f.train_step(x, e, model.optimizer)
            c.synchronize()
            thetas[mode] = f.theta.numpy().copy()
            ev.record(0)
            l0 = v._abi.launch_count()
            for _ in range(50):
                f.train_step(x, e, model.optimizer)
            ev.record(1)
            c.synchronize()
            print('  mode %d (%s): %.4f ms / train_step, %.1f launches per step' % (
                mode, f.path(B), ev.elapsed_ms(0, 1) / 50, (v._abi.launch_count() - l0) / 50.0))
        d = thetas[3] - thetas[4]
        print('  parameters after 5 train steps, mode 3 vs mode 4: max abs diff %.2e (update size %.2e)' % (
            float(np.abs(d).max()), float(np.abs(thetas[4] - theta0).max())))
        c.lib.vms_memcpy_h2d(f.theta.ptr, theta0.ctypes.data, theta0.nbytes, c.stream)
        c.synchronize()
        f.invalidate()
        del f, model


if __name__ == '__main__':
    main()
