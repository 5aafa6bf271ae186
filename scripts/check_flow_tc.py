"""Device cross-check of the tensor-core coupling-block kernels (csrc/flow_tc.cu, plan mode 2) against the float32
FFMA per-layer plan (mode 1) and the single fused kernel (mode 0): forward outputs per row, loss scalars, flat
gradient; then timings at the C5 shard size.  Run on a B200:  python scripts/check_flow_tc.py [batch ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (model builder)
import vaemolsim_b200 as v  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def main():
    batches = [int(a) for a in sys.argv[1:]] or [10007, 65536 + 37]
    w = bench.WORKLOADS['c2']
    c = v._abi.ctx()
    if os.environ.get('FLOW_TC_PROFILE'):  # under ncu: two mode-2 steps at the first batch, nothing else
        B = batches[0]
        model = bench.build_model(v, w, B)
        f = model.fused(B)
        rng = np.random.default_rng(11)
        x = v.Tensor.from_numpy(rng.standard_normal((B, w['dx']), dtype=np.float32))
        e = v.Tensor.from_numpy(rng.standard_normal((B, w['dz']), dtype=np.float32))
        f.set_mode(2)
        for _ in range(2):
            f.forward_backward(x, e)
        c.synchronize()
        return
    for B in batches:
        model = bench.build_model(v, w, B)
        f = model.fused(B)
        rng = np.random.default_rng(11)
        x = v.Tensor.from_numpy(rng.standard_normal((B, w['dx']), dtype=np.float32))
        e = v.Tensor.from_numpy(rng.standard_normal((B, w['dz']), dtype=np.float32))
        res = {}
        for mode in (1, 2, 0):
            f.set_mode(mode)
            out = f.forward(x, e)
            fw = {k: out[k].numpy() for k in ('z', 'logq', 'logpz', 'logpx', 'scalars')}
            f.forward_backward(x, e)
            c.synchronize()
            res[mode] = (fw, f.grad.numpy().copy(), f.scalars.numpy().copy())
        print('B = %d   tensor-core wait time-out: %s' % (B, f.tc_status()))
        for other in (1, 0):
            fa, ga, sa = res[2]
            fb, gb, sb = res[other]
            print('  mode 2 vs mode %d: ' % other + '  '.join('%s %.2e' % (k, rel(fa[k], fb[k])) for k in fa) +
                  '  | grad max-rel %.2e  norm-rel %.2e | scalars(bwd) %.2e' %
                  (rel(ga, gb), float(np.linalg.norm(ga - gb) / np.linalg.norm(gb)), rel(sa[:3], sb[:3])))
        # per-layer gradient error of the flow blocks (offsets: enc/dec = 5216 parameters, then 9795 per block)
        ga, gb = res[2][1], res[1][1]
        o = 5216
        for blk in range(4):
            seg = [('d1W', 100), ('d1b', 100), ('hW', 9500), ('hb', 95)]
            s = []
            for name, n in seg:
                s.append('%s %.1e' % (name, float(np.linalg.norm(ga[o:o + n] - gb[o:o + n]) /
                                                  (np.linalg.norm(gb[o:o + n]) + 1e-30))))
                o += n
            print('    block %d: %s' % (blk, '  '.join(s)))
        o = 5216
        print('    |g| mode 2 / mode 1, block 0: d1W %.3e / %.3e   d1b %.3e / %.3e   hW %.3e / %.3e   hb %.3e / %.3e' % (
            np.linalg.norm(ga[o:o + 100]), np.linalg.norm(gb[o:o + 100]), np.linalg.norm(ga[o + 100:o + 200]),
            np.linalg.norm(gb[o + 100:o + 200]), np.linalg.norm(ga[o + 200:o + 9700]), np.linalg.norm(gb[o + 200:o + 9700]),
            np.linalg.norm(ga[o + 9700:o + 9795]), np.linalg.norm(gb[o + 9700:o + 9795])))
        print('    hb mode 2:', ga[o + 9700:o + 9706], ' mode 1:', gb[o + 9700:o + 9706])
        print('    hW[0,:4] mode 2:', ga[o + 200:o + 204], ' mode 1:', gb[o + 200:o + 204])
        print('    enc/dec grad norm-rel %.2e' % (np.linalg.norm(ga[:5216] - gb[:5216]) / np.linalg.norm(gb[:5216])))
        # timings
        ev = bench.Events(c, 1)
        for mode in (2, 0):
            f.set_mode(mode)
            for _ in range(2):
                f.forward_backward(x, e)
            c.synchronize()
            ev.record(0)
            n = 5
            for _ in range(n):
                f.forward_backward(x, e)
            ev.record(1)
            c.synchronize()
            ms = ev.elapsed_ms(0, 1) / n
            print('  mode %d: %.3f ms / step (fwd+bwd)  %.1f M configs/s' % (mode, ms, B / ms / 1e3))
        f.set_mode(2)
        out = f.forward(x, e)
        c.synchronize()
        ev.record(0)
        for _ in range(5):
            f.forward(x, e)
        ev.record(1)
        c.synchronize()
        print('  mode 2 forward only: %.3f ms' % (ev.elapsed_ms(0, 1) / 5))
        del f, model


if __name__ == '__main__':
    main()
