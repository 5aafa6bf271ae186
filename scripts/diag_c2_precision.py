"""Precision diagnostic at the exact C2 shape (B = 4096, H = 200, FH = 100, K = 32): flat-gradient error of every plan of
vms_elbo_forward_backward against the float64 oracle, per layer, next to the error of the float32 ORACLE itself (what the
reference's own float32 arithmetic would give).    python scripts/diag_c2_precision.py [widen]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import vaemolsim_b200 as v  # noqa: E402
from helpers import flat_grad_from_oracle, vae_from_oracle  # noqa: E402
from oracle import vae as ovae  # noqa: E402


def main():
    widen = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    P = ovae.init_vae(2003, dx=6, dz=2, hidden=200, prior='realnvp', num_blocks=4, num_bins=32, flow_hidden=100)
    rng = np.random.default_rng(2004)
    if widen > 0:
        for blk in P['flow']:
            for k in ('w', 'h', 's'):
                blk[k] = ((blk[k][0] * widen).astype(np.float32), rng.normal(0, 0.5, blk[k][1].shape).astype(np.float32))
    B = 4096
    rng = np.random.default_rng(1001)
    x = rng.standard_normal((B, 6), dtype=np.float32)
    eps = rng.standard_normal((B, 2), dtype=np.float32)
    P64 = ovae.cast_params(P, np.float64)
    out64, G64 = ovae.elbo_backward(P64, x.astype(np.float64), eps.astype(np.float64))
    want = flat_grad_from_oracle(P64, G64).astype(np.float64)
    out32, G32 = ovae.elbo_backward(P, x, eps)
    g32 = flat_grad_from_oracle(P, G32).astype(np.float64)
    names = [('enc0W', 1200), ('enc0b', 200), ('enc1W', 800), ('enc1b', 4), ('dec0W', 400), ('dec0b', 200), ('dec1W', 2400),
             ('dec1b', 12)]
    for blk in range(4):
        names += [('b%d.d1W' % blk, 100), ('b%d.d1b' % blk, 100), ('b%d.hW' % blk, 9500), ('b%d.hb' % blk, 95)]

    def report(tag, g, scal):
        rel = np.linalg.norm(g - want) / np.linalg.norm(want)
        o, parts = 0, []
        for name, n in names:
            e = np.linalg.norm(g[o:o + n] - want[o:o + n]) / np.linalg.norm(want)
            if e > 1e-6:
                parts.append('%s %.1e' % (name, e))
            o += n
        print('%-28s grad norm-rel %.2e | loss rel %.1e nll %.1e kl %.1e | layers > 1e-6 of total: %s' % (
            tag, rel, abs(scal[0] - out64['loss']) / abs(out64['loss']), abs(scal[1] - out64['nll']) / abs(out64['nll']),
            abs(scal[2] - out64['kl']) / abs(out64['kl']), ' '.join(parts)))

    print('widen = %g, B = %d' % (widen, B))
    report('float32 oracle (NumPy)', g32, [out32['loss'], out32['nll'], out32['kl']])
    f = vae_from_oracle(v, P, max_batch=B).fused(B)
    f.set_tc_auto_batch(1 << 40)
    for mode in (4, 1, 2, 3):
        try:
            f.set_mode(mode)
        except Exception as e:
            print('mode %d unavailable: %s' % (mode, e))
            continue
        scal = f.forward_backward(v.as_tensor(x), v.as_tensor(eps)).numpy()
        report('mode %d (%s)' % (mode, f.path(B)), f.grad.numpy().astype(np.float64), scal)


if __name__ == '__main__':
    main()
