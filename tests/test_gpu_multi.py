"""Multi-GPU test of the fused NVLink peer-memory allreduce + Adam kernel (csrc/peer.cu).  Needs >= 2 GPUs on the box
(skipped otherwise): two processes, one per GPU, exchange CUDA IPC handles over torch.distributed and run several
training-style steps; the result must equal NumPy's rank-ordered sum + Adam and be bit-identical on both ranks."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import vaemolsim_b200 as v
from vaemolsim_b200 import parallel
from oracle import vae as ovae

grp = parallel.Group()
rank, world = grp.rank, grp.world
n = 44396
pe = parallel.PeerExchange(grp, n)

class F(object):
    pass

f = F()
rng0 = np.random.default_rng(0)
theta0 = rng0.standard_normal(n).astype(np.float32)
f.theta, f.m, f.v, f.t = v.Tensor.from_numpy(theta0.copy()), v.Tensor.zeros((n, )), v.Tensor.zeros((n, )), 0
opt = v.models.Adam(learning_rate=1e-3)
c = v._abi.ctx()
want, m, vv = theta0.astype(np.float32).copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
gout = v.Tensor((n, ))
for step in range(1, 6):
    grads = [np.random.default_rng([step, r]).standard_normal(n).astype(np.float32) for r in range(world)]
    slot = pe.next_slot()
    c.lib.vms_memcpy_h2d(slot, grads[rank].ctypes.data, 4 * n, c.stream)   # "the rank's gradient of this step"
    pe.allreduce_adam(f, opt, grad_out=gout)
    g = np.zeros(n, np.float32)
    for r in range(world):
        g = g + grads[r]                      # rank order, float32: what the kernel does
    g = g * np.float32(1.0 / world)
    assert np.array_equal(gout.numpy(), g), 'step %%d: reduced gradient differs' %% step
    lr_t = np.float32(1e-3 * np.sqrt(1.0 - 0.999 ** step) / (1.0 - 0.9 ** step))
    m = m + (g - m) * np.float32(1.0 - 0.9)
    vv = vv + (g * g - vv) * np.float32(1.0 - 0.999)
    want = want - lr_t * m / (np.sqrt(vv) + np.float32(1e-7))
got = f.theta.numpy()
err = float(np.abs(got - want).max())
import hashlib
print('RESULT %%d %%.3e %%s' %% (rank, err, hashlib.sha1(got.tobytes()).hexdigest()), flush=True)
pe.close()
grp.close()
'''


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_peer_allreduce_adam_two_gpus(tmp_path):
    import ctypes
    from vaemolsim_b200 import _abi
    n = ctypes.c_int(0)
    _abi.load().vms_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=280)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    res = [[ln for ln in o.splitlines() if ln.startswith('RESULT')][0].split() for o in outs]
    assert all(float(r[2]) < 2e-6 for r in res), res     # Adam on the summed gradient (float32 rounding only)
    assert res[0][3] == res[1][3], 'replicas diverged'    # bit-identical parameters on both ranks
