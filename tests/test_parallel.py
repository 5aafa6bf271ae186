"""CPU tests of the multi-GPU host logic: row sharding, rank-count-independent synthetic inputs, and the
single data-path collective (flat-gradient sum) over gloo with world_size 2.  The data-parallel step is emulated with the
NumPy oracle standing in for the per-rank ELBO kernel: what is under test is `vaemolsim_b200/parallel.py`."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vaemolsim_b200 import parallel  # noqa: E402


def test_shard_rows_partitions_exactly():
    for n in (0, 1, 7, 4096, 65536, 262144, 10007):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_global_row_seed_is_rank_count_independent():
    a = parallel.global_row_seed(1001, 4096).standard_normal(8)
    b = parallel.global_row_seed(1001, 4096).standard_normal(8)
    c = parallel.global_row_seed(1001, 0).standard_normal(8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
from vaemolsim_b200 import parallel
from oracle import vae as ovae

grp = parallel.Group(backend=%(backend)r)
rank, world = grp.rank, grp.world
assert world == 2
# global batch of 64 configurations, sharded by rows; every rank builds the same model
B = 64
rng = np.random.default_rng(7)
x = rng.standard_normal((B, 6)).astype(np.float64)
eps = rng.standard_normal((B, 2)).astype(np.float64)
P = ovae.cast_params(ovae.init_vae(3, prior='realnvp', hidden=16, flow_hidden=8, num_bins=8), np.float64)
lo, hi = parallel.shard_rows(B, rank, world)
out, G = ovae.elbo_backward(P, x[lo:hi], eps[lo:hi])
g = ovae.flatten(ovae.grad_list(P, G)).astype(np.float32)
parts = grp.all_gather_bytes(('rank%%d' %% rank).encode())
assert parts == [b'rank0', b'rank1'], parts
grp.allreduce_sum_numpy_(g)          # the one collective of the training step
g = g / world                          # loss is a batch MEAN over equal shards (losses.py:253)
full, Gf = ovae.elbo_backward(P, x, eps)
want = ovae.flatten(ovae.grad_list(P, Gf)).astype(np.float32)
err = float(np.linalg.norm(g - want) / np.linalg.norm(want))
mx = grp.max(float(rank + 1))
sm = grp.sum(float(rank + 1))
grp.barrier()
grp.close()
print('RESULT %%d %%.3e %%g %%g' %% (rank, err, mx, sm), flush=True)
'''


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
@pytest.mark.parametrize('backend', ['socket', 'gloo'])
def test_gradient_allreduce_world2(tmp_path, backend):
    """world_size 2 on the CPU: the product's own socket rendezvous (no torch) and the torch.distributed gloo backend."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT, 'backend': backend})
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), OMP_NUM_THREADS='1')
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=280)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    for o in outs:
        line = [ln for ln in o.splitlines() if ln.startswith('RESULT')][0].split()
        assert float(line[2]) < 1e-6, o       # sharded + allreduced gradient == full-batch gradient
        assert float(line[3]) == 2.0 and float(line[4]) == 3.0
