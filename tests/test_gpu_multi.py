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


DP_WORKER = r'''
import os, sys, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
import vaemolsim_b200 as v
from vaemolsim_b200 import parallel
import vaemolsim_b200._protocols as PR

grp = parallel.Group()
rank, world = grp.rank, grp.world
d = v.dists


def make():
    v.set_seed(3)
    kinds = [d.Normal] * 2 + [d.VonMises] * 2
    dist = d.AutoregressiveBlockwise(4, kinds, conditional=True, conditional_event_shape=3,
                                     auto_net_params={'hidden_units': [16], 'activation': 'tanh'})
    m = v.models.MappingToDistribution(dist, mapping=v.mappings.FCDeepNN(dist.params_size(), hidden_dim=24, activation='tanh'),
                                       name='decoder')
    m.compile(optimizer=v.models.Adam(2e-3), loss=v.losses.LogProbLoss())
    return m


rng = np.random.default_rng(3)
B = 128
z = rng.normal(size=(B, 3)).astype(np.float32)
x = np.concatenate([rng.normal(size=(B, 2)), rng.uniform(-3, 3, (B, 2))], axis=1).astype(np.float32)
lo, hi = parallel.shard_rows(B, rank, world)
model = make()
model(z[:2])
model.distribute(grp)
for step in range(6):
    model.train_on_batch(z[lo:hi], x[lo:hi])
got = np.concatenate([w.numpy().ravel() for w in model.weights])
# the same six steps on the whole batch in one process
ref = make()
ref(z[:2])
for step in range(6):
    ref.train_on_batch(z, x)
want = np.concatenate([w.numpy().ravel() for w in ref.weights])
moved = float(np.abs(want - np.concatenate([w.numpy().ravel() for w in (lambda m: (m(z[:2]), m)[1])(make()).weights])).max())
err = float(np.abs(got - want).max())
print('RESULT %%d %%.3e %%.3e %%s' %% (rank, err, moved, hashlib.sha1(got.tobytes()).hexdigest()), flush=True)
grp.close()
'''


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_data_parallel_training_of_a_tape_model_two_gpus(tmp_path):
    """`Model.distribute(group)`: the generic (tape) training path sharded over two GPUs -- one NCCL allreduce of the flat
    gradient per step -- against the same steps on the whole batch in one process; replicas bit-identical."""
    import ctypes
    from vaemolsim_b200 import _abi
    n = ctypes.c_int(0)
    _abi.load().vms_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'dp_worker.py'
    script.write_text(DP_WORKER % {'root': ROOT})
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
    assert all(float(r[3]) > 1e-3 for r in res), res          # the six steps moved the weights
    assert all(float(r[2]) < 2e-5 for r in res), res          # sharded == whole batch up to float32 summation order
    assert res[0][4] == res[1][4], 'replicas diverged'


ELBO_DP_WORKER = r'''
import os, sys, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import vaemolsim_b200 as v
from vaemolsim_b200 import parallel
from oracle import vae as ovae
from helpers import vae_from_oracle

grp = parallel.Group()
rank, world = grp.rank, grp.world
P = ovae.init_vae(5, dx=6, dz=2, hidden=200, prior='realnvp', num_blocks=4, num_bins=32, flow_hidden=100)
B = 4096
rng = np.random.default_rng(17)
x = rng.standard_normal((world * B, 6), dtype=np.float32)
eps = rng.standard_normal((world * B, 2), dtype=np.float32)
lo, hi = rank * B, (rank + 1) * B
xd, ed = v.Tensor.from_numpy(x[lo:hi]), v.Tensor.from_numpy(eps[lo:hi])
opt = v.models.Adam(1e-3)
res = {}
for name in ('fused', 'separate'):
    model = vae_from_oracle(v, P, max_batch=B, weight=1.0)
    f = model.fused(B)
    assert f.path(B) == 'tensor-core-fused'
    pe = parallel.PeerExchange(grp, f.n_params)
    n0 = v._abi.launch_count()
    for step in range(5):
        if name == 'fused':
            pe.train_step(f, xd, ed, B, opt)            # tile kernel + finish / exchange / Adam kernel
        else:
            f.forward_backward(xd, ed, grad_ptr=pe.next_slot())
            pe.allreduce_adam(f, opt)                   # the separate exchange kernel
    v.synchronize()
    res[name] = (f.theta.numpy().copy(), f.scalars.numpy().copy(), v._abi.launch_count() - n0)
    assert not pe.timed_out()
    pe.close()
same = bool(np.array_equal(res['fused'][0], res['separate'][0]))
print('RESULT %%d %%d %%d %%d %%s' %% (rank, int(same), res['fused'][2], res['separate'][2],
                                     hashlib.sha1(res['fused'][0].tobytes()).hexdigest()), flush=True)
grp.close()
'''


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_elbo_data_parallel_step_with_exchange_in_the_finish_kernel_two_gpus(tmp_path):
    """`vms_elbo_train_step_peer` at the C2 shape on two GPUs: the finish kernel of the whole-step tensor-core plan performs
    the NVLink exchange + Adam + next weight images itself (two launches per step after the first) -- bit-identical parameters
    to forward_backward + the separate `vms_peer_allreduce_adam` kernel, replicas bit-identical."""
    import ctypes
    from vaemolsim_b200 import _abi
    n = ctypes.c_int(0)
    _abi.load().vms_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'elbo_dp_worker.py'
    script.write_text(ELBO_DP_WORKER % {'root': ROOT})
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
    assert all(r[2] == '1' for r in res), res                 # fused finish + exchange == separate kernels, bit for bit
    assert all(int(r[3]) == 11 for r in res), res             # 5 steps x 2 launches + the first step's pre-pack
    assert all(int(r[4]) == 20 for r in res), res             # separate path: pre-pack, tile, finish, exchange per step
    assert res[0][5] == res[1][5], 'replicas diverged'


DP_BN_WORKER = r'''
import os, sys, hashlib
import numpy as np
sys.path.insert(0, %(root)r)
import vaemolsim_b200 as v
from vaemolsim_b200 import parallel
import vaemolsim_b200._protocols as PR

grp = parallel.Group()
rank, world = grp.rank, grp.world
d = v.dists


def make():
    v.set_seed(4)
    flow = v.flows.RQSSplineRealNVP(num_blocks=2, batch_norm=True, rqs_params={'hidden_dim': 16, 'num_bins': 8})
    dist = d.FlowedDistribution(flow, d.IndependentBlockwise(4, [d.Normal] * 4))
    dist.flow(np.ones((1, 4), np.float32))
    m = v.models.MappingToDistribution(dist, mapping=v.mappings.FCDeepNN(dist.params_size(), hidden_dim=24, activation='tanh',
                                                                         batch_norm=True), name='decoder')
    m.compile(optimizer=v.models.Adam(2e-3), loss=v.losses.LogProbLoss())
    return m


rng = np.random.default_rng(3)
B = 128
z = rng.normal(size=(B, 3)).astype(np.float32) * 2 + 1
x = rng.normal(size=(B, 4)).astype(np.float32)
lo, hi = parallel.shard_rows(B, rank, world)
model = make()
model(z[:2])
w0 = np.concatenate([w.numpy().ravel() for w in model.weights])
model.distribute(grp)
for step in range(5):
    model.train_on_batch(z[lo:hi], x[lo:hi])
got = np.concatenate([w.numpy().ravel() for w in model.weights])
ref = make()
ref(z[:2])
for step in range(5):
    ref.train_on_batch(z, x)
want = np.concatenate([w.numpy().ravel() for w in ref.weights])
err = float(np.abs(got - want).max())
moved = float(np.abs(want - w0).max())
print('RESULT %%d %%.3e %%.3e %%s' %% (rank, err, moved, hashlib.sha1(got.tobytes()).hexdigest()), flush=True)
grp.close()
'''


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_cross_replica_batch_norm_statistics_two_gpus(tmp_path):
    """Batch normalisation under data-parallel training: the Keras layer inside FCDeepNN and the tfp bijector
    between flow blocks normalise with the moments of the WHOLE batch (two small allreduces forward, one in the reverse mode),
    so five sharded steps on two GPUs equal five steps on the whole batch in one process -- weights AND moving statistics --
    and the replicas stay bit-identical."""
    import ctypes
    from vaemolsim_b200 import _abi
    n = ctypes.c_int(0)
    _abi.load().vms_device_count(ctypes.byref(n))
    if n.value < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'dp_bn_worker.py'
    script.write_text(DP_BN_WORKER % {'root': ROOT})
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
    assert all(float(r[3]) > 1e-3 for r in res), res          # the steps moved the weights
    assert all(float(r[2]) < 5e-5 for r in res), res          # sharded == whole batch up to float32 summation order
    assert res[0][4] == res[1][4], 'replicas diverged'
