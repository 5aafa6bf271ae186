import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import vaemolsim_b200 as v
import test_gpu_autodiff as T

which = sys.argv[1] if len(sys.argv) > 1 else 'autoregressive'
v.set_seed(5)
rng = np.random.default_rng(3)
model = T._decoder(v, which)
model.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
z = rng.normal(size=(96, 3)).astype(np.float32)
x = np.concatenate([rng.normal(size=(96, 2)), rng.uniform(-3, 3, (96, 2))], axis=1).astype(np.float32)
model(z)
c = v._abi.ctx()
for w in model.weights:
    a = w.numpy(); m = getattr(w, '_grad_mask', None)
    a = a + rng.normal(0, 0.1, a.shape).astype(np.float32) * (m.numpy() if m is not None else 1.0)
    c.lib.vms_memcpy_h2d(w.ptr, np.ascontiguousarray(a).ctypes.data, a.nbytes, c.stream)
c.synchronize()
loss0, ws, grads = T._tape_gradients(v, model, z, x)
print('loss', loss0, 'n weights', len(ws))
h = 2e-3
for i, (w, g) in enumerate(zip(ws, grads)):
    b0 = w.numpy().copy()
    d = rng.normal(size=b0.shape); m = getattr(w, '_grad_mask', None)
    if m is not None: d = d * m.numpy()
    vals = []
    for sgn in (1, -1):
        a = np.ascontiguousarray((b0 + sgn * h * d).astype(np.float32))
        c.lib.vms_memcpy_h2d(w.ptr, a.ctypes.data, a.nbytes, c.stream); c.synchronize()
        vals.append(float(model._loss_tensor(v.as_tensor(z), v.as_tensor(x), False).numpy()))
    c.lib.vms_memcpy_h2d(w.ptr, np.ascontiguousarray(b0).ctypes.data, b0.nbytes, c.stream); c.synchronize()
    fd = (vals[0] - vals[1]) / (2 * h)
    print('%2d shape %-12s mask %-5s tape %+.5f  fd %+.5f' % (i, b0.shape, m is not None, float((g * d).sum()), fd))
