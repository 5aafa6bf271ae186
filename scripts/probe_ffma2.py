import ctypes as C, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vaemolsim_b200 as v
c = v._abi.ctx()
for name in ('vms_probe_ffma', 'vms_probe_ffma2'):
    t, ms = C.c_double(), C.c_double()
    getattr(c.lib, name)(2000, 3, C.byref(t), C.byref(ms), c.stream)
    print(name, '%.1f TFLOP/s  %.3f ms' % (t.value, ms.value))
