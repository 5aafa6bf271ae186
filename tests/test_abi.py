"""CPU: the C-ABI library loads and exports every symbol include/vms_b200.h declares (no compute calls)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'vms_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(vms_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_and_library_exports_match():
    from vaemolsim_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        from vaemolsim_b200 import build
        build.build()
    lib = _abi.load()
    syms = header_symbols()
    assert len(syms) >= 50
    assert sorted(_abi.EXPORTS) == syms, set(syms) ^ set(_abi.EXPORTS)
    out = subprocess.run(['nm', '-D', '--defined-only', _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (vms_[a-z0-9_]+)', out))
    assert set(syms) <= exported, set(syms) - exported
    assert lib.vms_abi_version() == 1
    assert lib.vms_launch_count() == 0


def test_struct_layouts_match_header():
    """ctypes mirrors of the by-pointer structs have the sizes a C compiler gives the header's definitions."""
    import ctypes
    from vaemolsim_b200 import _abi
    code = '#include <stdio.h>\n#include "vms_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(vms_rqs_args), ' \
           'sizeof(vms_rqs_bwd_args), sizeof(vms_elbo_desc), sizeof(vms_gaa_weights));return 0;}'
    exe = '/tmp/vms_sizes'
    subprocess.run(['gcc', '-x', 'c', '-', '-I', os.path.join(ROOT, 'include'), '-o', exe], input=code, text=True, check=True)
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_abi.RqsArgs), ctypes.sizeof(_abi.RqsBwdArgs), ctypes.sizeof(_abi.ElboDesc),
                     ctypes.sizeof(_abi.GaaWeights)]


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    from vaemolsim_b200 import _abi
    import ctypes
    lib = _abi.load()
    n = ctypes.c_int(0)
    try:
        lib.vms_device_count(ctypes.byref(n))
        have = n.value > 0
    except RuntimeError:
        have = False
    if have:
        pytest.skip('a CUDA device is present')
    import vaemolsim_b200 as v
    with pytest.raises(RuntimeError, match='CUDA device'):
        v.Tensor.zeros((4, ))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'vaemolsim_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), fn
