"""CPU tests of host-side logic that needs no device: the PCG64 jump-ahead description handed to `vms_mc_run_pcg64`, the
DLPack capsule plumbing, struct sizes of the new ABI structs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MULT = 0x2360ED051FC65DA44385DF649FCCF645
M128 = (1 << 128) - 1


def _pcg_double(state):
    hi, lo = state >> 64, state & ((1 << 64) - 1)
    v, rot = hi ^ lo, state >> 122
    out = ((v >> rot) | (v << ((64 - rot) & 63))) & ((1 << 64) - 1)
    return (out >> 11) * (1.0 / 9007199254740992.0)


@pytest.mark.parametrize('seed,chain0,n_global,B,n_steps', [(4002, 0, 64, 64, 5), (7, 1000, 65536, 37, 4), (None, 5, 6, 1, 3)])
def test_pcg_stream_descriptor_reproduces_numpy_columns(seed, chain0, n_global, B, n_steps):
    """What the device kernel does with the descriptor (per-chain advance by chain0 + c + 1, then the affine stride per MC
    step, XSL-RR output, >> 11 * 2^-53), restated with Python integers, equals NumPy's `random(size=(n_steps, n_global))`
    columns [chain0, chain0 + B) -- mcmc.py:119."""
    from vaemolsim_b200 import mcmc
    mc = mcmc.MCMC(None, None, random_seed=seed, stream_layout=(chain0, n_global))
    twin = np.random.default_rng(mc._rng.bit_generator.seed_seq) if seed is None else np.random.default_rng(seed)
    mc._rng.random(size=3)  # not at the start of the stream
    twin.random(size=3)
    st = mc._pcg_stream(chain0, n_global)
    state = (st.state_hi << 64) | st.state_lo
    inc = (st.inc_hi << 64) | st.inc_lo
    jm = (st.stride_mul_hi << 64) | st.stride_mul_lo
    ja = (st.stride_add_hi << 64) | st.stride_add_lo
    assert st.chain0 == chain0
    want = twin.random(size=(n_steps, n_global))[:, chain0:chain0 + B]
    got = np.empty((n_steps, B))
    for c in range(B):
        s = state
        for _ in range(chain0 + c + 1):
            s = (s * MULT + inc) & M128
        for k in range(n_steps):
            got[k, c] = _pcg_double(s)
            s = (jm * s + ja) & M128
    assert np.array_equal(got, want)
    # and the generator is advanced the way run_fused does it afterwards
    mc._rng.bit_generator.advance(n_steps * n_global)
    assert mc._rng.random() == twin.random()


def test_new_struct_sizes_match_header():
    from vaemolsim_b200 import _abi
    code = '#include <stdio.h>\n#include "vms_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(vms_pcg64_stream), ' \
           'sizeof(vms_mc_desc));return 0;}'
    exe = '/tmp/vms_sizes2'
    subprocess.run(['gcc', '-x', 'c', '-', '-I', os.path.join(ROOT, 'include'), '-o', exe], input=code, text=True, check=True)
    sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(_abi.Pcg64Stream), C.sizeof(_abi.McDesc)]


def test_dlpack_import_rejects_host_memory_without_touching_a_device():
    """The DLPack bridge wraps CUDA memory only: a NumPy (kDLCPU) producer is refused before any device call, and the
    capsule is left unconsumed (there is no host compute path to fall back to)."""
    from vaemolsim_b200 import _abi
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    with pytest.raises(ValueError, match='CUDA device memory'):
        _abi.Tensor.from_dlpack(a)
