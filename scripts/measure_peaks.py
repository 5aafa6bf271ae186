"""Measures the compute peaks the roofline fractions are quoted against with the library's own probe
kernels (csrc/probe.cu) and writes them to profiles/<tag>_measured_peaks.json.    python scripts/measure_peaks.py [tag]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vaemolsim_b200 as v  # noqa: E402


def measure():
    c = v._abi.ctx()
    t, ms = C.c_double(0), C.c_double(0)
    out = {'device_sms': c.sm_count, 'how': 'csrc/probe.cu: best of 5 launches, CUDA events on the launching stream'}
    c.lib.vms_probe_ffma(4096, 5, C.byref(t), C.byref(ms), c.stream)
    out['fp32_ffma_tflops'] = t.value
    out['fp32_ffma_probe_ms'] = ms.value
    out['fp32_ffma_nominal_tflops'] = c.sm_count * 128 * 2 * 1.965e9 / 1e12
    for kind, name in ((0, 'bf16'), (1, 'tf32')):
        for M, N in ((128, 256), (128, 96), (64, 96)):
            c.lib.vms_probe_mma(kind, M, N, 20000, 5, C.byref(t), C.byref(ms), c.stream)
            out['tcgen05_%s_m%d_n%d_tflops' % (name, M, N)] = t.value
    out['fp32_equiv_3xbf16_m128_n256_tflops'] = out['tcgen05_bf16_m128_n256_tflops'] / 6.0
    out['fp32_equiv_3xtf32_m128_n256_tflops'] = out['tcgen05_tf32_m128_n256_tflops'] / 3.0
    return out


if __name__ == '__main__':
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    res = measure()
    path = os.path.join(ROOT, 'gpurun_out', '%s_measured_peaks.json' % tag)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(res, open(path, 'w'), indent=1)
    print(json.dumps(res, indent=1))
