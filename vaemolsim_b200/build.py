"""Build libvms_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:

    python -m vaemolsim_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind a C ABI (include/vms_b200.h): no torch, no pybind.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvms_b200.so')
SOURCES = ['runtime.cu', 'rqs.cu', 'rqs_stream.cu', 'dense.cu', 'gemm_tc.cu', 'flow_tc.cu', 'mlp_stream.cu', 'logprob.cu', 'reduce.cu', 'batchnorm.cu', 'distsel.cu', 'mcmc.cu', 'adam.cu', 'elbo.cu', 'elbo_fused.cu', 'elbo_tcf.cu', 'mc_fused.cu', 'mc_chain.cu', 'mc_nb.cu', 'peer.cu', 'probe.cu', 'autodiff.cu', 'gaa.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr'
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'vms_b200.h'))
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace('.cu', '.o'))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('--- nvcc %s\n%s\n' % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
