#!/usr/bin/env python
"""Benchmark of the vaemolsim hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c1]

Workload (BASELINE.json configs[1], SURVEY 8d "C2"): flow-prior VAE -- encoder FCDeepNN 6->200->4 + IndependentNormal,
prior = RQSSplineRealNVP(4 blocks, 32 bins, hidden 100) over N(0, I_2), decoder FCDeepNN 2->200->12 + IndependentNormal,
KLDivergenceEstimate + LogProbLoss -- batch 4096 configurations PER GPU (weak scaling), synthetic N(0,1) inputs,
random-init weights (44,396 parameters).  One step = ELBO forward + backward (+ the single gradient allreduce when
N > 1) + Adam update.  Metric: configs/sec, whole job.

Lines printed (one JSON object, rank 0):
  value        inputs resident in HBM; per-step CUDA events on the launching stream; L2 flushed between timed steps
  e2e          same step through the public API (`VAE.train_step`) from PINNED HOST buffers: H2D of x and eps, D2H of
               the loss scalars, wall clock, every step
  roofline     the dominant kernel of the step (tcf_kernel: the whole-step tcgen05 kernel), its device time measured with
               CUDA events around every launch INSIDE the timed region (vms_elbo_plan_set_timing), against the measured
               bf16 peak (MEASURED_PEAKS.json) and against the FFMA / tcgen05 issue peaks measured live by the library's
               probe kernels (csrc/probe.cu); HBM-bound kernels timed alone are under "kernels"
  legs         the other BASELINE.json configurations as short legs, each with value / e2e / roofline / cpu_baseline:
               c1 (Gaussian VAE), c3 (neighbour selection through DistanceSelection.__call__, rows sharded over ranks),
               c4a_mc (= "mc": MC proposals/sec, 65,536 chains over all GPUs), c5 (= "large_batch": global batch 262,144)
  cpu_baseline the NumPy oracle (CPU restatement; TF/TFP are not installable here) on a bounded sample, rank 0
`--impl reference` times that CPU restatement as the reference arm (the reference's TF path cannot run in this image).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dx, dz, hidden, prior, num_blocks, num_bins, flow_hidden, batch per GPU)
    'c2': dict(dx=6, dz=2, hidden=200, prior='realnvp', num_blocks=4, num_bins=32, flow_hidden=100, batch=4096,
               label='C2 flow-prior VAE (RealNVP-RQS 4 blocks K=32 H=100 over N(0,I_2); enc 6-200-4, dec 2-200-12), '
               'ELBO fwd+bwd+Adam, batch 4096 per GPU'),
    'c1': dict(dx=6, dz=2, hidden=200, prior='normal', num_blocks=0, num_bins=32, flow_hidden=100, batch=4096,
               label='C1 Gaussian VAE (N(0,I) prior; enc 6-200-4, dec 2-200-12), ELBO fwd+bwd+Adam, batch 4096 per GPU'),
}
METRIC = 'configs/sec ELBO fwd+bwd'
UNIT = 'configs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='rows per GPU (default: the named config, 4096)')
    ap.add_argument('--no-extras', action='store_true', help='skip roofline microbenchmarks and the CPU baseline')
    ap.add_argument('--collective', default='auto', choices=['auto', 'peer', 'nccl'],
                    help='N > 1: fused NVLink peer-memory allreduce+Adam kernel (peer), NCCL allreduce + Adam (nccl)')
    ap.add_argument('--c4b-only', action='store_true', help='development aid: run only the C4b MC leg and print it')
    ap.add_argument('--backmap-only', action='store_true', help='development aid: run only the backmapping leg and print it')
    ap.add_argument('--mc-only', action='store_true', help='development aid: run only the MC leg and print it')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------- CPU arm
def oracle_step_fn(w):
    """One CPU step of the same workload on the NumPy oracle: ELBO forward + analytic backward + Adam."""
    from oracle import vae as ovae
    P = ovae.init_vae(2003, dx=w['dx'], dz=w['dz'], hidden=w['hidden'], prior=w['prior'], num_blocks=w['num_blocks'],
                      num_bins=w['num_bins'], flow_hidden=w['flow_hidden'])
    theta = ovae.flatten(ovae.param_list(P))
    m, v = np.zeros_like(theta), np.zeros_like(theta)
    state = {'t': 0}

    def step(x, eps):
        out, G = ovae.elbo_backward(P, x, eps)
        state['t'] += 1
        g = ovae.flatten(ovae.grad_list(P, G))
        ovae.adam_step(theta, g, m, v, state['t'])  # (the oracle keeps per-layer arrays; the flat update is timed too)
        return float(out['loss'])

    return step


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu(w, rows, steps, warmup):
    step = oracle_step_fn(w)
    rng = np.random.default_rng(1001)
    x = rng.standard_normal((rows, w['dx']), dtype=np.float32)
    eps = rng.standard_normal((rows, w['dz']), dtype=np.float32)
    for _ in range(warmup):
        step(x, eps)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(x, eps)
    dt = time.perf_counter() - t0
    return rows * steps / dt, dt / steps


# The same step with the rows of the batch split over worker PROCESSES (one per host core, BLAS pinned to one thread
# each): NumPy's elementwise arithmetic is single-threaded, so a single process leaves most of the host idle.  Every worker
# runs the oracle's forward + analytic backward on its rows and writes its share of the mean gradient into a shared
# buffer; the parent sums the shares in worker order and applies Adam -- data-parallel training on the CPU.
_PAR = {}


def _par_init(w, x, eps, shm, n_params):
    try:
        from threadpoolctl import threadpool_limits
        _PAR['limit'] = threadpool_limits(1)
    except Exception:  # noqa: BLE001 -- without threadpoolctl the workers keep their BLAS pools
        pass
    from oracle import vae as ovae
    _PAR['ovae'] = ovae
    _PAR['P'] = ovae.init_vae(2003, dx=w['dx'], dz=w['dz'], hidden=w['hidden'], prior=w['prior'], num_blocks=w['num_blocks'],
                              num_bins=w['num_bins'], flow_hidden=w['flow_hidden'])
    _PAR['x'], _PAR['eps'] = x, eps
    _PAR['g'] = np.frombuffer(shm, np.float32).reshape(-1, n_params)


def _par_task(task):
    i, lo, hi, n = task
    ovae, P = _PAR['ovae'], _PAR['P']
    out, G = ovae.elbo_backward(P, _PAR['x'][lo:hi], _PAR['eps'][lo:hi])
    share = np.float32((hi - lo) / n)
    _PAR['g'][i][:] = ovae.flatten(ovae.grad_list(P, G)) * share
    return float(out['loss']) * float(share)


def time_cpu_parallel(w, rows, steps, warmup, workers):
    import multiprocessing as mp
    from oracle import vae as ovae
    P = ovae.init_vae(2003, dx=w['dx'], dz=w['dz'], hidden=w['hidden'], prior=w['prior'], num_blocks=w['num_blocks'],
                      num_bins=w['num_bins'], flow_hidden=w['flow_hidden'])
    theta = ovae.flatten(ovae.param_list(P))
    m, v = np.zeros_like(theta), np.zeros_like(theta)
    rng = np.random.default_rng(1001)
    x = rng.standard_normal((rows, w['dx']), dtype=np.float32)
    eps = rng.standard_normal((rows, w['dz']), dtype=np.float32)
    workers = max(1, min(workers, rows // 32))
    ctx = mp.get_context('fork')
    shm = ctx.RawArray('f', workers * theta.size)
    g_all = np.frombuffer(shm, np.float32).reshape(workers, theta.size)
    cuts = [rows * i // workers for i in range(workers + 1)]
    tasks = [(i, cuts[i], cuts[i + 1], rows) for i in range(workers)]
    with ctx.Pool(workers, initializer=_par_init, initargs=(w, x, eps, shm, theta.size)) as pool:
        t = 0

        def step():
            nonlocal t
            losses = pool.map_async(_par_task, tasks, chunksize=1).get(timeout=300)
            t += 1
            ovae.adam_step(theta, g_all.sum(axis=0), m, v, t)
            return sum(losses)

        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    return rows * steps / dt, dt / steps, workers


def cpu_baseline_subprocess(args, w, batch):
    """The main workload's CPU baseline for the GPU arm: the reference arm itself (worker processes over all host cores) in a
    child interpreter -- this process holds a CUDA context and must not fork workers."""
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '20', '--warmup', '3',
                            '--workload', args.workload, '--batch', str(batch), '--no-extras'],
                           env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        ref = json.loads(r.stdout.strip().splitlines()[-1])
        cb = dict(ref['cpu_baseline'])
        cb['ms_per_step'] = ref['ms_per_step']
        return cb
    except Exception as ex:  # noqa: BLE001 -- fall back to one process in this interpreter
        cpu_val, cpu_sec = time_cpu(w, batch, 10, 2)
        return {'value': cpu_val, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                'sample': '10 steps of %d configs, NumPy oracle in one process (child interpreter failed: %s)' % (batch, ex),
                'ms_per_step': cpu_sec * 1e3}


def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = args.batch or w['batch']
    # bounded sample: probe one full batch, then size the per-step sample so the whole run stays within ~2 minutes
    _, t_full = time_cpu(w, batch, 1, 1)
    budget = 120.0
    rows = batch
    cores = cpu_threads()
    how = 'rows split over %d worker processes (one BLAS thread each), shares of the mean gradient summed by the parent'
    try:
        value, sec, used = time_cpu_parallel(w, rows, args.steps, args.warmup, cores)
        how = how % used
    except Exception as ex:  # noqa: BLE001 -- a host that cannot fork workers still gets the single-process number
        total = (args.steps + args.warmup) * t_full
        if total > budget:
            rows = max(32, int(batch * budget / total) // 32 * 32)
        value, sec = time_cpu(w, rows, args.steps, args.warmup)
        used, how = 1, 'single process (worker pool failed: %s: %s)' % (type(ex).__name__, ex)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': w['label'], 'rows_per_step': rows,
                   'note': 'CPU NumPy restatement of the reference path (oracle/); TF<=2.15 / TFP<=0.23 are not '
                           'installable in this image (Python 3.12, no network)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': used, 'kind': 'port',
                         'sample': '%d steps of %d configs (of the %d-config batch); %s; one process alone: %.1f ms per '
                                   '%d-config step' % (args.steps, rows, batch, how, t_full * 1e3, batch)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if not args.no_extras:
        mcb = mc_cpu_baseline()
        line['mc'] = {'metric': 'MC proposals/sec', 'value': mcb['value'], 'unit': 'proposals/s', 'workload': MC_LABEL,
                      'cpu_baseline': mcb}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def count(self):
        return len(self.rows)

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, ln in self.rows:
            if t < t0 or t > t1:
                continue
            f = [s.strip() for s in ln.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------- GPU arm
def pinned_array(lib, shape, dtype=np.float32):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    lib.vms_malloc_host(C.byref(p), n)
    buf = (C.c_byte * n).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Events(object):

    def __init__(self, c, n):
        self.c = c
        self.ev = []
        for _ in range(2 * n):
            e = C.c_void_p()
            c.lib.vms_event_create(C.byref(e))
            self.ev.append(e.value)

    def record(self, i):
        self.c.lib.vms_event_record(self.ev[i], self.c.stream)

    def elapsed_ms(self, i, j):
        ms = C.c_float(0)
        self.c.lib.vms_event_elapsed_ms(self.ev[i], self.ev[j], C.byref(ms))
        return ms.value


def build_model(v, w, batch):
    from vaemolsim_b200 import dists, flows, losses, models
    import vaemolsim_b200._protocols as PR
    v.set_seed(2003)
    enc = models.MappingToDistribution(PR.IndependentNormal(w['dz']), name='encoder')
    dec = models.MappingToDistribution(PR.IndependentNormal(w['dx']), name='decoder')
    enc.mapping.hidden_dim = [w['hidden']]
    dec.mapping.hidden_dim = [w['hidden']]
    latent = PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], w['dz']))
    if w['prior'] == 'normal':
        prior = latent
    else:
        flow = flows.RQSSplineRealNVP(num_blocks=w['num_blocks'],
                                      rqs_params=dict(bin_range=[-10.0, 10.0], num_bins=w['num_bins'],
                                                      hidden_dim=w['flow_hidden']))
        prior = dists.FlowedDistribution(flow, latent)
    model = models.VAE(enc, dec, prior, regularizer=losses.KLDivergenceEstimate())
    model(np.zeros((2, w['dx']), np.float32))  # builds every layer (reference idiom: call once on data)
    model.compile(optimizer=models.Adam(learning_rate=1e-3), loss=losses.LogProbLoss())
    model.fused(batch)
    return model


def load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, name)))
    except (OSError, ValueError):
        return {}


def measured_peaks(v):
    """Roofline denominators: the driver-written MEASURED_PEAKS.json (HBM copy, cuBLAS bf16) and, measured LIVE by the
    library's probe kernels (csrc/probe.cu), the FP32 FFMA issue peak and the tcgen05 bf16 / tf32 issue peaks
    (BASELINE.md section 2 asks for them)."""
    c = v._abi.ctx()
    drv = load_json('MEASURED_PEAKS.json')
    out = {'hbm_gbs': float(drv.get('hbm_gbs', 6650.0)), 'bf16_tflops': float(drv.get('bf16_tflops', 1590.0)),
           'bf16_tflops_sustained': float(drv.get('bf16_tflops_sustained', 1400.0)),
           'kind': 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in drv else 'fallback (B200_PROFILING.md)'}
    t, ms = C.c_double(0), C.c_double(0)
    c.lib.vms_probe_ffma(2048, 3, C.byref(t), C.byref(ms), c.stream)
    out['fp32_ffma_tflops'] = t.value
    out['fp32_ffma_nominal_tflops'] = c.sm_count * 128 * 2 * 1.965e9 / 1e12
    c.lib.vms_probe_ffma2(2048, 3, C.byref(t), C.byref(ms), c.stream)
    out['fp32_ffma2_packed_tflops'] = t.value  # fma.rn.f32x2: the same FLOP rate with half the issue slots
    for kind, name in ((0, 'bf16'), (1, 'tf32')):
        for M, N in ((128, 256), (64, 96)):
            c.lib.vms_probe_mma(kind, M, N, 8000, 3, C.byref(t), C.byref(ms), c.stream)
            out['tcgen05_%s_m%d_n%d_tflops' % (name, M, N)] = t.value
    out['probe'] = 'csrc/probe.cu, best of 3 launches, CUDA events (this run)'
    return out


def kernel_microbench(v, w, batch, peak, reps=20):
    """CUDA-event timings of the HBM-bound kernels, alone, at streaming sizes (cold L2)."""
    c = v._abi.ctx()
    K = w['num_bins']
    flush = v.Tensor((64 << 20, ))  # 256 MiB > 126 MB L2
    ev = Events(c, 1)
    rng = np.random.default_rng(5)
    res = {}

    def timed(fn, n_rep, do_flush=True):
        ts = []
        for _ in range(n_rep):
            if do_flush:
                c.lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
            ev.record(0)
            fn()
            ev.record(1)
            c.synchronize()
            ts.append(ev.elapsed_ms(0, 1))
        return float(np.mean(ts[2:])) if len(ts) > 4 else float(np.mean(ts))

    n = 1 << 21
    rw = v.Tensor.from_numpy(rng.standard_normal((n, K), dtype=np.float32))
    rh = v.Tensor.from_numpy(rng.standard_normal((n, K), dtype=np.float32))
    rs = v.Tensor.from_numpy(rng.standard_normal((n, K - 1), dtype=np.float32))
    x = v.Tensor.from_numpy(rng.uniform(-10, 10, n).astype(np.float32))
    g = v.Tensor.from_numpy(rng.standard_normal(n, dtype=np.float32))
    y, l = v.Tensor((n, )), v.Tensor((n, ))
    gi, gw, gh, gs = v.Tensor((n, )), v.Tensor((n, K)), v.Tensor((n, K)), v.Tensor((n, K - 1))
    fwd_bytes = n * (4 * (3 * K - 1) + 12)
    bwd_bytes = n * (2 * 4 * (3 * K - 1) + 4 + 8 + 4)
    for name, nbytes, fn in (
        ('rqs_forward', fwd_bytes, lambda: c.lib.vms_rqs_forward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, y.ptr, l.ptr,
                                                                 c.stream)),
        ('rqs_inverse', fwd_bytes, lambda: c.lib.vms_rqs_inverse(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, y.ptr, l.ptr,
                                                                 c.stream)),
        ('rqs_backward', bwd_bytes, lambda: c.lib.vms_rqs_backward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0, 1, g.ptr,
                                                                   g.ptr, gi.ptr, gw.ptr, gh.ptr, gs.ptr, c.stream)),
    ):
        ms = timed(fn, reps)
        gbs = nbytes / (ms * 1e-3) / 1e9
        res['%s@stream' % name] = {'n_elem': n, 'ms': ms, 'bytes': nbytes, 'gbs': gbs, 'frac': gbs / peak}
    del rw, rh, rs, gw, gh, gs
    n, D = 1 << 22, w['dx']
    xx = v.Tensor.from_numpy(rng.standard_normal((n, D), dtype=np.float32))
    pp = v.Tensor.from_numpy(rng.standard_normal((n, 2 * D), dtype=np.float32))
    lp = v.Tensor((n, ))
    i32 = lambda a: (C.c_int32 * len(a))(*a)
    kind, loc, loc2, sc = i32([0] * D), i32(list(range(D))), i32([-1] * D), i32(list(range(D, 2 * D)))
    ms = timed(lambda: c.lib.vms_blockwise_log_prob(xx.ptr, D, pp.ptr, 2 * D, n, D, kind, loc, loc2, sc, 1, lp.ptr, 0,
                                                    c.stream), reps)
    nbytes = n * (3 * D * 4 + 4)
    res['normal_log_prob@stream'] = {'n_rows': n, 'ms': ms, 'bytes': nbytes, 'gbs': nbytes / ms / 1e6,
                                     'frac': nbytes / ms / 1e6 / peak}
    return res


# ---------------------------------------------------------------------------------------------------- C3 leg
C3 = dict(rows=4096, N=10000, L=46.416, k=50, P=2, cutoff=3.0)
C3_LABEL = ('C3 backmapping selection: DistanceSelection(cutoff 3.0, max_included 50, box 46.416^3) over one frame of 10,000 '
            'particles replicated per reference row (the API takes [B, N, 3]), one-hot particle_info P = 2, int32 indices; '
            '4096 reference rows sharded over the GPUs')


def c3_leg(v, grp, peak_hbm, reps=10):
    """Neighbour selection at the C3 shape: `value` = rows/s with coords / info resident in HBM (CUDA events, L2 flushed),
    `e2e` = the same call through `DistanceSelection.__call__` from pinned HOST arrays (H2D of coords / ref / info and D2H
    of the selection inside the timed region), rows sharded over the ranks, no collective."""
    from vaemolsim_b200 import parallel
    from oracle import mappings as omap
    c = v._abi.ctx()
    lib = c.lib
    lo, hi = parallel.shard_rows(C3['rows'], grp.rank, grp.world)
    B, N, k, P = hi - lo, C3['N'], C3['k'], C3['P']
    L = np.float32(C3['L'])
    rng = np.random.default_rng(3001)
    frame = rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32)
    kinds = rng.integers(0, 2, N)
    ref_all = np.random.default_rng(3002).uniform(-L / 2, L / 2, (C3['rows'], 1, 3)).astype(np.float32)
    coords_h = pinned_array(lib, (B, N, 3))
    info_h = pinned_array(lib, (B, N, P))
    ref_h = pinned_array(lib, (B, 1, 3))
    coords_h[...] = frame[None]
    info_h[...] = np.eye(2, dtype=np.float32)[kinds][None]
    ref_h[...] = ref_all[lo:hi]
    box = np.array([L, L, L], np.float32)
    layer = v.mappings.DistanceSelection(C3['cutoff'], max_included=k, box_lengths=box)
    coords_d, info_d, ref_d = v.Tensor.from_numpy(coords_h), v.Tensor.from_numpy(info_h), v.Tensor.from_numpy(ref_h)
    flush = v.Tensor((64 << 20, ))
    ev = Events(c, 1)
    for _ in range(3):
        out = layer(coords_d, ref_d, particle_info=info_d, return_indices=True)
    c.synchronize()
    grp.barrier()
    l0 = v._abi.launch_count()
    dev_ms = 0.0
    for _ in range(reps):
        lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
        ev.record(0)
        out = layer(coords_d, ref_d, particle_info=info_d, return_indices=True)
        ev.record(1)
        c.synchronize()
        dev_ms += ev.elapsed_ms(0, 1)
    launches = v._abi.launch_count() - l0  # (vms_memset is a runtime call, not counted)
    dev_ms = grp.max(dev_ms) / reps
    # end to end from pinned host arrays through the layer call
    n_e2e = 3
    sel = layer(coords_h, ref_h, particle_info=info_h, return_indices=True)
    [t.numpy() for t in sel]
    grp.barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        sel = layer(coords_h, ref_h, particle_info=info_h, return_indices=True)
        got = [t.numpy() for t in sel]
    e2e_s = grp.max(time.perf_counter() - t0) / n_e2e
    bytes_row = 12 * N + 12 + k * (12 + 4 * P) + 4 * k * P + 4 * k  # SURVEY 8d (+ the int32 indices)
    gbs = B * bytes_row / (dev_ms * 1e-3) / 1e9
    res = {'metric': 'reference rows/sec (neighbour selection)', 'value': C3['rows'] / (dev_ms * 1e-3), 'unit': 'rows/s',
           'workload': C3_LABEL, 'rows_per_gpu': B, 'ms_per_call': dev_ms, 'scaling': 'strong (rows sharded, no collective)',
           'gpu_launches_per_call': max(1, int(round(launches / reps))) if launches > 0 else 1,
           'e2e': {'value': C3['rows'] / e2e_s, 'unit': 'rows/s', 'ms_per_call': e2e_s * 1e3,
                   'h2d_bytes_per_step': int(coords_h.nbytes + info_h.nbytes + ref_h.nbytes),
                   'd2h_bytes_per_step': int(B * k * (12 + 4 * P + 4)),
                   'api': 'DistanceSelection.__call__(coords, ref, particle_info=..., return_indices=True) from pinned host arrays',
                   'note': 'PCIe-bound by construction: the reference API takes the frame replicated per row, 120 KB per row'},
           'roofline': {'bound': 'hbm', 'kernel': 'dist_select_kernel', 'achieved': gbs, 'peak': peak_hbm, 'unit': 'GB/s',
                        'frac': gbs / peak_hbm, 'algorithmic_bytes_per_row': bytes_row, 'traffic': None}}
    # the same selection through the shared-frame entry (`DistanceSelection.select_from_frame`, vms_dist_select_frame): the C3
    # workload IS one box of 10,000 particles with 4,096 sites -- the reference API makes the caller tile the box per site
    frame_d, kinds_d = v.Tensor.from_numpy(frame), v.Tensor.from_numpy(np.eye(2, dtype=np.float32)[kinds])
    for _ in range(3):
        fsel = layer.select_from_frame(frame_d, ref_d, particle_info=kinds_d, return_indices=True)
    c.synchronize()
    grp.barrier()
    f_ms = 0.0
    for _ in range(reps):
        lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
        ev.record(0)
        fsel = layer.select_from_frame(frame_d, ref_d, particle_info=kinds_d, return_indices=True)
        ev.record(1)
        c.synchronize()
        f_ms += ev.elapsed_ms(0, 1)
    f_ms = grp.max(f_ms) / reps
    frame_h, kinds_h = pinned_array(lib, (N, 3)), pinned_array(lib, (N, P))
    frame_h[...] = frame
    kinds_h[...] = np.eye(2, dtype=np.float32)[kinds]
    fgot = [t.numpy() for t in layer.select_from_frame(frame_h, ref_h, particle_info=kinds_h, return_indices=True)]
    grp.barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        fgot = [t.numpy() for t in layer.select_from_frame(frame_h, ref_h, particle_info=kinds_h, return_indices=True)]
    f_e2e = grp.max(time.perf_counter() - t0) / n_e2e
    res['shared_frame'] = {
        'value': C3['rows'] / (f_ms * 1e-3), 'unit': 'rows/s', 'ms_per_call': f_ms,
        'e2e': {'value': C3['rows'] / f_e2e, 'unit': 'rows/s', 'ms_per_call': f_e2e * 1e3,
                'h2d_bytes_per_step': int(frame_h.nbytes + kinds_h.nbytes + ref_h.nbytes),
                'd2h_bytes_per_step': int(B * k * (12 + 4 * P + 4)),
                'api': 'DistanceSelection.select_from_frame(frame [N, 3], ref [B, 3], particle_info=[N, P], return_indices=True)'},
        'equals_tiled_call': bool(all(np.array_equal(a, b) for a, b in zip(fgot, got))),
        'note': 'extension without a reference counterpart: one frame shared by all sites, read from HBM once and then from '
                'L2; the roofline above is for the reference-shaped (tiled) call'}
    if grp.rank == 0:
        n_cpu = 32
        t0 = time.perf_counter()
        want = omap.distance_selection(coords_h[:n_cpu], ref_h[:n_cpu].reshape(n_cpu, 3), C3['cutoff'], k, box_lengths=box,
                                       particle_info=info_h[:n_cpu], return_indices=True)
        dt = time.perf_counter() - t0
        res['cpu_baseline'] = {'value': n_cpu / dt, 'unit': 'rows/s', 'cores': cpu_threads(), 'kind': 'port',
                               'sample': '%d rows, oracle/mappings.py (restatement of mappings.py:362-455)' % n_cpu}
        res['parity_vs_oracle'] = bool(np.array_equal(got[0][:n_cpu], want[0]) and np.array_equal(got[1][:n_cpu], want[1])
                                       and np.array_equal(got[2][:n_cpu], want[2]))
    return res


# --------------------------------------------------------------- backmapping descriptors (GAA embedding)
def backmap_leg(v, grp, ffma_peak, reps=5):
    """Local descriptors of the C3 box: 4,096 sites select their k nearest of 10,000 particles (shared frame) and embed the
    selected cloud with `ParticleEmbedding(20)` (2 attention blocks + the permutation-invariant final attention, hidden 40,
    mappings.py:564-688) -- one fused kernel per attention layer.  `value` = sites/s device-resident at k = 50 (the layer's
    default `max_included`), sites sharded over the ranks; k = 10 (the backmapping notebook) beside it; the op-by-op
    (training) path timed on 256 sites; a `BackmappingOnly` training step at the shape of tests/test_models.py:265-308."""
    from vaemolsim_b200 import parallel
    from oracle import gaa as ogaa
    from oracle import mappings as omap
    c = v._abi.ctx()
    M = v.mappings
    lo, hi = parallel.shard_rows(C3['rows'], grp.rank, grp.world)
    B, N, P, E, H = hi - lo, C3['N'], C3['P'], 20, 40
    L = np.float32(C3['L'])
    rng = np.random.default_rng(3001)
    frame = rng.uniform(-L / 2, L / 2, (N, 3)).astype(np.float32)
    kinds = np.eye(2, dtype=np.float32)[rng.integers(0, 2, N)]
    ref = np.random.default_rng(3002).uniform(-L / 2, L / 2, (C3['rows'], 3)).astype(np.float32)[lo:hi]
    box = np.array([L, L, L], np.float32)
    frame_d, kinds_d, ref_d = v.Tensor.from_numpy(frame), v.Tensor.from_numpy(kinds), v.Tensor.from_numpy(ref)
    v.set_seed(8)
    ev = Events(c, 1)
    res = {}
    flop_pair = 2 * (2 * H + H * E + 2 * E * E + E * H + H)  # value net, two join products, score net (per pair and layer)
    for k in (50, 10):
        sel = M.DistanceSelection(C3['cutoff'], max_included=k, box_lengths=box)
        pe = M.ParticleEmbedding(E, hidden_dim=H)
        xyz, info = sel.select_from_frame(frame_d, ref_d, particle_info=kinds_d)
        for _ in range(2):
            out = pe(xyz, info)
        c.synchronize()
        grp.barrier()
        l0 = v._abi.launch_count()
        ms = 0.0
        for _ in range(reps):
            ev.record(0)
            out = pe(xyz, info)
            ev.record(1)
            c.synchronize()
            ms += ev.elapsed_ms(0, 1)
        launches = (v._abi.launch_count() - l0) / reps
        ms = grp.max(ms) / reps
        ms_all = 0.0
        for _ in range(reps):  # selection + embedding (LocalParticleDescriptors on a shared frame)
            ev.record(0)
            xyz2, info2 = sel.select_from_frame(frame_d, ref_d, particle_info=kinds_d)
            out2 = pe(xyz2, info2)
            ev.record(1)
            c.synchronize()
            ms_all += ev.elapsed_ms(0, 1)
        ms_all = grp.max(ms_all) / reps
        # end to end from HOST arrays: upload of the frame / sites, selection, embedding, descriptors read back
        d_host = pe(*sel.select_from_frame(frame, ref, particle_info=kinds)).numpy()
        grp.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            d_host = pe(*sel.select_from_frame(frame, ref, particle_info=kinds)).numpy()
        e2e_s = grp.max(time.perf_counter() - t0) / reps
        flop = B * k * k * 3 * flop_pair
        rec = {'sites_per_gpu': B, 'ms_embedding': ms, 'ms_select_and_embed': ms_all, 'gpu_launches': launches,
               'sites_per_s': C3['rows'] / (ms * 1e-3), 'sites_per_s_with_selection': C3['rows'] / (ms_all * 1e-3),
               'e2e': {'value': C3['rows'] / e2e_s, 'unit': 'sites/s', 'ms_per_call': e2e_s * 1e3,
                       'h2d_bytes_per_step': int(frame.nbytes + kinds.nbytes + ref.nbytes), 'd2h_bytes_per_step': int(B * E * 4),
                       'api': 'ParticleEmbedding(*DistanceSelection.select_from_frame(frame, ref, particle_info=...)).numpy() '
                              'from host arrays (the shared-frame form of LocalParticleDescriptors)'},
               'tflops': flop / (ms * 1e-3) / 1e12, 'frac_of_fp32_ffma_measured': flop / (ms * 1e-3) / 1e12 / ffma_peak}
        # the op-by-op (training) path on a slice, same weights: pair tensors in HBM
        nb = min(256, B)
        xs, is_ = v.Tensor.from_numpy(xyz.numpy()[:nb]), v.Tensor.from_numpy(info.numpy()[:nb])
        os.environ['VMS_GAA_FUSED'] = '0'
        try:
            o_ops = pe(xs, is_)
            c.synchronize()
            l0 = v._abi.launch_count()
            ev.record(0)
            o_ops = pe(xs, is_)
            ev.record(1)
            c.synchronize()
            rec['op_by_op'] = {'sites': nb, 'ms': ev.elapsed_ms(0, 1), 'gpu_launches': v._abi.launch_count() - l0,
                               'sites_per_s': nb / (ev.elapsed_ms(0, 1) * 1e-3)}
        finally:
            del os.environ['VMS_GAA_FUSED']
        rec['fused_equals_op_by_op_1e-5'] = bool(np.allclose(out.numpy()[:nb], o_ops.numpy(), rtol=1e-5,
                                                            atol=1e-5 * float(np.abs(o_ops.numpy()).max())))
        if grp.rank == 0:  # oracle on a bounded sample: parity flag and CPU baseline
            n_cpu = 64 if k == 50 else 256
            w = {'info': [pe.info_net.kernel.numpy(), pe.info_net.bias.numpy()], 'blocks': [], 'final': None}

            def att_w(va):
                s_, v_ = va.score_net.layers, va.value_net.layers
                return {'merge': [t.numpy() for t in va.merge_kernels], 'join': [t.numpy() for t in va.join_kernels],
                        'score': [s_[0].kernel.numpy(), s_[0].bias.numpy(), s_[1].kernel.numpy(), s_[1].bias.numpy()],
                        'value': [v_[0].kernel.numpy(), v_[0].bias.numpy(), v_[1].gamma.numpy(), v_[1].beta.numpy(),
                                  v_[3].kernel.numpy(), v_[3].bias.numpy()]}

            for blk in pe.block_list:
                wb = att_w(blk.attn)
                nl = blk.nonlinearity.layers
                wb['nonlin'] = [nl[0].kernel.numpy(), nl[0].bias.numpy(), nl[1].gamma.numpy(), nl[1].beta.numpy(),
                                nl[3].kernel.numpy(), nl[3].bias.numpy()]
                w['blocks'].append(wb)
            w['final'] = att_w(pe.final_attn)
            xh, ih = xyz.numpy()[:n_cpu], info.numpy()[:n_cpu]
            t0 = time.perf_counter()
            want32 = ogaa.particle_embedding(xh, ih, w)
            dt = time.perf_counter() - t0
            want = ogaa.particle_embedding(xh.astype(np.float64), ih.astype(np.float64), ogaa.cast(w, np.float64))
            err = float(np.abs(out.numpy()[:n_cpu] - want).max() / np.abs(want).max())
            rec['max_rel_err_vs_float64_oracle'] = err
            rec['parity_vs_oracle_1e-5'] = bool(err < 1e-5)
            rec['cpu_baseline'] = {'value': n_cpu / dt, 'unit': 'sites/s', 'cores': cpu_threads(), 'kind': 'port',
                                   'sample': '%d sites, float32 NumPy oracle/gaa.py (restatement of the published '
                                             'geometric-algebra-attention algorithm; parity unpinned)' % n_cpu}
        res['k_%d' % k] = rec
    # one training step of BackmappingOnly at the reference test's shape (tests/test_models.py:265-308; the backmapping
    # notebook's progress bar shows 37-39 ms/step at batch 32 with a 3-block conditional MAF decoder)
    D = v.dists
    n_b = 32
    r2 = np.random.default_rng(9)
    coords = [r2.uniform(-5, 5, (40, 3)).astype(np.float32) for _ in range(n_b)]
    infos = [np.eye(2, dtype=np.float32)[r2.integers(0, 2, 40)] for _ in range(n_b)]
    x = [r2.uniform(-5, 5, (n_b, 1, 3)).astype(np.float32), M.RaggedTensor.from_rows(coords, inner=3),
         M.RaggedTensor.from_rows(infos, inner=2)]
    y = r2.uniform(-3, 3, (n_b, 6)).astype(np.float32)
    bm = v.models.BackmappingOnly(
        M.LocalParticleDescriptors(M.DistanceSelection(3.0, max_included=10, box_lengths=[10.0, 10.0, 10.0]),
                                   M.ParticleEmbedding(20)),
        v.models.MappingToDistribution(D.AutoregressiveBlockwise(6, [D.Normal] * 3 + [D.VonMises] * 3), name='decoder'))
    bm.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
    bm(x)
    losses = [bm.train_on_batch(x, y) for _ in range(5)]
    c.synchronize()
    l0 = v._abi.launch_count()
    t0 = time.perf_counter()
    n_st = 20
    for _ in range(n_st):
        losses.append(bm.train_on_batch(x, y))
    c.synchronize()
    dt = (time.perf_counter() - t0) / n_st
    res['train_step'] = {'batch': n_b, 'particles_per_frame': 40, 'max_included': 10, 'ms_per_step': dt * 1e3,
                         'configs_per_s': n_b / dt, 'gpu_launches_per_step': (v._abi.launch_count() - l0) / n_st,
                         'loss_first': losses[0], 'loss_last': losses[-1], 'loss_finite': bool(np.isfinite(losses).all()),
                         'reference_notebook': 'Molecular_Backmapping.ipynb progress bar: 37-39 ms/step at batch 32 (other '
                                               'hardware, 3-block conditional MAF decoder; orientation only)'}
    k50 = res['k_50']
    return {'metric': 'sites/sec (local descriptors: ParticleEmbedding over the selected cloud)', 'value': k50['sites_per_s'],
            'unit': 'sites/s', 'scaling': 'strong (sites sharded, no collective)',
            'workload': 'C3 box (N = 10,000, L = 46.416, cutoff 3.0): %d sites x k nearest particles -> ParticleEmbedding(20, '
                        'hidden 40, 2 blocks), fused forward' % C3['rows'],
            'roofline': {'bound': 'ffma', 'kernel': 'gaa_attention_fwd_x2_kernel<40, 20, relu> (packed fma.rn.f32x2, two pairs per trip)', 'achieved': k50['tflops'], 'peak': ffma_peak,
                         'unit': 'TFLOP/s', 'frac': k50['frac_of_fp32_ffma_measured'], 'traffic': None,
                         'algorithmic_flop_per_pair_and_layer': flop_pair},
            'e2e': k50['e2e'], 'cpu_baseline': k50.get('cpu_baseline'), **res}


# ------------------------------------------------------------------------------ FlowModel NLL training (generic tape path)
def flow_model_leg(v, grp):
    """`FlowModel` of examples/Using_Normalizing_Flows.ipynb (cells 18-24: 1-D RQSSplineRealNVP, 4 blocks, K = 32, H = 100
    over N(0, 1), NLL `fit`) through the generic path -- op-by-op kernels recorded on a tape, reverse mode, Adam per weight
    (no fused plan exists for this family): training steps per second at the notebook's batch size 32 (BASELINE.md quotes the
    notebook's Keras progress bar: 2 ms/step on an M1/M2 Mac) and at batch 4096, host-driven, wall clock."""
    import vaemolsim_b200._protocols as PR
    c = v._abi.ctx()
    v.set_seed(7)
    flow = v.flows.RQSSplineRealNVP(num_blocks=4, rqs_params=dict(bin_range=[-10.0, 10.0], num_bins=32, hidden_dim=100))
    fm = v.models.FlowModel(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 1)))
    fm.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
    rng = np.random.default_rng(21)
    out = {}
    losses = []
    for B, n in ((32, 60), (4096, 30)):
        x = (rng.normal(size=(B, 1)) * 1.5 + 0.5).astype(np.float32)
        for _ in range(5):
            fm.train_on_batch(x, x)
        c.synchronize()
        l0 = v._abi.launch_count()
        t0 = time.perf_counter()
        for _ in range(n):
            losses.append(fm.train_on_batch(x, x))
        c.synchronize()
        dt = (time.perf_counter() - t0) / n
        out['batch_%d' % B] = {'ms_per_step': dt * 1e3, 'configs_per_s': B / dt,
                               'gpu_launches_per_step': (v._abi.launch_count() - l0) / n}
    # the decoder-only model of examples/Training_VAEs_and_Decoders.ipynb cells 39-43 (FCDeepNN + conditional autoregressive von
    # Mises decoder, MADE hidden [10, 100, 10], 3,938 parameters; Keras progress bar: 471-555 us per step at batch 32)
    d = v.dists
    v.set_seed(9)
    dec = v.models.MappingToDistribution(d.AutoregressiveBlockwise(2, [d.VonMises] * 2, conditional=True,
                                                                    conditional_event_shape=(1, ),
                                                                    auto_net_params={'hidden_units': [10, 100, 10]}))
    dec.compile(optimizer=v.models.Adam(1e-3), loss=v.losses.LogProbLoss())
    cg = rng.uniform(-np.pi, np.pi, (32, 1)).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, (32, 2)).astype(np.float32)
    for _ in range(5):
        dec.train_on_batch(cg, ang)
    c.synchronize()
    l0 = v._abi.launch_count()
    t0 = time.perf_counter()
    dl = [dec.train_on_batch(cg, ang) for _ in range(60)]
    c.synchronize()
    dt = (time.perf_counter() - t0) / 60
    out['decoder_only_batch_32'] = {'ms_per_step': dt * 1e3, 'configs_per_s': 32 / dt, 'params': dec.count_params(),
                                    'gpu_launches_per_step': (v._abi.launch_count() - l0) / 60,
                                    'graph_replay': bool(getattr(dec._trainer(), '_graphs', {})),
                                    'loss_finite': bool(np.isfinite(dl).all()),
                                    'reference_notebook': 'Training_VAEs_and_Decoders.ipynb:648-666: 471-555 us/step at batch 32 '
                                                          '(Apple M-series CPU, orientation only)'}
    out['graph_replay'] = bool(getattr(fm._trainer(), '_graphs', {}))
    return {'metric': 'FlowModel NLL training steps (tape path)', 'value': out['batch_4096']['configs_per_s'], 'unit': 'configs/s',
            'workload': 'Using_Normalizing_Flows.ipynb FlowModel: 1-D RQSSplineRealNVP 4 blocks K=32 H=100 over N(0,1), NLL + Adam',
            'scaling': 'replicas (every rank runs the same steps)', **out, 'loss_finite': bool(np.isfinite(losses).all()),
            'reference_notebook': 'Keras progress bar of the notebook (BASELINE.md 2): 2 ms/step at batch 32 on an Apple M1/M2 '
                                  '= 16-20 k configs/s; other hardware, quoted for orientation only',
            'e2e': {'value': out['batch_4096']['configs_per_s'], 'unit': 'configs/s', 'h2d_bytes_per_step': 4096 * 4 * 2,
                    'd2h_bytes_per_step': 4, 'api': 'FlowModel.train_on_batch(x, x) from host arrays'}}


# ---------------------------------------------------------------------------------------------------- MC leg (C4a)
MC_CHAINS, MC_STEPS = 65536, 100
MC_LABEL = ('C4a VAE-proposal MC: Gaussian VAE (enc 6-200-4, dec 2-200-12, N(0,I) prior), quadratic energy on device, '
            '%d independent chains over all GPUs, %d steps per launch' % (MC_CHAINS, MC_STEPS))


def mc_cpu_baseline(chains=4096, steps=3):
    """The CPU restatement of MCMC.single_step (oracle/mcmc.py, pinned bit for bit by the reference's own mcmc.py
    goldens) over the NumPy oracle VAE: proposals/sec on the host cores."""
    from oracle import mcmc as omc
    from oracle import vae as ovae
    P = ovae.init_vae(2003, dx=6, dz=2, hidden=200, prior='normal')
    model = omc.OracleVAE(P, noise_seed=1)
    rng = np.random.default_rng(4002)
    x = np.random.default_rng(4001).standard_normal((chains, 6)).astype(np.float32)
    x, e, _ = omc.single_step(model, omc.quadratic_energy, rng, x, None)
    t0 = time.perf_counter()
    for _ in range(steps):
        x, e, _ = omc.single_step(model, omc.quadratic_energy, rng, x, e)
    dt = time.perf_counter() - t0
    return {'value': chains * steps / dt, 'unit': 'proposals/s', 'cores': cpu_threads(), 'kind': 'port',
            'sample': '%d steps of %d chains, oracle/mcmc.py (restatement of mcmc.py:68-130) over the NumPy oracle VAE' %
                      (steps, chains)}


def mc_flips_vs_oracle(v, model, chains=1024, steps=20):
    """Decisions of the device path against the oracle restatement of mcmc.py on a small sample of the job (same noise,
    same PCG64 columns): number of chains whose decision trace differs anywhere (tracked in the bench line)."""
    from oracle import mcmc as omc
    from oracle import vae as ovae
    P = ovae.init_vae(2003, dx=6, dz=2, hidden=200, prior='normal')
    theta = np.concatenate([a.reshape(-1) for W, b in P['enc'] + P['dec'] for a in (W, b)]).astype(np.float32)
    f = model.fused()
    saved = f.theta.numpy().copy()
    c = v._abi.ctx()
    c.lib.vms_memcpy_h2d(f.theta.ptr, theta.ctypes.data, theta.nbytes, c.stream)
    c.synchronize()
    x0 = np.random.default_rng(4001).standard_normal((MC_CHAINS, 6), dtype=np.float32)[:chains]
    rng = np.random.default_rng(777)
    noise = np.empty((steps, chains, 10), np.float32)
    for s in range(steps):
        noise[s, :, :2] = rng.standard_normal((chains, 2), dtype=np.float32)
        noise[s, :, 2:4] = rng.standard_normal((chains, 2), dtype=np.float32)
        noise[s, :, 4:] = rng.standard_normal((chains, 6), dtype=np.float32)
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002, stream_layout=(0, MC_CHAINS))
    mc.run_fused(x0, n_steps=steps, noise=noise, trace=True)
    acc_d = mc._last_trace['acc'].astype(bool)
    u = np.random.default_rng(4002).random(size=(steps, MC_CHAINS))[:, :chains]

    class Cols(object):
        s = 0

        def random(self, size):
            self.s += 1
            return u[self.s - 1]

    ovm, cols = omc.OracleVAE(P, noise_seed=777), Cols()
    xo, eo = x0.copy(), None
    acc_o = np.empty((steps, chains), bool)
    for s in range(steps):
        xo, eo, acc_o[s] = omc.single_step(ovm, omc.quadratic_energy, cols, xo, eo)
    c.lib.vms_memcpy_h2d(f.theta.ptr, saved.ctypes.data, saved.nbytes, c.stream)
    c.synchronize()
    return {'chains': chains, 'steps': steps, 'flipped_chains': int((acc_d != acc_o).any(axis=0).sum()),
            'accepted_device': int(acc_d.sum()), 'accepted_oracle': int(acc_o.sum()),
            'host_stream_reruns': int(mc.host_stream_reruns)}


# ---------------------------------------------------------------------------------------------- C4b: the MC notebook's model
# FLOP of the reference procedure per proposal as the fused kernel organises it: 2 encoder + 2 decoder-mapping evaluations (1-hidden-layer FCDeepNN, H = 200) and
# 5 MADE passes ([3 -> 10 -> 100 -> 10 -> 4]; 3 sampling + 2 log_prob).  SURVEY 8d quotes 64,352 for the op-by-op form, which
# also evaluates the prior's 12 conditioner networks per chain; the kernel replaces those by per-call knot tables.
C4B_FLOP = 2 * (2 * 200 + 200 * 2) * 2 + 2 * (200 + 200 * 4) * 2 + 2 * (3 * 10 + 10 * 100 + 100 * 10 + 10 * 4) * 5
C4B_FLOP_EXEC = 2 * (2 * 200 + 200 * 2) * 2 + 2 * (200 + 200 * 4) * 2 + 2 * (3 * 10 + 10 * 100 + 100 * 10 + 10 * 4) * 2
C4B_LABEL = ('C4b: MC notebook model (enc 2-200-2 Normal(1); prior RQSSplineMAF 4 blocks K=20 H=40 over N(0,1); dec FCDeepNN '
             '1-200-(2,2) + AutoregressiveBlockwise(2 Normal, cond 1, hidden [10,100,10])), Gaussian-mixture energy, '
             '%d chains x %d steps' % (MC_CHAINS, MC_STEPS))


def build_c4b_model(v, seed=2003):
    """examples/MC_Moves_with_VAEs.ipynb cells 11-20 through the host API, randomly initialised; the (untrained) decoder's
    output layer is aimed at the mixture's main component so that the chains see both accepts and rejects."""
    import vaemolsim_b200._protocols as PR
    d = v.dists
    v.set_seed(seed)
    enc = v.models.MappingToDistribution(PR.IndependentNormal(1), name='encoder')
    dec_dist = d.AutoregressiveBlockwise(2, [d.Normal] * 2, conditional=True, conditional_event_shape=(1, ),
                                         auto_net_params={'hidden_units': [10, 100, 10]})
    dec = v.models.MappingToDistribution(dec_dist, name='decoder')
    flow = v.flows.RQSSplineMAF(num_blocks=4, rqs_params={'bin_range': [-10.0, 10.0], 'num_bins': 20, 'hidden_dim': 40})
    flow(np.zeros((2, 1), np.float32))
    prior = d.FlowedDistribution(flow, PR.DistributionLambda(lambda t: PR.StandardNormal(t.shape[0], 1)), name='prior')
    model = v.models.VAE(enc, dec, prior)
    model(np.zeros((2, 2), np.float32))
    rng = np.random.default_rng(seed)
    last = [l for l in dec.mapping.layer_list if hasattr(l, 'kernel')][-1]
    inv_softplus = lambda y: float(np.log(np.expm1(y)))
    last.assign(last.kernel.numpy() * np.float32(0.2),
                np.array([-0.5, inv_softplus(0.1), 0.0, inv_softplus(0.5)], np.float32))
    net = dec_dist.auto_net
    arrs = []
    for k, lay in enumerate(net.layers):
        sc = np.float32(0.2 if k == len(net.layers) - 1 else 1.0)
        arrs += [lay.kernel.numpy() * sc, rng.normal(0, 0.3, lay.units).astype(np.float32) * sc, net.cond_kernels[k].numpy() * sc]
    net.set_weights(arrs)
    for bij in flow.chain.bijectors:
        msb = bij.bijector_fn
        for sub in (msb.bin_widths, msb.bin_heights, msb.knot_slopes):
            sub.set_weights([a for lay in sub.layers for a in (lay.kernel.numpy(), rng.normal(0, 0.3, lay.units).astype(np.float32))])
    return model


def c4b_oracle_params(model):
    """The product model's weights in the layout of oracle.mcmc.init_vae_b (checker side only)."""
    dense = lambda m: [(l.kernel.numpy(), l.bias.numpy()) for l in m.layer_list if hasattr(l, 'kernel')]
    net = model.decoder.distribution.auto_net
    made = [dict(W=lay.kernel.numpy(), b=lay.bias.numpy(), Wc=net.cond_kernels[k].numpy(), mask=None)
            for k, lay in enumerate(net.layers)]
    maf = []
    for bij in model.prior.flow.chain.bijectors[::-1]:
        msb = bij.bijector_fn
        maf.append({key: [dict(W=lay.kernel.numpy(), b=lay.bias.numpy(), Wc=None, mask=None) for lay in sub.layers]
                    for key, sub in (('w', msb.bin_widths), ('h', msb.bin_heights), ('s', msb.knot_slopes))})
    return dict(enc=dense(model.encoder.mapping), dec=dense(model.decoder.mapping), made=made, maf=maf, num_bins=20,
                bin_range=(-10.0, 10.0), hidden=200)


def gmm_start(n, seed=5001):
    """Chains start in the notebook's data distribution (cell 5, 41: `mc_sim.run(data_sample, ...)`)."""
    rng = np.random.default_rng(seed)
    probs = np.array([0.7, 0.2, 0.1])
    locs = np.array([[-0.5, 0.0], [1.0, 2.0], [-1.5, 0.0]], np.float32)
    scales = np.array([[0.05, 0.5], [1.0, 0.5], [0.5, 0.2]], np.float32)
    k = rng.choice(3, size=n, p=probs)
    return (locs[k] + scales[k] * rng.standard_normal((n, 2))).astype(np.float32)


def c4b_check_and_cpu(v, model, chains=2048, steps=3):
    """Checker + CPU baseline of the C4b leg (rank 0): the oracle restatement of mcmc.py over the NumPy notebook model with
    the product's weights, timed on the host cores; its decisions against the device path's under the same noise and
    uniforms (chains whose trace differs anywhere)."""
    import vaemolsim_b200._protocols as PR
    from oracle import mcmc as omc
    P = c4b_oracle_params(model)
    x0 = gmm_start(MC_CHAINS)[:chains]
    twin, rng = omc.OracleVAEb(P, noise_seed=888), np.random.default_rng(5002)
    xo, eo = x0, None
    acc_o = np.empty((steps, chains), bool)
    t0 = time.perf_counter()
    for s in range(steps):
        xo, eo, acc_o[s] = omc.single_step(twin, omc.gmm_energy, rng, xo, eo)
    dt = time.perf_counter() - t0
    mc = v.mcmc.MCMC(model, v.mcmc.GaussianMixtureEnergy(), random_seed=5002)
    v.set_seed(888)
    xd, ed = x0, None
    acc_d = np.empty((steps, chains), bool)
    for s in range(steps):
        xd, ed = mc.single_step(xd, energies=ed)
        acc_d[s] = mc._last_acc.numpy().astype(bool)
    return ({'value': chains * steps / dt, 'unit': 'proposals/s', 'cores': cpu_threads(), 'kind': 'port',
             'sample': '%d steps of %d chains, oracle/mcmc.py (restatement of mcmc.py:68-130) over the NumPy notebook model '
                       '(oracle.mcmc.OracleVAEb)' % (steps, chains)},
            {'chains': chains, 'steps': steps, 'flipped_chains': int((acc_d != acc_o).any(axis=0).sum()),
             'accepted_device': int(acc_d.sum()), 'accepted_oracle': int(acc_o.sum())})


def c4b_bench(v, grp, ffma_peak, reps=2):
    """MC proposals/sec for the notebook's model family: chains sharded over the ranks, no collective; whole MC steps in the
    fused kernel (`vms_mc_nb_run`), accept uniforms regenerated on the device from the one PCG64 stream, sampling noise
    keyed by the global chain index.  The op-by-op path of the same model is timed beside it."""
    from vaemolsim_b200 import parallel
    c = v._abi.ctx()
    lo, hi = parallel.shard_rows(MC_CHAINS, grp.rank, grp.world)
    B = hi - lo
    model = build_c4b_model(v)
    energy = v.mcmc.GaussianMixtureEnergy()
    mc = v.mcmc.MCMC(model, energy, random_seed=5002, stream_layout=(lo, MC_CHAINS))
    if mc._nb_plan() is None:
        raise RuntimeError('bench: the fused notebook-family MC plan is unavailable')
    x0 = gmm_start(MC_CHAINS)[lo:hi]
    # device-resident leg: chain state in HBM, noise and uniforms generated in the kernel; one launch = MC_STEPS steps
    xd = v.Tensor.from_numpy(np.ascontiguousarray(x0))
    xd, ed = mc.run_nb(None, n_steps=MC_STEPS, configs_dev=xd)
    for _ in range(2):
        mc.run_nb(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed)
    ev = Events(c, reps)
    grp.barrier()
    c.synchronize()
    l0 = v._abi.launch_count()
    for i in range(reps):
        ev.record(2 * i)
        mc.run_nb(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed)
        ev.record(2 * i + 1)
    c.synchronize()
    grp.barrier()
    launches = v._abi.launch_count() - l0
    dev_ms = grp.max(sum(ev.elapsed_ms(2 * i, 2 * i + 1) for i in range(reps)))
    uncertain = mc.uncertain()
    mc.sync_counters()
    acc_rate = mc.acceptance_rate
    # end to end through the public API: MCMC.run(configs, n_steps) from host arrays
    mc2 = v.mcmc.MCMC(model, energy, random_seed=5002, stream_layout=(lo, MC_CHAINS))
    mc2.run(x0, n_steps=2)
    grp.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        xe, ee = mc2.run(x0, n_steps=MC_STEPS)
    e2e_s = grp.max(time.perf_counter() - t0)
    # the op-by-op path of the same model (what every other model family runs), a few steps
    mc3 = v.mcmc.MCMC(model, energy, random_seed=5002)
    mc3.fuse_notebook = False
    v.set_seed(888 + grp.rank)
    xo, eo = mc3.run_device(None, n_steps=2, configs_dev=v.Tensor.from_numpy(np.ascontiguousarray(x0)))
    c.synchronize()
    t0 = time.perf_counter()
    mc3.run_device(None, n_steps=10, configs_dev=xo, energies_dev=eo)
    c.synchronize()
    op_ms = (time.perf_counter() - t0) / 10 * 1e3
    tot = MC_CHAINS * MC_STEPS * reps
    tflops = tot * C4B_FLOP / (dev_ms * 1e-3) / 1e12
    res = {'metric': 'MC proposals/sec', 'value': tot / (dev_ms * 1e-3), 'unit': 'proposals/s', 'workload': C4B_LABEL,
           'chains_global': MC_CHAINS, 'chains_per_gpu': B, 'steps_per_launch': MC_STEPS, 'launches_timed': reps,
           'scaling': 'strong (65,536 chains split over the GPUs, no data-path collective)',
           'ms_per_mc_step': dev_ms / (reps * MC_STEPS), 'gpu_launches': int(launches), 'acceptance_rate': acc_rate,
           'path': 'mc_nb_kernel (fused, 4 lanes per chain); per call: 24 conditioner launches + 4 knot tables + 1 MC launch',
           'uniform_stream': 'NumPy PCG64 regenerated on the device, uncertain decisions: %d, host-stream re-runs: %d'
                             % (uncertain, mc2.host_stream_reruns),
           'op_by_op_ms_per_mc_step': op_ms,
           'e2e': {'value': tot / e2e_s, 'unit': 'proposals/s', 'h2d_bytes_per_step': int(B * 8 / MC_STEPS),
                   'd2h_bytes_per_step': int(B * 12 / MC_STEPS), 'ms_per_mc_step': e2e_s / (reps * MC_STEPS) * 1e3,
                   'api': 'MCMC.run(configs, n_steps=%d) from host arrays' % MC_STEPS},
           'roofline': {'bound': 'ffma', 'kernel': 'mc_nb_kernel', 'achieved': tflops / grp.world, 'peak': ffma_peak,
                        'unit': 'TFLOP/s', 'frac': tflops / grp.world / ffma_peak, 'traffic': None,
                        'algorithmic_flop_per_proposal': C4B_FLOP,
                        'executed_flop_per_proposal': C4B_FLOP_EXEC,
                        'frac_executed': tflops / grp.world / ffma_peak * C4B_FLOP_EXEC / C4B_FLOP,
                        'note': 'achieved counts the FLOP of the reference procedure (5 MADE passes per proposal: tfp runs '
                                'D + 1 = 3 sampling passes + 2 log_prob passes); the kernel proves 3 of them bit-identical '
                                'from the MADE masks and executes 2 (executed_*)',
                        'peak_kind': 'FP32 FFMA issue peak measured in this run (csrc/probe.cu)'}}
    if grp.world > 1:
        res['weak'] = mc_weak_record(
            v, grp, lambda layout: v.mcmc.MCMC(model, energy, random_seed=5002, stream_layout=layout),
            lambda m, xd, ed: m.run_nb(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed),
            lambda a, b: gmm_start(MC_CHAINS * grp.world)[a:b], 2)
    if grp.rank == 0:
        res['cpu_baseline'], res['flips_vs_oracle'] = c4b_check_and_cpu(v, model)
    return res


def mc_weak_record(v, grp, make_mc, run, x_all_fn, dx, reps=3):
    """The same kernel with 65,536 chains on EVERY rank (weak scaling: 65,536 x N chains, one global PCG64 stream / noise
    stream sharded by chain index) -- reported beside the named strong-scaling configuration when N > 1."""
    c = v._abi.ctx()
    n_global = MC_CHAINS * grp.world
    lo = MC_CHAINS * grp.rank
    mc = make_mc((lo, n_global))
    xd = v.Tensor.from_numpy(np.ascontiguousarray(x_all_fn(lo, lo + MC_CHAINS)))
    xd, ed = run(mc, xd, None)
    run(mc, xd, ed)
    ev = Events(c, reps)
    grp.barrier()
    c.synchronize()
    for i in range(reps):
        ev.record(2 * i)
        run(mc, xd, ed)
        ev.record(2 * i + 1)
    c.synchronize()
    grp.barrier()
    ms = grp.max(sum(ev.elapsed_ms(2 * i, 2 * i + 1) for i in range(reps)))
    return {'value': n_global * MC_STEPS * reps / (ms * 1e-3), 'unit': 'proposals/s', 'chains_global': n_global,
            'chains_per_gpu': MC_CHAINS, 'scaling': 'weak', 'ms_per_mc_step': ms / (reps * MC_STEPS)}


def mc_bench(v, grp, ffma_peak, reps=3):
    """MC proposals/sec: chains sharded over ranks with no collective; the accept uniforms are ONE PCG64 stream
    whose columns every rank regenerates on the device for its own chains, and the sampling noise is keyed by the global
    chain index, so decisions do not depend on the rank count."""
    from vaemolsim_b200 import parallel
    c = v._abi.ctx()
    lo, hi = parallel.shard_rows(MC_CHAINS, grp.rank, grp.world)
    B = hi - lo
    model = build_model(v, WORKLOADS['c1'], 4096)
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002, stream_layout=(lo, MC_CHAINS))
    if mc._fused_plan() is None:
        raise RuntimeError('bench: the fused MC plan is unavailable')
    x0 = np.random.default_rng(4001).standard_normal((MC_CHAINS, 6), dtype=np.float32)[lo:hi]
    # device-resident leg: chain state lives in HBM, uniforms and noise are generated in the kernel; one launch = MC_STEPS
    # steps of B chains
    xd = v.Tensor.from_numpy(np.ascontiguousarray(x0))
    xd, ed = mc.run_fused(None, n_steps=MC_STEPS, configs_dev=xd)  # warm-up (also computes E)
    for _ in range(2):
        mc.run_fused(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed)
    ev = Events(c, reps)
    grp.barrier()
    c.synchronize()
    l0 = v._abi.launch_count()
    for i in range(reps):
        ev.record(2 * i)
        mc.run_fused(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed)
        ev.record(2 * i + 1)
    c.synchronize()
    grp.barrier()
    launches = v._abi.launch_count() - l0
    dev_ms = grp.max(sum(ev.elapsed_ms(2 * i, 2 * i + 1) for i in range(reps)))
    uncertain = mc.uncertain()
    mc.sync_counters()
    acc_rate = mc.acceptance_rate
    # end to end through the public API: MCMC.run(configs, n_steps) from host arrays -- H2D of x, the launch(es), D2H of
    # x / E and of the counters
    mc2 = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002, stream_layout=(lo, MC_CHAINS))
    mc2.run(x0, n_steps=2)
    grp.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        xe, ee = mc2.run(x0, n_steps=MC_STEPS)
    e2e_s = grp.max(time.perf_counter() - t0)
    tot = MC_CHAINS * MC_STEPS * reps
    tflops = tot * 19200 / (dev_ms * 1e-3) / 1e12
    res = {'metric': 'MC proposals/sec', 'value': tot / (dev_ms * 1e-3), 'unit': 'proposals/s', 'workload': MC_LABEL,
           'chains_global': MC_CHAINS, 'chains_per_gpu': B, 'steps_per_launch': MC_STEPS, 'launches_timed': reps,
           'scaling': 'strong (65,536 chains split over the GPUs, no data-path collective)',
           'ms_per_mc_step': dev_ms / (reps * MC_STEPS), 'gpu_launches': int(launches), 'acceptance_rate': acc_rate,
           'uniform_stream': 'NumPy PCG64 regenerated on the device (per-chain LCG jump-ahead), uncertain decisions: %d, '
                             'host-stream re-runs: %d' % (uncertain, mc2.host_stream_reruns),
           'e2e': {'value': tot / e2e_s, 'unit': 'proposals/s', 'h2d_bytes_per_step': int(B * 24 / MC_STEPS),
                   'd2h_bytes_per_step': int(B * 32 / MC_STEPS), 'ms_per_mc_step': e2e_s / (reps * MC_STEPS) * 1e3,
                   'api': 'MCMC.run(configs, n_steps=%d) from host arrays' % MC_STEPS},
           'roofline': {'bound': 'ffma',
                        'kernel': ('mc_chain_warp_kernel (a warp per four chains: fewer than 128 chains per SM)'
                                   if B < 128 * 148 and not os.environ.get('VMS_MC_TPC') else 'mc_chain_kernel'),
                        'achieved': tflops / grp.world, 'peak': ffma_peak,
                        'unit': 'TFLOP/s', 'frac': tflops / grp.world / ffma_peak, 'traffic': None,
                        'algorithmic_flop_per_proposal': 19200,
                        'peak_kind': 'FP32 FFMA issue peak measured in this run (csrc/probe.cu)'},
           'tflops_fp32': tflops}
    if grp.world > 1:
        x_all = lambda a, b: np.random.default_rng(4001).standard_normal((MC_CHAINS * grp.world, 6), dtype=np.float32)[a:b]
        res['weak'] = mc_weak_record(
            v, grp, lambda layout: v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002, stream_layout=layout),
            lambda m, xd, ed: m.run_fused(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed), x_all, 6)
    if grp.rank == 0:
        res['flips_vs_oracle'] = mc_flips_vs_oracle(v, model)
    return res


def large_batch_leg(v, w, opt, grp, peaks, collective='auto', global_batch=262144, steps=5):
    """C5 (BASELINE.json configs[4]: data-parallel training, GLOBAL batch 262,144): every rank takes 262,144 / N rows; one
    step = ELBO forward + backward on the shard (auto mode: the tensor-core plan at these sizes) + the gradient exchange
    (N > 1: the fused peer-memory allreduce + Adam kernel, NCCL fallback) + Adam.  Strong scaling.  Device-resident
    inputs, CUDA events on the launching stream, max over ranks."""
    from vaemolsim_b200 import parallel
    c = v._abi.ctx()
    world, rank = grp.world, grp.rank
    lo, hi = parallel.shard_rows(global_batch, rank, world)
    batch = hi - lo
    model = build_model(v, w, batch)
    f = model.fused(batch)
    rng = parallel.global_row_seed(77, lo)
    x = v.Tensor.from_numpy(rng.standard_normal((batch, w['dx']), dtype=np.float32))
    e = v.Tensor.from_numpy(rng.standard_normal((batch, w['dz']), dtype=np.float32))
    peer = None
    if world > 1 and collective in ('auto', 'peer'):
        try:
            peer = parallel.PeerExchange(grp, f.n_params)
        except Exception as ex:
            sys.stderr.write('bench (C5 leg): peer exchange unavailable (%s); using NCCL\n' % ex)

    def step():
        if world == 1:
            f.train_step(x, e, opt)
        elif peer is not None:
            peer.train_step(f, x, e, x.shape[0], opt)
        else:
            f.forward_backward(x, e)
            grp.allreduce_sum_device_(f.grad.ptr, f.n_params, c.stream)
            f.adam_step(opt, grad_scale=1.0 / world)

    for _ in range(3):
        step()
    ev = Events(c, 1)
    c.synchronize()
    grp.barrier()
    ev.record(0)
    for _ in range(steps):
        step()
    ev.record(1)
    c.synchronize()
    grp.barrier()
    ms = grp.max(ev.elapsed_ms(0, 1)) / steps
    # end to end: the same steps through the public pipelined loop (`FusedELBO.train_loop`, what `VAE.fit` runs), every step's
    # shard of x / eps leaving PINNED host memory on the copy stream while the previous step trains, every step's scalars
    # read back; the peer exchange runs inside the loop when N > 1
    e2e = None
    if world == 1 or peer is not None:
        n_e2e = steps
        xh, eh = pinned_array(c.lib, (2 * batch, w['dx'])), pinned_array(c.lib, (2 * batch, w['dz']))
        xh[...] = rng.standard_normal((2 * batch, w['dx']), dtype=np.float32)
        eh[...] = rng.standard_normal((2 * batch, w['dz']), dtype=np.float32)
        f.train_loop(xh, opt, batch, eps_host=eh, n_steps=2, exchange=peer)
        grp.barrier()
        t0 = time.perf_counter()
        f.train_loop(xh, opt, batch, eps_host=eh, n_steps=n_e2e, exchange=peer)
        e2e_s = grp.max(time.perf_counter() - t0) / n_e2e
        e2e = {'value': global_batch / e2e_s, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3,
               'h2d_bytes_per_step': int(batch * (w['dx'] + w['dz']) * 4), 'd2h_bytes_per_step': 16,
               'api': 'FusedELBO.train_loop (VAE.fit inner loop): per-step H2D of the shard from pinned memory, step, '
                      'scalars read-back, pipelined'}
    timed_out = False
    if peer is not None:
        timed_out = grp.sum(1.0 if peer.timed_out() else 0.0) > 0.0
        peer.close()
    tc_bad = grp.sum(1.0 if f.tc_status() else 0.0) > 0.0
    flop = 259200 if w['prior'] != 'normal' else 28800
    # bf16 MMA work the coupling-block kernels issue per 64-row tile and block (flow_tc.cu: six MMAs per float32 product):
    # forward 7 k-steps of 64 x 96 x 16, backward the same recompute + 6 k-steps of 64 x 112 x 16 + 4 of 128 x 96 x 16
    mma = lambda m, n, k: 2.0 * m * n * k
    per_tile = 6 * (2 * 7 * mma(64, 96, 16) + 6 * mma(64, 112, 16) + 4 * mma(128, 96, 16))
    mma_flop = per_tile * ((batch + 63) // 64) * w['num_blocks'] if f.path(batch) == 'tensor-core' else 0.0
    tfl = batch * flop / (ms * 1e-3) / 1e12  # per GPU, algorithmic (float32-equivalent)
    bf16_peak = peaks['bf16_tflops_sustained']
    roof = {'bound': 'tensor', 'kernel': 'flow_tc_kernel<fwd|bwd> (8 launches per step, ~80 % of it)', 'unit': 'TFLOP/s',
            'achieved': tfl, 'peak': bf16_peak, 'frac': tfl / bf16_peak, 'traffic': None,
            'peak_kind': 'measured cuBLAS bf16, sustained (MEASURED_PEAKS.json): the kernels run inside a long step',
            'issued_bf16_mma_tflops_over_whole_step': mma_flop / (ms * 1e-3) / 1e12,
            'issued_frac_of_peak_over_whole_step': mma_flop / (ms * 1e-3) / 1e12 / bf16_peak,
            'note': 'achieved = algorithmic float32 FLOPs (259,200 per configuration) per GPU over the WHOLE step; issued = the '
                    'bf16 MMAs actually launched (3 x BF16 split: six per float32 product, padded tiles)'}
    return {'workload': 'C5: same model, GLOBAL batch %d over %d GPU(s) (%d rows per GPU), fwd + bwd + gradient exchange '
                        '+ Adam' % (global_batch, world, batch),
            'scaling': 'strong', 'ms_per_step': ms, 'plan': f.path(batch), 'value': global_batch / (ms * 1e-3),
            'unit': UNIT, 'configs_per_s': global_batch / (ms * 1e-3),
            'collective': 'none' if world == 1 else ('peer kernel' if peer is not None else 'nccl'),
            'tflops_fp32': global_batch * flop / (ms * 1e-3) / 1e12, 'roofline': roof, 'e2e': e2e,
            'valid': not (timed_out or tc_bad), 'last_loss': float(f.scalars.numpy()[0])}


# ---------------------------------------------------------------------------------------------------- ELBO legs (C2, C1)
def elbo_leg(v, w, grp, batch, K, W, collective, sampler=None):
    """One ELBO training workload: `value` (inputs resident in HBM, per-step CUDA events, L2 flushed between steps) and
    `e2e` (the public training loop from pinned host memory, per-step H2D of x / eps and read-back of the loss)."""
    from vaemolsim_b200 import parallel
    c = v._abi.ctx()
    lib = c.lib
    rank, world = grp.rank, grp.world
    model = build_model(v, w, batch)
    f = model.fused(batch)
    opt = model.optimizer
    # synthetic inputs keyed by the GLOBAL row index of this rank's shard (results independent of the rank count)
    rng = parallel.global_row_seed(1001, rank * batch)
    n_sets = 4  # rotate a few input sets so consecutive steps do not see identical data
    xh = pinned_array(lib, (n_sets * batch, w['dx']))
    eh = pinned_array(lib, (n_sets * batch, w['dz']))
    xh[...] = rng.standard_normal(xh.shape, dtype=np.float32)
    eh[...] = rng.standard_normal(eh.shape, dtype=np.float32)
    xs = [v.Tensor.from_numpy(xh[k * batch:(k + 1) * batch]) for k in range(n_sets)]
    es = [v.Tensor.from_numpy(eh[k * batch:(k + 1) * batch]) for k in range(n_sets)]
    flush = v.Tensor((64 << 20, ))  # 256 MiB float32 > 126 MB L2
    peer = None
    note = None
    if world > 1:
        if collective in ('auto', 'peer'):
            try:
                peer = parallel.PeerExchange(grp, f.n_params)
            except Exception as e:  # no P2P mapping between the ranks' GPUs: NCCL path
                if collective == 'peer':
                    raise
                sys.stderr.write('bench: peer exchange unavailable (%s); using NCCL\n' % e)
        if grp.sum(1.0 if peer is not None else 0.0) != world:  # all ranks must agree on the path
            peer = None

    def step(i):
        if world == 1:
            f.train_step(xs[i % n_sets], es[i % n_sets], opt)  # forward + backward + Adam: 2 launches
        elif peer is not None:
            # gradient straight into this step's slot of the exchange buffer, then ONE kernel: flags + peer reads over
            # NVLink + rank-ordered sum + Adam
            peer.train_step(f, xs[i % n_sets], es[i % n_sets], batch, opt)
        else:
            f.forward_backward(xs[i % n_sets], es[i % n_sets])
            grp.allreduce_sum_device_(f.grad.ptr, f.n_params, c.stream)  # fallback: NCCL allreduce on the library's stream
            f.adam_step(opt, grad_scale=1.0 / world)

    for i in range(max(W, 3)):
        step(i)
    c.synchronize()
    if peer is not None and grp.sum(1.0 if peer.timed_out() else 0.0) > 0.0:
        note = 'peer exchange timed out during warm-up; NCCL fallback used'
        sys.stderr.write('bench: %s\n' % note)
        peer.close()
        peer = None
    # ---- timed region 1
    ev = Events(c, K)
    # kernel time for the roofline record: live, inside the timed region, on every 8th step -- the two extra event records
    # around the kernel cost a step ~5 us (0.111 -> 0.1055 ms), which would otherwise sit in `value` itself
    KT_EVERY = 8
    lib.vms_elbo_plan_set_timing_every(f.handle, K, KT_EVERY)
    grp.barrier()
    c.synchronize()
    launches0 = v._abi.launch_count()
    t_clock0 = time.perf_counter()
    for i in range(K):
        lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
        ev.record(2 * i)
        step(i)
        ev.record(2 * i + 1)
    c.synchronize()
    grp.barrier()
    wall = time.perf_counter() - t_clock0
    launches = v._abi.launch_count() - launches0
    dev_ms = grp.max(sum(ev.elapsed_ms(2 * i, 2 * i + 1) for i in range(K)))
    k_ms, k_n = C.c_double(0), C.c_int(0)
    lib.vms_elbo_plan_kernel_ms(f.handle, C.byref(k_ms), C.byref(k_n))
    lib.vms_elbo_plan_set_timing(f.handle, 0)
    # ---- timed region 2: end to end through the public training loop (`VAE.fit` -> FusedELBO.train_loop): every step's x
    # and eps leave PINNED host memory on a copy stream while the previous step trains, every step's {loss, nll, kl} is
    # read back (ring of 8 steps); N > 1: the same loop with the peer exchange inside
    exchange = peer
    if world > 1 and peer is None:
        e2e_api = 'FusedELBO.forward_backward + NCCL allreduce + Adam per step (synchronous), H2D / D2H per step'
        def e2e_step(i):
            xd, ed = v.Tensor.from_numpy(xh[(i % n_sets) * batch:(i % n_sets + 1) * batch]), \
                v.Tensor.from_numpy(eh[(i % n_sets) * batch:(i % n_sets + 1) * batch])
            f.forward_backward(xd, ed)
            grp.allreduce_sum_device_(f.grad.ptr, f.n_params, c.stream)
            f.adam_step(opt, grad_scale=1.0 / world)
            return f.scalars.numpy()
        for i in range(3):
            e2e_step(i)
        grp.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            last = e2e_step(i)
        e2e_s = grp.max(time.perf_counter() - t0)
    else:
        e2e_api = ('VAE.fit inner loop (FusedELBO.train_loop): per-step H2D of x / eps, step%s, loss read-back, pipelined' %
                   ('' if world == 1 else ' + peer-memory gradient exchange'))
        f.train_loop(xh, opt, batch, eps_host=eh, n_steps=8, exchange=exchange)
        c.synchronize()
        grp.barrier()
        t0 = time.perf_counter()
        scal = f.train_loop(xh, opt, batch, eps_host=eh, n_steps=K, exchange=exchange)
        c.synchronize()
        e2e_s = grp.max(time.perf_counter() - t0)
        last = scal[-1]
    if sampler is not None:
        # keep the same load going until the clock sampler has a few samples inside a loaded window (rank 0 decides for
        # all ranks: step() contains the gradient exchange when N > 1, so every rank must run the same number of steps)
        t_load = time.perf_counter()
        while grp.max(1.0 if rank == 0 and sampler.count() < 8 and time.perf_counter() - t_load < 3.0 else 0.0) > 0.0:
            for i in range(50):
                step(i)
            c.synchronize()
    replicas_ok = None
    if world > 1:
        # replicas must hold bit-identical parameters after the run (same reduced gradient on every rank)
        digest = float(np.frombuffer(f.theta.numpy().tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 40))
        replicas_ok = grp.max(digest) == -grp.max(-digest)
        if peer is not None:
            replicas_ok = replicas_ok and grp.sum(1.0 if peer.timed_out() else 0.0) == 0.0
            peer.close()
    return dict(model=model, f=f, opt=opt, ms_per_step=dev_ms / K, value=world * batch * K / (dev_ms * 1e-3),
                e2e_value=world * batch * K / e2e_s, e2e_ms=e2e_s / K * 1e3, e2e_api=e2e_api, last_loss=float(last[0]),
                launches=int(launches), kernel_ms=k_ms.value, kernel_launches=k_n.value, wall=wall, t0=t_clock0,
                t1=time.perf_counter(), path=f.path(batch), h2d=int(batch * (w['dx'] + w['dz']) * 4),
                collective=('none' if world == 1 else
                            'one fused kernel per step: NVLink peer-memory gradient allreduce + Adam (csrc/peer.cu)'
                            if exchange is not None else 'one NCCL allreduce(sum) of the flat gradient per step'),
                replicas_ok=replicas_ok, note=note, tc_bad=bool(f.tc_status()))


def tcf_issued_mma_flop(w, batch, rows=32):
    """bf16 MMA FLOPs the whole-step tensor-core kernel issues per launch (elbo_tcf.cu: six MMAs per float32 product; per
    32-row tile and block: forward 7 k-steps of 64 x 96 x 16, backward the same recompute + 6 k-steps of 64 x 112 x 16 +
    rows / 16 k-steps of 128 x 96 x 16)."""
    mma = lambda m, n, k: 2.0 * m * n * k
    per_tile = 6 * (2 * 7 * mma(64, 96, 16) + 6 * mma(64, 112, 16) + (rows // 16) * mma(128, 96, 16))
    return per_tile * ((batch + rows - 1) // rows) * w['num_blocks']


def run_b200(args, w):
    import vaemolsim_b200 as v
    from vaemolsim_b200 import parallel
    grp = parallel.Group()
    rank, world = grp.rank, grp.world
    c = v._abi.ctx()
    if args.backmap_only:
        line = backmap_leg(v, grp, measured_peaks(v)['fp32_ffma_tflops'])
        if rank == 0:
            print(json.dumps(line), flush=True)
        grp.close()
        return
    if args.mc_only or args.c4b_only:
        line = (c4b_bench if args.c4b_only else mc_bench)(v, grp, measured_peaks(v)['fp32_ffma_tflops'])
        if rank == 0:
            print(json.dumps(line), flush=True)
        grp.close()
        return
    batch = args.batch or w['batch']
    K, W = args.steps, args.warmup
    sampler = ClockSampler(c.device)
    if rank == 0:  # one nvidia-smi loop per job (rank 0's GPU), not one per rank
        sampler.start()
    r = elbo_leg(v, w, grp, batch, K, W, args.collective, sampler)
    clocks = sampler.stop(r['t0'], r['t1'])
    f = r['f']
    legs, micro, peaks = {}, None, None
    if not args.no_extras:
        peaks = measured_peaks(v)
        try:
            legs['c4a_mc'] = mc_bench(v, grp, peaks['fp32_ffma_tflops'])
        except Exception as ex:  # the headline line must survive a failure of an extra leg
            legs['c4a_mc'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
        try:
            legs['c4b_mc'] = c4b_bench(v, grp, peaks['fp32_ffma_tflops'])
        except Exception as ex:
            legs['c4b_mc'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
        try:
            legs['flow_model'] = flow_model_leg(v, grp)
        except Exception as ex:
            legs['flow_model'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
        try:
            legs['c5'] = large_batch_leg(v, w, r['opt'], grp, peaks, args.collective)
        except Exception as ex:
            legs['c5'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
        try:
            legs['c3'] = c3_leg(v, grp, peaks['hbm_gbs'])
        except Exception as ex:
            legs['c3'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
        try:
            legs['backmap'] = backmap_leg(v, grp, peaks['fp32_ffma_tflops'])
        except Exception as ex:
            legs['backmap'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
        try:
            if args.workload != 'c1':
                r1 = elbo_leg(v, WORKLOADS['c1'], grp, WORKLOADS['c1']['batch'], min(K, 100), 5, args.collective)
                k1 = r1['kernel_ms'] / max(r1['kernel_launches'], 1)
                tfl1 = 28800 * WORKLOADS['c1']['batch'] / (k1 * 1e-3) / 1e12 if k1 else None
                legs['c1'] = {'metric': METRIC, 'value': r1['value'], 'unit': UNIT, 'workload': WORKLOADS['c1']['label'],
                              'ms_per_step': r1['ms_per_step'], 'scaling': 'weak', 'plan': r1['path'],
                              'e2e': {'value': r1['e2e_value'], 'unit': UNIT, 'ms_per_step': r1['e2e_ms'],
                                      'h2d_bytes_per_step': r1['h2d'], 'd2h_bytes_per_step': 16, 'api': r1['e2e_api']},
                              'roofline': {'bound': 'ffma', 'kernel': 'elbo_fused_kernel<bwd>', 'achieved': tfl1,
                                           'peak': peaks['fp32_ffma_tflops'], 'unit': 'TFLOP/s',
                                           'frac': tfl1 / peaks['fp32_ffma_tflops'] if tfl1 else None, 'traffic': None,
                                           'launch_ms': k1, 'algorithmic_flop_per_launch': 28800 * WORKLOADS['c1']['batch']}}
                if rank == 0:
                    cv, cs = time_cpu(WORKLOADS['c1'], WORKLOADS['c1']['batch'], 10, 2)
                    legs['c1']['cpu_baseline'] = {'value': cv, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
                                                  'sample': '10 steps of 4096 configs, NumPy oracle', 'ms_per_step': cs * 1e3}
        except Exception as ex:
            legs['c1'] = {'error': '%s: %s' % (type(ex).__name__, ex)}
    if rank != 0:
        grp.close()
        return
    line = {
        'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': w['label'], 'global_batch': world * batch, 'params': f.n_params, 'plan': r['path'],
                   'parallelism': 'dp%d' % world if world > 1 else 'single', 'collective': r['collective'],
                   'l2': 'flushed between timed steps (256 MiB memset, outside the per-step events)',
                   'timing': 'sum of per-step CUDA-event intervals on the launching stream, max over ranks',
                   'wall_s_timed_region_incl_flush': r['wall']},
        'e2e': {'value': r['e2e_value'], 'unit': UNIT, 'h2d_bytes_per_step': r['h2d'], 'd2h_bytes_per_step': 16,
                'ms_per_step': r['e2e_ms'], 'api': r['e2e_api'],
                'note': 'back-to-back steps (L2 stays warm across steps), whereas `value` flushes L2 before every timed step: '
                        'e2e can therefore exceed `value`; every step still pays its own H2D (x, eps) and loss read-back',
                'last_loss': r['last_loss']},
        'gpu_launches': r['launches'],
        'clocks': clocks,
    }
    if r['replicas_ok'] is not None:
        line['config']['replicas_bit_identical'] = bool(r['replicas_ok'])
    if r['note']:
        line['config']['note'] = r['note']
    if r['tc_bad']:
        line['config']['tensor_core_wait_timeout'] = True
    if not args.no_extras:
        flop_per_config = 259200 if w['prior'] != 'normal' else 28800  # GEMM FLOPs fwd + bwd per configuration
        traffic = load_json(os.path.join('profiles', 'r02_traffic.json'))
        if r['kernel_launches']:
            launch_ms = r['kernel_ms'] / r['kernel_launches']
            tflops = flop_per_config * batch / (launch_ms * 1e-3) / 1e12
            tc = r['path'] == 'tensor-core-fused'
            kname = 'tcf_kernel<96, true> (whole-step tcgen05 kernel, elbo_tcf.cu)' if tc else 'elbo_fused_kernel<bwd>'
            peak = peaks['bf16_tflops'] if tc else peaks['fp32_ffma_tflops']
            roof = {'bound': 'tensor' if tc else 'ffma', 'kernel': kname, 'achieved': tflops, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': tflops / peak,
                    'traffic': (traffic.get('tcf_kernel' if tc else 'elbo_fused_kernel') or {}).get('dram_bytes_per_launch'),
                    'traffic_note': (traffic.get('tcf_kernel' if tc else 'elbo_fused_kernel') or {}).get('note'),
                    'peak_kind': ('measured cuBLAS bf16 dense, burst (MEASURED_PEAKS.json)' if tc else
                                  'FP32 FFMA issue peak measured in this run (csrc/probe.cu)'),
                    'algorithmic_flop_per_launch': flop_per_config * batch, 'launch_ms': launch_ms,
                    'launches_timed': r['kernel_launches'],
                    'launches_timed_note': 'every 8th step of the timed region carries the two event records around the '
                                           'kernel (they cost a step ~5 us)',
                    'share_of_step': launch_ms / r['ms_per_step'],
                    'fp32_ffma_peak_measured_tflops': peaks['fp32_ffma_tflops'],
                    'frac_of_fp32_ffma_measured': tflops / peaks['fp32_ffma_tflops']}
            if tc:
                issued = tcf_issued_mma_flop(w, batch) / (launch_ms * 1e-3) / 1e12
                roof.update({'issued_bf16_mma_tflops': issued,
                             'issued_frac_of_bf16_peak': issued / peaks['bf16_tflops'],
                             'issued_frac_of_tcgen05_m64_n96_issue_peak': issued / peaks['tcgen05_bf16_m64_n96_tflops'],
                             'note': 'achieved = ALGORITHMIC float32 FLOPs (259,200 per configuration) over the kernel time; the '
                                     'kernel issues six bf16 MMAs per float32 product on 64-row tiles of which 32 rows are valid '
                                     '(issued_*); batch 4096 = 128 tiles of a strictly sequential chain, so the kernel is bounded '
                                     'by phase latency, not by the tensor pipe (ncu: profiles/r02_*)'})
            line['roofline'] = roof
        micro = kernel_microbench(v, w, batch, peaks['hbm_gbs'])
        line['peaks'] = peaks
        line['roofline_hbm_kernels'] = {
            'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'peak_kind': peaks['kind'],
            'rqs_forward@stream': micro['rqs_forward@stream']['frac'],
            'rqs_inverse@stream': micro['rqs_inverse@stream']['frac'],
            'rqs_backward@stream': micro['rqs_backward@stream']['frac'],
            'normal_log_prob@stream': micro['normal_log_prob@stream']['frac'],
            'dist_select@C3': (legs.get('c3', {}).get('roofline') or {}).get('frac')}
        line['kernels'] = micro
        if 'error' not in legs.get('c4a_mc', {'error': 1}):
            legs['c4a_mc']['cpu_baseline'] = mc_cpu_baseline()
        line['legs'] = legs
        # short per-leg summary (kept flat so that per-N records retain every leg's number)
        line['leg_values'] = {k: {'value': d.get('value'), 'unit': d.get('unit'), 'e2e': (d.get('e2e') or {}).get('value'),
                                  'roofline_frac': (d.get('roofline') or {}).get('frac'),
                                  'weak_value': (d.get('weak') or {}).get('value')}
                              for k, d in legs.items() if 'error' not in d}
        line['mc'] = legs.get('c4a_mc')
        line['large_batch'] = legs.get('c5')
        line['cpu_baseline'] = cpu_baseline_subprocess(args, w, batch)
    print(json.dumps(line), flush=True)
    grp.close()


def main():
    # the contract is ONE JSON line on stdout: keep the real stdout for it and send every library's chatter (NCCL's
    # version banner, torch warnings) to stderr
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, 'w')
    args = parse()
    w = WORKLOADS[args.workload]
    if os.environ.get('VMS_BENCH_FAULT_AFTER'):  # development aid: dump every thread's Python stack and exit if stuck
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['VMS_BENCH_FAULT_AFTER']), exit=True)
    if args.impl == 'reference':
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == '__main__':
    main()
