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
  roofline     the dominant kernel of the step (elbo_fused_kernel), its device time measured with CUDA events around
               every launch INSIDE the timed region (vms_elbo_plan_set_timing); HBM-bound kernels timed alone are under
               "kernels"
  mc           the second headline metric: MC proposals/sec (C4a: 65,536 chains over all GPUs, 100 steps, fused kernel)
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


def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = args.batch or w['batch']
    # bounded sample: probe one full batch, then size the per-step sample so the whole run stays within ~2 minutes
    _, t_full = time_cpu(w, batch, 1, 1)
    budget = 120.0
    rows = batch
    total = (args.steps + args.warmup) * t_full
    if total > budget:
        rows = max(32, int(batch * budget / total) // 32 * 32)
    value, sec = time_cpu(w, rows, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': w['label'], 'rows_per_step': rows,
                   'note': 'CPU NumPy restatement of the reference path (oracle/); TF<=2.15 / TFP<=0.23 are not '
                           'installable in this image (Python 3.12, no network)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
                         'sample': '%d steps of %d configs (of the %d-config batch)' % (args.steps, rows, batch)},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
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


def kernel_microbench(v, w, batch, reps=20):
    """CUDA-event timings of the HBM-bound kernels, alone: at the workload's shape (cold L2) and at a streaming size."""
    c = v._abi.ctx()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_kind = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    K = w['num_bins']
    flush = v.Tensor((64 << 20, ))  # 256 MiB > 126 MB L2
    ev = Events(c, 1)
    rng = np.random.default_rng(5)
    res = {}

    def timed(fn, n_rep, do_flush):
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

    for tag, n in (('workload', batch), ('stream', 1 << 21)):
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
            ('rqs_forward', fwd_bytes, lambda: c.lib.vms_rqs_forward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0,
                                                                     y.ptr, l.ptr, c.stream)),
            ('rqs_inverse', fwd_bytes, lambda: c.lib.vms_rqs_inverse(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0,
                                                                     y.ptr, l.ptr, c.stream)),
            ('rqs_backward', bwd_bytes, lambda: c.lib.vms_rqs_backward(x.ptr, rw.ptr, rh.ptr, rs.ptr, n, K, -10.0, 10.0,
                                                                       1, g.ptr, g.ptr, gi.ptr, gw.ptr, gh.ptr, gs.ptr,
                                                                       c.stream)),
        ):
            ms = timed(fn, reps, True)
            gbs = nbytes / (ms * 1e-3) / 1e9
            res['%s@%s' % (name, tag)] = {'n_elem': n, 'ms': ms, 'bytes': nbytes, 'gbs': gbs, 'frac': gbs / peak}
    # decoder-distribution log_prob (K4) and neighbour selection (K6) at streaming sizes
    n, D = 1 << 22, w['dx']
    xx = v.Tensor.from_numpy(rng.standard_normal((n, D), dtype=np.float32))
    pp = v.Tensor.from_numpy(rng.standard_normal((n, 2 * D), dtype=np.float32))
    lp = v.Tensor((n, ))
    i32 = lambda a: (C.c_int32 * len(a))(*a)
    kind, loc, loc2, sc = i32([0] * D), i32(list(range(D))), i32([-1] * D), i32(list(range(D, 2 * D)))
    ms = timed(lambda: c.lib.vms_blockwise_log_prob(xx.ptr, D, pp.ptr, 2 * D, n, D, kind, loc, loc2, sc, 1, lp.ptr, 0,
                                                    c.stream), reps, True)
    nbytes = n * (3 * D * 4 + 4)
    res['normal_log_prob@stream'] = {'n_rows': n, 'ms': ms, 'bytes': nbytes, 'gbs': nbytes / ms / 1e6,
                                     'frac': nbytes / ms / 1e6 / peak}
    Bs, N, k = 4096, 10000, 50  # C3 shape: 4096 reference rows x 10,000 particles (coords replicated per row, as the API requires)
    coords = v.Tensor.from_numpy(np.broadcast_to(rng.uniform(-23.2, 23.2, (1, N, 3)).astype(np.float32), (Bs, N, 3)))
    ref = v.Tensor.from_numpy(rng.uniform(-23.2, 23.2, (Bs, 3)).astype(np.float32))
    box = v.Tensor.from_numpy(np.full(3, 46.416, np.float32))
    oxyz = v.Tensor((Bs, k, 3))
    ms = timed(lambda: c.lib.vms_dist_select(coords.ptr, None, Bs, N, ref.ptr, box.ptr, 0, 9.0, k, None, 0, oxyz.ptr,
                                             None, None, c.stream), reps, True)
    nbytes = Bs * (12 * N + 12 + k * 12)
    res['dist_select@C3'] = {'rows': Bs, 'N': N, 'k': k, 'ms': ms, 'bytes': nbytes, 'gbs': nbytes / ms / 1e6,
                                  'frac': nbytes / ms / 1e6 / peak}
    return res, peak, peak_kind


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


def mc_bench(v, grp, reps=3):
    """MC proposals/sec: chains sharded over ranks with no collective; the accept uniforms are ONE PCG64
    stream sliced per rank, so decisions do not depend on the rank count."""
    from vaemolsim_b200 import parallel
    c = v._abi.ctx()
    lo, hi = parallel.shard_rows(MC_CHAINS, grp.rank, grp.world)
    B = hi - lo
    model = build_model(v, WORKLOADS['c1'], 4096)
    mc = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002)
    if mc._fused_plan() is None:
        raise RuntimeError('bench: the fused MC plan is unavailable')
    x0 = np.random.default_rng(4001).standard_normal((MC_CHAINS, 6), dtype=np.float32)[lo:hi]
    t_rng = time.perf_counter()
    log_u = np.ascontiguousarray(np.log(np.random.default_rng(4002).random(size=(MC_STEPS, MC_CHAINS)))[:, lo:hi])
    t_rng = time.perf_counter() - t_rng
    # device-resident leg: chain state and the uniforms live in HBM; one launch = MC_STEPS steps of B chains
    xd, lud = v.Tensor.from_numpy(np.ascontiguousarray(x0)), v.Tensor.from_numpy(log_u)
    xd, ed = mc.run_fused(None, n_steps=MC_STEPS, configs_dev=xd, log_u_dev=lud)  # warm-up (also computes E)
    for _ in range(2):
        mc.run_fused(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed, log_u_dev=lud)
    ev = Events(c, reps)
    grp.barrier()
    c.synchronize()
    l0 = v._abi.launch_count()
    for i in range(reps):
        ev.record(2 * i)
        mc.run_fused(None, n_steps=MC_STEPS, configs_dev=xd, energies_dev=ed, log_u_dev=lud)
        ev.record(2 * i + 1)
    c.synchronize()
    grp.barrier()
    launches = v._abi.launch_count() - l0
    dev_ms = grp.max(sum(ev.elapsed_ms(2 * i, 2 * i + 1) for i in range(reps)))
    mc.sync_counters()
    acc_rate = mc.acceptance_rate
    # end to end through the public API: MCMC.run(configs, n_steps) from host arrays -- host PCG64 + log for the
    # uniforms (mcmc.py:119), H2D of x / log u, the launch, D2H of x / E
    mc2 = v.mcmc.MCMC(model, v.mcmc.QuadraticEnergy(6), random_seed=4002)
    mc2.run(x0, n_steps=2)
    grp.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        xe, ee = mc2.run(x0, n_steps=MC_STEPS)
    e2e_s = grp.max(time.perf_counter() - t0)
    tot = MC_CHAINS * MC_STEPS * reps
    return {'metric': 'MC proposals/sec', 'value': tot / (dev_ms * 1e-3), 'unit': 'proposals/s', 'workload': MC_LABEL,
            'chains_global': MC_CHAINS, 'chains_per_gpu': B, 'steps_per_launch': MC_STEPS, 'launches_timed': reps,
            'scaling': 'strong (65,536 chains split over the GPUs, no data-path collective)',
            'ms_per_mc_step': dev_ms / (reps * MC_STEPS), 'gpu_launches': int(launches), 'acceptance_rate': acc_rate,
            'e2e': {'value': tot / e2e_s, 'unit': 'proposals/s', 'h2d_bytes_per_step': int(B * 8 + B * 24 / MC_STEPS),
                    'd2h_bytes_per_step': int(B * 32 / MC_STEPS), 'ms_per_mc_step': e2e_s / (reps * MC_STEPS) * 1e3,
                    'api': 'MCMC.run(configs, n_steps=%d) from host arrays' % MC_STEPS,
                    'note': 'includes NumPy PCG64 + log of the accept uniforms on the host (%.1f ms per %d x %d block '
                            'on this box): the stream is kept on the host so decisions stay bit-identical to the '
                            'reference under the same seed' % (t_rng * 1e3, MC_STEPS, B)},
            'algorithmic_flop_per_proposal': 19200,
            'tflops_fp32': tot * 19200 / (dev_ms * 1e-3) / 1e12}



def large_batch_leg(v, w, opt, grp, collective='auto', global_batch=262144, steps=5):
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
    peer = gt = None
    if world > 1:
        if collective in ('auto', 'peer'):
            try:
                peer = parallel.PeerExchange(grp, f.n_params)
            except Exception as ex:
                sys.stderr.write('bench (C5 leg): peer exchange unavailable (%s); using NCCL\n' % ex)
        if peer is None:
            gt, _ = grp.wrap_device_buffer(f.grad.ptr, f.n_params, c.stream)

    def step():
        if world == 1:
            f.train_step(x, e, opt)
        elif peer is not None:
            f.forward_backward(x, e, grad_ptr=peer.next_slot())
            peer.allreduce_adam(f, opt)
        else:
            f.forward_backward(x, e)
            grp.allreduce_sum_(gt, host_sync=c.synchronize)
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
    mma_flop = per_tile * ((batch + 63) // 64) * w['num_blocks'] * world if f.path(batch) == 'tensor-core' else 0.0
    try:
        bf16_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('bf16_tflops', 1590.0))
    except (OSError, ValueError):
        bf16_peak = 1590.0
    tensor = {'bound': 'tensor', 'kernel': 'flow_tc_kernel<fwd|bwd> (8 launches per step)', 'unit': 'TFLOP/s',
              'bf16_mma_flop_per_step': mma_flop, 'achieved_over_whole_step': mma_flop / (ms * 1e-3) / 1e12 / world,
              'peak': bf16_peak, 'frac_over_whole_step': mma_flop / (ms * 1e-3) / 1e12 / world / bf16_peak,
              'note': 'MMA FLOPs actually issued (3 x BF16 split: six bf16 MMAs per float32 product) divided by the WHOLE '
                      'step time per GPU -- a lower bound of the kernels\' own rate (they are ~80 % of the step); ncu: '
                      'tensor pipe active 15-18 % (profiles/r01_ncu_flow_tc_pipelined.txt)'}
    return {'workload': 'C5: same model, GLOBAL batch %d over %d GPU(s) (%d rows per GPU), fwd + bwd + gradient exchange '
                        '+ Adam' % (global_batch, world, batch),
            'scaling': 'strong', 'ms_per_step': ms, 'plan': f.path(batch), 'configs_per_s': global_batch / (ms * 1e-3),
            'collective': 'none' if world == 1 else ('peer kernel' if peer is not None else 'nccl'),
            'tflops_fp32': global_batch * flop / (ms * 1e-3) / 1e12, 'roofline': tensor,
            'valid': not (timed_out or tc_bad),
            'last_loss': float(f.scalars.numpy()[0])}


def run_b200(args, w):
    import vaemolsim_b200 as v
    from vaemolsim_b200 import parallel
    grp = parallel.Group()
    rank, world = grp.rank, grp.world
    c = v._abi.ctx()
    lib = c.lib
    if args.mc_only:
        line = mc_bench(v, grp)
        if rank == 0:
            print(json.dumps(line), flush=True)
        grp.close()
        return
    batch = args.batch or w['batch']
    K, W = args.steps, args.warmup
    model = build_model(v, w, batch)
    f = model.fused(batch)
    opt = model.optimizer
    # synthetic inputs keyed by the GLOBAL row index of this rank's shard (results independent of the rank count)
    row0 = rank * batch
    rng = parallel.global_row_seed(1001, row0)
    n_sets = 4  # rotate a few input sets so consecutive steps do not see identical data
    xs_host = [pinned_array(lib, (batch, w['dx'])) for _ in range(n_sets)]
    es_host = [pinned_array(lib, (batch, w['dz'])) for _ in range(n_sets)]
    for a in xs_host + es_host:
        a[...] = rng.standard_normal(a.shape, dtype=np.float32)
    xs = [v.Tensor.from_numpy(a) for a in xs_host]
    es = [v.Tensor.from_numpy(a) for a in es_host]
    flush = v.Tensor((64 << 20, ))  # 256 MiB float32 > 126 MB L2
    gt = ext = peer = None
    if world > 1:
        if args.collective in ('auto', 'peer'):
            try:
                peer = parallel.PeerExchange(grp, f.n_params)
            except Exception as e:  # no P2P mapping between the ranks' GPUs: NCCL path
                if args.collective == 'peer':
                    raise
                sys.stderr.write('bench: peer exchange unavailable (%s); using NCCL\n' % e)
        # all ranks must agree on the path
        if grp.sum(1.0 if peer is not None else 0.0) != world:
            peer = None
        if peer is None:
            gt, ext = grp.wrap_device_buffer(f.grad.ptr, f.n_params, c.stream)

    def step(i):
        if world == 1:
            f.train_step(xs[i % n_sets], es[i % n_sets], opt)  # forward + backward + Adam: 2 launches
            return
        if peer is not None:
            # gradient straight into this step's slot of the exchange buffer, then ONE kernel: flags + peer reads over
            # NVLink + rank-ordered sum + Adam
            f.forward_backward(xs[i % n_sets], es[i % n_sets], grad_ptr=peer.next_slot())
            peer.allreduce_adam(f, opt)
            return
        f.forward_backward(xs[i % n_sets], es[i % n_sets])
        grp.allreduce_sum_(gt, host_sync=c.synchronize)  # fallback: NCCL allreduce, host-synchronised
        f.adam_step(opt, grad_scale=1.0 / world)

    sampler = ClockSampler(c.device)
    if rank == 0:  # one nvidia-smi loop per job (rank 0's GPU), not one per rank
        sampler.start()
    for i in range(max(W, 3)):
        step(i)
    c.synchronize()
    peer_note = None
    if peer is not None and grp.sum(1.0 if peer.timed_out() else 0.0) > 0.0:
        # a rank gave up waiting for a peer's flag during warm-up (the kernel's bound): do not spend the timed region on
        # 2 s time-outs -- finish on the NCCL exchange and say so
        peer_note = 'peer exchange timed out during warm-up; NCCL fallback used'
        sys.stderr.write('bench: %s\n' % peer_note)
        peer.close()
        peer = None
        gt, ext = grp.wrap_device_buffer(f.grad.ptr, f.n_params, c.stream)

    # ---- timed region 1: inputs resident in HBM, per-step CUDA events, L2 flushed (untimed) between steps
    ev = Events(c, K)
    lib.vms_elbo_plan_set_timing(f.handle, K)
    grp.barrier()
    c.synchronize()
    launches0 = v._abi.launch_count()
    t_clock0 = time.perf_counter()
    wall0 = time.perf_counter()
    for i in range(K):
        lib.vms_memset(flush.ptr, 0, flush.nbytes, c.stream)
        ev.record(2 * i)
        step(i)
        ev.record(2 * i + 1)
    c.synchronize()
    grp.barrier()
    wall1 = time.perf_counter()
    launches = v._abi.launch_count() - launches0
    dev_ms = sum(ev.elapsed_ms(2 * i, 2 * i + 1) for i in range(K))
    dev_ms = grp.max(dev_ms)
    k_ms, k_n = C.c_double(0), C.c_int(0)
    lib.vms_elbo_plan_kernel_ms(f.handle, C.byref(k_ms), C.byref(k_n))
    lib.vms_elbo_plan_set_timing(f.handle, 0)
    ms_per_step = dev_ms / K
    value = world * batch * K / (dev_ms * 1e-3)

    # ---- timed region 2: end to end through the public API from pinned host memory
    def e2e_step(i):
        xd = v.Tensor.from_numpy(xs_host[i % n_sets])  # H2D (pinned)
        ed = v.Tensor.from_numpy(es_host[i % n_sets])
        if world == 1:
            f.train_step(xd, ed, opt)
        elif peer is not None:
            f.forward_backward(xd, ed, grad_ptr=peer.next_slot())
            peer.allreduce_adam(f, opt)
        else:
            f.forward_backward(xd, ed)
            grp.allreduce_sum_(gt, host_sync=c.synchronize)
            f.adam_step(opt, grad_scale=1.0 / world)
        return f.scalars.numpy()  # D2H of {loss, nll, kl}: synchronises the step

    e2e_api = 'FusedELBO.train_step (the call behind VAE.train_step / VAE.fit)'
    if world == 1:
        # the public training loop (`VAE.fit` -> FusedELBO.train_loop): every step's x and eps leave PINNED host memory on
        # a copy stream while the previous step trains, every step's {loss, nll, kl} is read back (ring of 8 steps)
        e2e_api = 'VAE.fit inner loop (FusedELBO.train_loop): per-step H2D of x / eps, step, loss read-back, pipelined'
        xh = pinned_array(lib, (n_sets * batch, w['dx']))
        eh = pinned_array(lib, (n_sets * batch, w['dz']))
        for k in range(n_sets):
            xh[k * batch:(k + 1) * batch] = xs_host[k]
            eh[k * batch:(k + 1) * batch] = es_host[k]
        f.train_loop(xh, opt, batch, eps_host=eh, n_steps=8)
        c.synchronize()
        t0 = time.perf_counter()
        scal = f.train_loop(xh, opt, batch, eps_host=eh, n_steps=K)
        c.synchronize()
        e2e_s = time.perf_counter() - t0
        last = scal[-1]
    else:
        for i in range(3):
            e2e_step(i)
        grp.barrier()
        c.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            last = e2e_step(i)
        c.synchronize()
        e2e_s = grp.max(time.perf_counter() - t0)
    e2e_value = world * batch * K / e2e_s
    # keep the same load going until the clock sampler has a few samples inside a loaded window
    # (the decision is rank 0's, shared with every rank: step() contains the gradient exchange when N > 1, so all ranks
    # must run the same number of steps)
    t_load = time.perf_counter()
    while grp.max(1.0 if rank == 0 and sampler.count() < 8 and time.perf_counter() - t_load < 3.0 else 0.0) > 0.0:
        for i in range(50):
            step(i)
        c.synchronize()
    t_clock1 = time.perf_counter()
    clocks = sampler.stop(t_clock0, t_clock1)

    replicas_ok = None
    if world > 1:
        # replicas must hold bit-identical parameters after the run (same reduced gradient on every rank)
        digest = float(np.frombuffer(f.theta.numpy().tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 40))
        replicas_ok = grp.max(digest) == -grp.max(-digest)
        if peer is not None:
            replicas_ok = replicas_ok and grp.sum(1.0 if peer.timed_out() else 0.0) == 0.0
            peer.close()
    mc_line = None if args.no_extras else mc_bench(v, grp)
    lb_line = None
    if not args.no_extras:
        try:
            lb_line = large_batch_leg(v, w, opt, grp, args.collective)
        except Exception as ex:  # the headline line must survive a failure of this extra leg
            lb_line = {'error': '%s: %s' % (type(ex).__name__, ex)}
            sys.stderr.write('bench (C5 leg) failed: %s\n' % lb_line['error'])
    if rank != 0:
        grp.close()
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': w['label'], 'global_batch': world * batch, 'params': f.n_params,
                   'parallelism': 'dp%d' % world if world > 1 else 'single',
                   'collective': ('none' if world == 1 else
                                  'one fused kernel per step: NVLink peer-memory gradient allreduce + Adam (csrc/peer.cu)'
                                  if peer is not None else 'one NCCL allreduce(sum) of the flat gradient per step'),
                   'l2': 'flushed between timed steps (256 MiB memset, outside the per-step events)',
                   'timing': 'sum of per-step CUDA-event intervals on the launching stream, max over ranks',
                   'wall_s_timed_region_incl_flush': wall1 - wall0},
        'e2e': {'value': e2e_value, 'unit': UNIT,
                'h2d_bytes_per_step': int(xs_host[0].nbytes + es_host[0].nbytes), 'd2h_bytes_per_step': 16,
                'ms_per_step': e2e_s / K * 1e3, 'api': e2e_api,
                'note': 'back-to-back steps (L2 stays warm across steps), whereas `value` flushes L2 before every timed step: '
                        'e2e can therefore exceed `value`; every step still pays its own H2D (x, eps) and loss read-back',
                'last_loss': float(last[0])},
        'gpu_launches': int(launches),
        'clocks': clocks,
    }
    if replicas_ok is not None:
        line['config']['replicas_bit_identical'] = bool(replicas_ok)
    if peer_note:
        line['config']['note'] = peer_note
    if not args.no_extras:
        micro, peak, peak_kind = kernel_microbench(v, w, batch)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        flop_per_config = 259200 if w['prior'] != 'normal' else 28800  # GEMM FLOPs fwd + bwd per configuration
        if k_n.value:
            launch_ms = k_ms.value / k_n.value
            tflops = flop_per_config * batch / (launch_ms * 1e-3) / 1e12
            tens_peak = float(peaks.get('bf16_tflops', 1590.0))
            line['roofline'] = {
                'bound': 'tensor', 'kernel': 'elbo_fused_kernel<bwd>', 'achieved': tflops, 'peak': tens_peak,
                'unit': 'TFLOP/s', 'frac': tflops / tens_peak, 'traffic': None,
                'peak_kind': ('measured bf16 dense, burst (MEASURED_PEAKS.json)' if 'bf16_tflops' in peaks else
                              'fallback (B200_PROFILING.md)'),
                'algorithmic_flop_per_launch': flop_per_config * batch, 'launch_ms': launch_ms, 'launches_timed': k_n.value,
                'share_of_step': launch_ms / ms_per_step,
                'fp32_ffma_peak_nominal_tflops': 74.4, 'frac_of_fp32_ffma': tflops / 74.4,
                'note': 'the step is one FP32-FFMA kernel (float32 parity forbids plain TF32 tensor-core inputs, DESIGN.md 6); '
                        'it is bounded by per-phase latency at 32 rows per SM, not by a pipe: the math-pipe roofline is '
                        'reported against the measured bf16 peak as the contract asks and against the nominal FP32 FFMA '
                        'peak; HBM-bound kernels and their fractions of the measured copy peak are under "kernels"'}
        line['roofline_hbm_kernels'] = {
            'peak': peak, 'unit': 'GB/s', 'peak_kind': peak_kind,
            'rqs_forward@stream': micro['rqs_forward@stream']['frac'],
            'rqs_inverse@stream': micro['rqs_inverse@stream']['frac'],
            'rqs_backward@stream': micro['rqs_backward@stream']['frac'],
            'normal_log_prob@stream': micro['normal_log_prob@stream']['frac'],
            'dist_select@C3': micro['dist_select@C3']['frac']}
        line['kernels'] = micro
        line['large_batch'] = lb_line
        line['mc'] = mc_line
        line['mc']['cpu_baseline'] = mc_cpu_baseline()
        rows = batch
        cpu_val, cpu_sec = time_cpu(w, rows, 10, 2)
        line['cpu_baseline'] = {'value': cpu_val, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
                                'sample': '10 steps of %d configs, NumPy oracle (ELBO fwd + analytic bwd + Adam)' % rows,
                                'ms_per_step': cpu_sec * 1e3}
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
