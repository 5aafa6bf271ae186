"""VAE-based Monte Carlo -- host-side mirror of `vaemolsim/mcmc.py` with the accept / reject step on the device.

Same class, methods and counters as the reference (mcmc.py:12 `MCMC`, :48 `acceptance_rate`, :52 `reset`, :68
`single_step`, :133 `run`).  The six distribution evaluations of a step (mcmc.py:100-108) are the kernels behind
`vae.encoder / prior / decoder`; the acceptance arithmetic of mcmc.py:116-128 is `vms_mc_accept` (float64, same
operation order).  The accept uniforms are NumPy's PCG64 stream (mcmc.py:119) under the same seed: the op-by-op path draws
them on the host (8 bytes per chain per step uploaded); the fused path for the C4a shape regenerates the SAME stream on
the device (`vms_mc_run_pcg64`: per-chain LCG jump-ahead, bit-identical uniforms, CUDA double log) and falls back to the
host stream for a call whose decisions could depend on the last ulps of the logarithm, so decisions are those of the
reference either way.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import Tensor, as_tensor, ctx


class QuadraticEnergy(object):
    """The test energy of the reference (tests/test_mcmc.py:28-32): E(x) = sum_d (x_d - means_d)^2, means =
    linspace(-2, 2, D) unless given.  Callable on NumPy arrays exactly like the reference's callback; an `MCMC` built on
    a Gaussian VAE additionally recognises it and runs whole MC steps in the fused device kernel (`vms_mc_run`)."""

    def __init__(self, n_dims=None, means=None):
        if means is None:
            means = np.linspace(-2, 2, int(n_dims))
        self.means = np.asarray(means, dtype=np.float64).reshape(-1)

    def __call__(self, configs):
        return np.sum((np.asarray(configs) - self.means[np.newaxis, :])**2, axis=-1)


class GaussianMixtureEnergy(object):
    """The "energy" of examples/MC_Moves_with_VAEs.ipynb (cell 5, 38): the log-density of a mixture of independent Normals,
    `tfp.distributions.Mixture(Categorical(probs), [Independent(Normal(loc_k, scale_k))...]).log_prob(configs)`.  Callable on
    NumPy arrays like the reference's callback (float32 result, tfp's arithmetic), and flagged `on_device`: an `MCMC` calls
    it with device tensors and gets device float32 energies back (`vms_energy_gmm`, the dtype of `log_prob(...).numpy()`), so an MC loop has no PCIe round trip.
    Defaults: the notebook's three components."""
    on_device = True

    def __init__(self, probs=(0.7, 0.2, 0.1), locs=((-0.5, 0.0), (1.0, 2.0), (-1.5, 0.0)),
                 scales=((0.05, 0.5), (1.0, 0.5), (0.5, 0.2))):
        self.probs = np.asarray(probs, np.float32)
        self.locs = np.asarray(locs, np.float32).reshape(len(self.probs), -1)
        self.scales = np.asarray(scales, np.float32).reshape(len(self.probs), -1)
        self._dev = None

    def _host(self, configs):
        x = np.asarray(configs, np.float32)
        lp = np.stack([np.sum(-0.5 * (x / s - m / s)**2 - (np.float32(0.9189385332046727) + np.log(s)), axis=-1,
                              dtype=np.float32) + np.log(p)
                       for p, m, s in zip(self.probs, self.locs, self.scales)], axis=-1).astype(np.float32)
        mx = lp.max(axis=-1, keepdims=True)
        return (mx[..., 0] + np.log(np.sum(np.exp(lp - mx), axis=-1, dtype=np.float32))).astype(np.float32)

    def __call__(self, configs):
        if not isinstance(configs, Tensor):
            return self._host(configs)
        c = ctx()
        if self._dev is None:
            self._dev = (Tensor.from_numpy(np.log(self.probs)), Tensor.from_numpy(self.locs), Tensor.from_numpy(self.scales))
        x = configs.contig()
        E = Tensor((x.shape[0], ), np.float32)
        c.lib.vms_energy_gmm(x.ptr, x.shape[0], x.shape[1], len(self.probs), self._dev[0].ptr, self._dev[1].ptr,
                             self._dev[2].ptr, E.ptr, c.stream)
        return E


class MCMC(object):
    """Markov chain Monte Carlo with a VAE proposal: as many independent chains as input configurations."""

    def __init__(self, vae, energy_func, random_seed=None, stream_layout=None):
        """`stream_layout = (chain0, n_chains_global)` (extension, default (0, B)): this object's chains are rows
        [chain0, chain0 + B) of a global set of n_chains_global chains that share ONE uniform stream -- how a multi-GPU job
        keeps the decisions of the single-process run."""
        self.vae = vae
        self.energy_func = energy_func
        self.stream_layout = stream_layout
        self._fused = None
        self._nb = None
        self._acc_folded = 0   # value of the device acceptance counter already folded into _num_acc
        self.device_rng = True  # fused path: draw the accept uniforms on the device when the kernel supports it
        self.host_stream_reruns = 0
        self.reset(random_seed)

    @property
    def acceptance_rate(self):
        return self._num_acc / self._num_trials

    def reset(self, random_seed=None):
        self._num_trials = 0.0
        self._num_acc = 0.0
        self._rng = np.random.default_rng(seed=random_seed)
        self._noise_seed = int(np.random.SeedSequence(random_seed).generate_state(2, np.uint32).astype(np.uint64) @
                               np.array([1, 1 << 32], np.uint64))
        self._step0 = 0
        if self._fused:  # accepts counted on the device before the reset must not leak into the new statistics
            self._fused['n_acc'].fill_zero()
        if getattr(self, '_nb', None):
            self._nb['n_acc'].fill_zero()
            self._nb['n_unc'].fill_zero()
            self._nb['folded'] = 0
        self._acc_folded = 0

    # ------------------------------------------------------------------------------------------ fused device path
    def _fused_plan(self):
        """`vms_mc_plan` for this (vae, energy) pair, or None when the pair is outside the fused kernel's family
        (Gaussian VAE with N(0, I) prior + QuadraticEnergy); the op-by-op path below handles everything else."""
        if self._fused is not None:
            return self._fused or None
        self._fused = False
        if not isinstance(self.energy_func, QuadraticEnergy):
            return None
        try:
            f = self.vae.fused()
        except (NotImplementedError, RuntimeError, AttributeError):
            return None
        if f.desc.num_blocks != 0 or self.energy_func.means.size != f.dx:
            return None
        c = ctx()
        desc = _abi.McDesc(f.dx, f.dz, f.desc.hidden)
        h = C.c_void_p()
        try:
            c.lib.vms_mc_plan_create(C.byref(desc), C.byref(h))
        except NotImplementedError:
            return None
        self._fused = dict(handle=h.value, elbo=f, means=Tensor.from_numpy(self.energy_func.means),
                           n_acc=Tensor.zeros((1, ), np.uint64), n_unc=Tensor.zeros((1, ), np.uint64),
                           device_rng=bool(c.lib.vms_mc_plan_has_device_rng(h.value)))
        return self._fused

    def close(self):
        """Releases the fused MC plan and the pinned uniform buffers."""
        lib = _abi.load()
        if self._fused:
            lib.vms_mc_plan_destroy(self._fused['handle'])
        self._fused = None
        ub = getattr(self, '_ub', None)
        if ub is not None:
            ctx().synchronize()
            host, dev, evs, raw = ub
            for e in evs:
                lib.vms_event_destroy(e)
            for p in raw:
                lib.vms_free_host(p)
            self._ub, self._ub_key = None, None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the PCG64 stream as the device kernel consumes it
    _PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645
    _M128 = (1 << 128) - 1

    def _pcg_stream(self, chain0, n_global):
        """`vms_pcg64_stream` describing self._rng's CURRENT position: state / increment and the affine jump over n_global
        draws (one MC step of the global chain set), computed with Python integers (pcg_advance_lcg_128)."""
        st = self._rng.bit_generator.state
        if st.get('bit_generator') != 'PCG64':
            return None
        state, inc = int(st['state']['state']), int(st['state']['inc'])
        key = (inc, int(n_global))
        if getattr(self, '_jump_key', None) != key:
            am, ap, cm, cp, d, M = 1, 0, self._PCG_MULT, inc, int(n_global), self._M128
            while d > 0:
                if d & 1:
                    am, ap = (am * cm) & M, (ap * cm + cp) & M
                cp, cm, d = ((cm + 1) * cp) & M, (cm * cm) & M, d >> 1
            self._jump_key, self._jump = key, (am, ap)
        am, ap = self._jump
        lo64 = (1 << 64) - 1
        return _abi.Pcg64Stream(state >> 64, state & lo64, inc >> 64, inc & lo64, am >> 64, am & lo64, ap >> 64, ap & lo64,
                                int(chain0))

    def _fold(self):
        """Fold the device acceptance counter into `_num_acc` (one bookkeeping for the host- and device-resident paths)."""
        total = int(self._fused['n_acc'].numpy()[0])
        self._num_acc += float(total - self._acc_folded)
        self._acc_folded = total

    def run_fused(self, configs, energies=None, n_steps=1, noise=None, trace=False, configs_dev=None, energies_dev=None,
                  log_u_dev=None):
        """n_steps MC steps in one `vms_mc_run` launch.  `noise` [n_steps, B, 2 dz + dx] injects the sampling noise
        (parity tests); otherwise the device Philox stream keyed by (seed, chain, step) is used.  With `*_dev` tensors the
        chain state stays on the device (the bench's device-resident leg); log_u always comes from the host PCG64 stream."""
        fp = self._fused_plan()
        if fp is None:
            raise NotImplementedError('run_fused: needs a Gaussian VAE (N(0, I) prior) and a QuadraticEnergy')
        c = ctx()
        f = fp['elbo']
        if configs_dev is None:
            configs = np.array(configs.numpy() if isinstance(configs, Tensor) else configs)
            x = Tensor.from_numpy(configs.reshape(configs.shape[0], -1), dtype=np.float32)
        else:
            x = configs_dev
        B = x.shape[0]
        if energies_dev is not None:
            e, valid = energies_dev, 1
        elif energies is None:
            e, valid = Tensor((B, ), np.float64), 0
        else:
            e, valid = as_tensor(np.asarray(energies, np.float64), dtype=np.float64), 1
        nz = None if noise is None else Tensor.from_numpy(np.ascontiguousarray(noise, np.float32))
        P_ = lambda t: None if t is None else t.ptr
        tr = {}
        chain0, n_global = self.stream_layout if self.stream_layout is not None else (0, B)
        c.lib.vms_mc_plan_set_chain_offset(fp['handle'], int(chain0))
        stream = self._pcg_stream(chain0, n_global) if (self.device_rng and fp['device_rng'] and log_u_dev is None) else None
        done = False
        if stream is not None:
            # the accept uniforms are drawn ON THE DEVICE from this generator's PCG64 stream (bit-identical u, device log):
            # no host draw, no upload.  The chain state is saved first so the call can be repeated on the host stream in the
            # (~1e-6 per run) case that a decision sits within ~1e-13 of its threshold.
            x_save, e_save = Tensor(x.shape), Tensor((B, ), np.float64)
            c.lib.vms_memcpy_d2d(x_save.ptr, x.ptr, x.nbytes, c.stream)
            if valid:
                c.lib.vms_memcpy_d2d(e_save.ptr, e.ptr, e.nbytes, c.stream)
            acc_before = Tensor((1, ), np.uint64)
            c.lib.vms_memcpy_d2d(acc_before.ptr, fp['n_acc'].ptr, 8, c.stream)
            fp['n_unc'].fill_zero()
            if trace:
                tr = dict(acc=Tensor((n_steps, B), np.uint8), fwd=Tensor((n_steps, B)), rev=Tensor((n_steps, B)),
                          e_new=Tensor((n_steps, B), np.float64), log_u=Tensor((n_steps, B), np.float64))
            c.lib.vms_mc_run_pcg64(fp['handle'], f.theta.ptr, x.ptr, e.ptr, valid, P_(nz), self._noise_seed, self._step0,
                                   C.byref(stream), fp['means'].ptr, B, n_steps, fp['n_acc'].ptr, fp['n_unc'].ptr,
                                   P_(tr.get('acc')), P_(tr.get('fwd')), P_(tr.get('rev')), P_(tr.get('e_new')),
                                   P_(tr.get('log_u')), c.stream)
            if configs_dev is not None and not self.check_uncertain:
                done = True  # device-resident loops check `uncertain()` themselves (one read-back per loop, not per call)
            elif int(fp['n_unc'].numpy()[0]) == 0:
                done = True
            else:  # repeat this call on the NumPy stream: restore the chain state and the acceptance counter
                self.host_stream_reruns += 1
                c.lib.vms_memcpy_d2d(x.ptr, x_save.ptr, x.nbytes, c.stream)
                if valid:
                    c.lib.vms_memcpy_d2d(e.ptr, e_save.ptr, e.nbytes, c.stream)
                c.lib.vms_memcpy_d2d(fp['n_acc'].ptr, acc_before.ptr, 8, c.stream)
                tr = {}
            if done:
                self._rng.bit_generator.advance(n_steps * n_global)
        chunk = 10
        if done:
            pass
        elif log_u_dev is None and nz is None and not trace and n_steps > chunk:
            # pipeline: the host draws / logs the uniforms of the next `chunk` steps (mcmc.py:119; ONE sequential PCG64
            # stream) while the device runs the previous chunk's launch
            host, dev, evs, _ = self._uniform_buffers(chunk, B)
            for k, s0 in enumerate(range(0, n_steps, chunk)):
                ns, i = min(chunk, n_steps - s0), k & 1
                if k >= 2:
                    c.lib.vms_event_synchronize(evs[i])  # the upload that last used this pinned buffer has finished
                h = host[i][:ns]
                self._host_uniform_logs(h, chain0, n_global)
                c.lib.vms_memcpy_h2d(dev[i].ptr, h.ctypes.data, ns * B * 8, c.stream)  # pinned => truly asynchronous
                c.lib.vms_event_record(evs[i], c.stream)
                c.lib.vms_mc_run(fp['handle'], f.theta.ptr, x.ptr, e.ptr, valid, None, self._noise_seed,
                                 self._step0 + s0, dev[i].ptr, fp['means'].ptr, B, ns, fp['n_acc'].ptr, None, None, None,
                                 None, c.stream)
                valid = 1
            c.synchronize()
        else:
            if log_u_dev is None:
                h = np.empty((n_steps, B))
                self._host_uniform_logs(h, chain0, n_global)  # mcmc.py:119, step by step
                log_u_dev = Tensor.from_numpy(h)
            if trace:
                tr = dict(acc=Tensor((n_steps, B), np.uint8), fwd=Tensor((n_steps, B)), rev=Tensor((n_steps, B)),
                          e_new=Tensor((n_steps, B), np.float64), log_u=log_u_dev)
            c.lib.vms_mc_run(fp['handle'], f.theta.ptr, x.ptr, e.ptr, valid, P_(nz), self._noise_seed, self._step0,
                             log_u_dev.ptr, fp['means'].ptr, B, n_steps, fp['n_acc'].ptr, P_(tr.get('acc')),
                             P_(tr.get('fwd')), P_(tr.get('rev')), P_(tr.get('e_new')), c.stream)
        self._step0 += n_steps
        self._num_trials += B * n_steps
        if configs_dev is not None:
            return x, e
        self._fold()
        self._last_trace = {k: t.numpy() for k, t in tr.items()}
        return x.numpy().reshape(configs.shape), e.numpy()

    check_uncertain = False  # device-resident calls: read the uncertainty counter back after every call (tests)

    def uncertain(self):
        """Number of chain-steps of device-resident `run_fused` calls (since the last call of this method) whose decision
        was within ~1e-13 of its threshold under the device logarithm; non-zero means those calls must be repeated on the
        host stream (`device_rng = False`).  Host-array calls handle this themselves."""
        fp = self._fused_plan() or (self._nb or None)
        if fp is None:
            return 0
        n = int(fp['n_unc'].numpy()[0])
        fp['n_unc'].fill_zero()
        return n

    def _host_uniform_logs(self, out, chain0, n_global):
        """out [n_steps, B] <- log of this object's columns [chain0, chain0 + B) of the global [n_steps, n_global] block of
        uniforms (mcmc.py:119), drawn from self._rng in stream order."""
        B = out.shape[1]
        if chain0 == 0 and n_global == B:
            self._rng.random(out=out)
        else:
            out[...] = self._rng.random(size=(out.shape[0], n_global))[:, chain0:chain0 + B]
        np.log(out, out=out)

    def _uniform_buffers(self, chunk, B):
        """Two pinned host buffers + two device buffers + two events for the pipelined uniform stream of `run`."""
        if getattr(self, '_ub_key', None) != (chunk, B):
            lib = ctx().lib
            host, dev, evs, raw = [], [], [], []
            for _ in range(2):
                p = C.c_void_p()
                lib.vms_malloc_host(C.byref(p), chunk * B * 8)
                raw.append(p.value)
                buf = (C.c_byte * (chunk * B * 8)).from_address(p.value)
                host.append(np.frombuffer(buf, dtype=np.float64).reshape(chunk, B))
                dev.append(Tensor((chunk, B), np.float64))
                ev = C.c_void_p()
                lib.vms_event_create(C.byref(ev))
                evs.append(ev.value)
            self._ub_key, self._ub = (chunk, B), (host, dev, evs, raw)
        return self._ub

    def sync_counters(self):
        """Fold the device acceptance counter into `_num_acc` after device-resident `run_fused` calls."""
        if self._fused_plan() is not None:
            self._fold()
        nb = self._nb or None
        if nb:
            total = int(nb['n_acc'].numpy()[0])
            self._num_acc += float(total - nb['folded'])
            nb['folded'] = total

    @staticmethod
    def _energy_tensor(e):
        """Energies keep the callback's floating type: mcmc.py:116 is evaluated by NumPy in float64 for a float64 callback
        (tests/test_mcmc.py:28-32) and in float32 for a float32 one (a tfp log_prob, MC notebook cell 38)."""
        if isinstance(e, Tensor):
            return e if e.dtype in (np.float32, np.float64) else as_tensor(e.numpy(), dtype=np.float64)
        e = np.asarray(e)
        return Tensor.from_numpy(e, dtype=np.float32 if e.dtype == np.float32 else np.float64)

    def _energies(self, configs_host, configs_dev):
        """energy_func is the reference's host callable on NumPy arrays; a callable flagged `on_device` instead maps a
        device Tensor [B, D] float32 to a device Tensor [B] (float32 or float64) with no PCIe round trip."""
        if getattr(self.energy_func, 'on_device', False):
            return self._energy_tensor(self.energy_func(configs_dev))
        if configs_host is None:
            configs_host = configs_dev.numpy()
        return self._energy_tensor(self.energy_func(configs_host))

    def _device_step(self, x1, e_old, n_acc=None):
        """One MC step with the chain state on the device: x1 [B, D] float32, e_old [B] float64 -> (x2, e_out, acc, n_acc).
        The six distribution evaluations of mcmc.py:100-108 are op-by-op kernels, the accept rule is `vms_mc_accept`; the
        uniforms are this step's draw of the host PCG64 stream (mcmc.py:119)."""
        c = ctx()
        B = x1.shape[0]
        # forward proposal: encode, move in latent space, decode (mcmc.py:100-103)
        z1, log_z1_given_x1 = self.vae.encoder(x1).experimental_sample_and_log_prob()
        z2, log_z2 = self.vae.prior(z1).experimental_sample_and_log_prob()
        x2, log_x2_given_z2 = self.vae.decoder(z2).experimental_sample_and_log_prob()
        forward_log_p = log_z1_given_x1 + log_z2 + log_x2_given_z2

        # reverse proposal probability (mcmc.py:106-109)
        log_z2_given_x2 = self.vae.encoder(x2).log_prob(z2)
        log_z1 = self.vae.prior(z2).log_prob(z1)
        log_x1_given_z1 = self.vae.decoder(z1).log_prob(x1)
        reverse_log_p = log_z2_given_x2 + log_z1 + log_x1_given_z1

        x2 = x2.contig()
        e_new = self._energies(None, x2)

        # accept / reject on the device with the host's uniform stream (mcmc.py:116-128)
        log_rand = Tensor.from_numpy(np.log(self._rng.random(size=B)))
        if e_new.dtype != e_old.dtype:  # NumPy would promote the mixed pair
            e_new, e_old = (e if e.dtype == np.float64 else Tensor.from_numpy(e.numpy(), dtype=np.float64) for e in (e_new, e_old))
        e_out = Tensor((B, ), e_new.dtype)
        acc = Tensor((B, ), np.uint8)
        if n_acc is None:
            n_acc = Tensor.zeros((1, ), np.uint64)
        accept = c.lib.vms_mc_accept if e_new.dtype == np.float64 else c.lib.vms_mc_accept_f32
        accept(e_new.ptr, e_old.ptr, forward_log_p.ptr, reverse_log_p.ptr, log_rand.ptr, B, x2.shape[1],
                            x1.ptr, x2.ptr, e_out.ptr, acc.ptr, n_acc.ptr, c.stream)
        return x2, e_out, acc, n_acc

    def single_step(self, configs, energies=None):
        configs = np.array(configs.numpy() if isinstance(configs, Tensor) else configs)
        x1 = Tensor.from_numpy(configs, dtype=np.float32)
        B = x1.shape[0]
        x1 = x1.reshape(B, -1)
        e_old = self._energies(configs, x1) if energies is None else self._energy_tensor(energies)
        x2, e_out, acc, n_acc = self._device_step(x1, e_old)
        self._num_trials += B
        self._num_acc += float(n_acc.numpy()[0])
        self._last_acc = acc
        new_configs = x2.numpy().reshape(configs.shape).astype(configs.dtype, copy=False)
        rej = acc.numpy() == 0
        if configs.dtype != np.float32 and rej.any():
            new_configs = np.array(new_configs)
            new_configs[rej, ...] = configs[rej, ...]  # mcmc.py:126: rejected rows are the caller's own rows, unrounded
        return new_configs, e_out.numpy()

    def run_device(self, configs, energies=None, n_steps=1, configs_dev=None, energies_dev=None):
        """`run` for models outside the fused kernel's family with the chain state kept on the device for the whole loop:
        per step the only host traffic is the upload of that step's sampling noise and B uniforms (and the energy
        callback's round trip when `energy_func` is a host function); the accept counters are read once at the end.  Same
        decisions as n_steps calls of `single_step` under the same seeds.  With `configs_dev` ([B, D] float32 device
        tensor) the state is taken from and returned on the device."""
        on_dev = configs_dev is not None
        if on_dev:
            x = configs_dev
            shape, B = x.shape, x.shape[0]
        else:
            configs = np.array(configs.numpy() if isinstance(configs, Tensor) else configs)
            shape, B = configs.shape, configs.shape[0]
            x = Tensor.from_numpy(configs, dtype=np.float32).reshape(B, -1)
        if energies_dev is not None:
            e = energies_dev
        elif energies is not None:
            e = self._energy_tensor(energies)
        else:
            e = self._energies(None if on_dev else configs, x)
        total = Tensor.zeros((1, ), np.uint64)
        acc = None
        for _ in range(n_steps):
            x, e, acc, _n = self._device_step(x, e, n_acc=total)
        self._num_trials += B * n_steps
        self._num_acc += float(total.numpy()[0])
        self._last_acc = acc
        if on_dev:
            return x, e
        return x.numpy().reshape(shape), e.numpy()

    # ------------------------------------------------------------------------------ fused path, MC-notebook family (C4b)
    fuse_notebook = True  # set False to force the op-by-op path for this family (cross-checks)

    def _nb_plan(self):
        """State of the fused kernel for the model family of examples/MC_Moves_with_VAEs.ipynb (`vms_mc_nb_run`), or None
        when (vae, energy) is anything else: FCDeepNN(2 -> H -> 2) + IndependentNormal(1) encoder, RQSSplineMAF prior over a
        one-dimensional N(0, 1), FCDeepNN(1 -> H -> (2, 2)) + conditional AutoregressiveBlockwise(2 Normal) decoder with a
        three-hidden-layer MADE, GaussianMixtureEnergy.  The structure is cached here; weight pointers and the prior's knot
        tables are re-derived whenever the package's parameter epoch moved (`_abi.param_epoch()`: bumped by set_weights /
        assign / training steps), so training between calls is seen.  Code that overwrites weights through raw device
        pointers must call `_abi.bump_param_epoch()` itself."""
        if self._nb is not None:
            return self._nb or None
        self._nb = False
        if not self.fuse_notebook:
            self._nb = None
            return None
        from . import dists, flows, mappings
        from . import _protocols as PR
        E, vae = self.energy_func, self.vae
        if not isinstance(E, GaussianMixtureEnergy) or E.locs.shape[1] != 2 or not 1 <= len(E.probs) <= 16:
            return None
        enc, dec, prior = (getattr(vae, k, None) for k in ('encoder', 'decoder', 'prior'))

        def two_dense(mtd, n_in, n_out):
            m = getattr(mtd, 'mapping', None)
            if not isinstance(m, mappings.FCDeepNN) or not getattr(m, 'layer_list', None):
                return None
            if m.batch_norm or m.any_periodic:
                return None
            d = [l for l in m.layer_list if isinstance(l, PR.Dense)]
            if len(d) != 2 or d[0].act != PR._act_code('relu') or d[1].act != 0 or not (d[0].use_bias and d[1].use_bias):
                return None
            if d[0].kernel.shape[0] != n_in or d[1].units != n_out:
                return None
            return d

        ed = getattr(enc, 'distribution', None)
        if not isinstance(ed, PR.IndependentNormal) or ed.event_size != 1:
            return None
        dd = getattr(dec, 'distribution', None)
        if not isinstance(dd, dists.AutoregressiveBlockwise) or dd.num_dofs != 2 or not dd.conditional:
            return None
        if any(d is not dists.Normal for d in dd.dist_classes) or list(dd.param_nums) != [2, 2]:
            return None
        if any(getattr(t, 'dist_kind', None) != PR.DIST_NORMAL for t in dd.param_transforms):
            return None
        net = getattr(dd, 'auto_net', None)
        if net is None or len(net.layers) != 4 or net.cond_size != 1 or net.params != 2:
            return None
        e_l, d_l = two_dense(enc, 2, 2), two_dense(dec, 1, 4)
        if e_l is None or d_l is None:
            return None
        flow = getattr(prior, 'flow', None)
        if not isinstance(prior, dists.FlowedDistribution) or not isinstance(flow, flows.RQSSplineMAF):
            return None
        if flow.conditional or flow.batch_norm or flow.before_flow_transform is not None or \
                flow.after_flow_transform is not None or getattr(flow, 'data_dim', None) != 1:
            return None
        try:
            base = prior.latent_dist(Tensor.zeros((2, 1)))
        except Exception:
            return None
        if not isinstance(base, PR.StandardNormal) or base.event_size != 1:
            return None
        blocks = [b.bijector_fn for b in flow.chain.bijectors[::-1]]  # sampling order: block 0 first
        msb = blocks[0]
        if any((b.num_bins, b.bin_min, b.bin_max) != (msb.num_bins, msb.bin_min, msb.bin_max) for b in blocks):
            return None
        m = _abi.McNbModel()
        m.dx, m.dz, m.enc_hidden, m.dec_hidden = 2, 1, e_l[0].units, d_l[0].units
        for k in range(3):
            m.made_hidden[k] = net.layers[k].units
        m.made_act = net.act
        # which dof the masks put first: the other dof's input row and the first dof's output columns are masked out
        # (set_weights and the gradient masks keep them exactly zero)
        mk = net.masks
        m.made_first_dof = -1
        for f in (0, 1):
            if not mk[0][1 - f, :].any() and not mk[-1][:, 2 * f:2 * f + 2].any():
                m.made_first_dof = f
        if os.environ.get('VMS_NB_GENERIC') == '1':  # cross-checks: tfp's D + 1 sampling passes + log_prob pass, no mask knowledge
            m.made_first_dof = -1
        m.n_blocks, m.n_bins, m.range_min, m.range_max = len(blocks), msb.num_bins, msb.bin_min, msb.bin_max
        m.n_comp = len(E.probs)
        if not _abi.load().vms_mc_nb_supported(C.byref(m)):
            return None
        self._nb = dict(model=m, enc=e_l, dec=d_l, net=net, blocks=blocks,
                        gmm=(Tensor.from_numpy(np.log(E.probs)), Tensor.from_numpy(E.locs), Tensor.from_numpy(E.scales)),
                        n_acc=Tensor.zeros((1, ), np.uint64), n_unc=Tensor.zeros((1, ), np.uint64), folded=0)
        return self._nb

    def _nb_model(self, nb):
        """Fills the kernel's model record from the live weights and rebuilds the prior's knot tables (one row through each
        block's three conditioner networks + vms_rqs_knot_table)."""
        c = ctx()
        m = nb['model']
        cached = nb.get('cache')
        if cached is not None and cached[0] == _abi.param_epoch():
            return m, cached[1]  # no parameter changed since the tables were built (same pointers, same values)
        keep = []

        def ptr(t):
            t = t if t.contiguous else t.contig()
            keep.append(t)
            return t.ptr

        (m.enc_W0, m.enc_b0), (m.enc_W1, m.enc_b1) = ((ptr(l.kernel), ptr(l.bias)) for l in nb['enc'])
        (m.dec_W0, m.dec_b0), (m.dec_W1, m.dec_b1) = ((ptr(l.kernel), ptr(l.bias)) for l in nb['dec'])
        net = nb['net']
        for k, lay in enumerate(net.layers):
            m.made_W[k], m.made_b[k], m.made_Wc[k] = ptr(lay.kernel), ptr(lay.bias), ptr(net.cond_kernels[k])
        K, nbk = m.n_bins, m.n_blocks
        ts = int(c.lib.vms_rqs_knot_table_doubles(K))
        tables = Tensor((nbk, ts), np.float64)
        zero = Tensor.zeros((1, 1))
        for b, msb in enumerate(nb['blocks']):
            rw, rh, rs_ = (n(zero).contig() for n in (msb.bin_widths, msb.bin_heights, msb.knot_slopes))
            keep += [rw, rh, rs_]
            c.lib.vms_rqs_knot_table(rw.ptr, rh.ptr, rs_.ptr, 1, K, m.range_min, m.range_max, tables.ptr + b * ts * 8, c.stream)
        m.tables = tables.ptr
        m.gmm_log_w, m.gmm_loc, m.gmm_scale = (t.ptr for t in nb['gmm'])
        keep.append(tables)
        nb['cache'] = (_abi.param_epoch(), keep)
        return m, keep

    def run_nb(self, configs, energies=None, n_steps=1, noise=None, trace=False, configs_dev=None, energies_dev=None):
        """n_steps MC steps of the notebook family in one `vms_mc_nb_run` launch.  `noise` [n_steps, B, 4] injects the
        sampling noise (eps(z1) | eps(z2) | eps(x2), the order mcmc.py:100-102 draws it; parity tests); otherwise the device
        Philox stream keyed by (seed, global chain, step).  Accept uniforms: this object's PCG64 stream, regenerated on the
        device (`device_rng`) with the host-stream re-run for uncertain decisions, or drawn on the host and uploaded."""
        nb = self._nb_plan()
        if nb is None:
            raise NotImplementedError('run_nb: needs the MC notebook model family and a GaussianMixtureEnergy')
        c = ctx()
        if configs_dev is None:
            configs = np.array(configs.numpy() if isinstance(configs, Tensor) else configs)
            x = Tensor.from_numpy(configs.reshape(configs.shape[0], -1), dtype=np.float32)
        else:
            x = configs_dev
        B = x.shape[0]
        if energies_dev is not None:
            e, valid = energies_dev, 1
        elif energies is None:
            e, valid = Tensor((B, ), np.float32), 0
        else:
            e, valid = Tensor.from_numpy(np.asarray(energies, np.float32)), 1
        if e.dtype != np.float32:
            raise TypeError('run_nb: energies are float32 (the dtype of the mixture log_prob)')
        m, keep = self._nb_model(nb)
        nz = None if noise is None else Tensor.from_numpy(np.ascontiguousarray(noise, np.float32))
        P_ = lambda t: None if t is None else t.ptr
        chain0, n_global = self.stream_layout if self.stream_layout is not None else (0, B)

        def traces(log_u=None):
            if not trace:
                return {}
            return dict(acc=Tensor((n_steps, B), np.uint8), fwd=Tensor((n_steps, B)), rev=Tensor((n_steps, B)),
                        e_new=Tensor((n_steps, B)), log_u=log_u if log_u is not None else Tensor((n_steps, B), np.float64))

        def launch(log_u, stream, tr):
            c.lib.vms_mc_nb_run(C.byref(m), x.ptr, e.ptr, valid, P_(nz), self._noise_seed, self._step0, P_(log_u),
                                None if stream is None else C.byref(stream), int(chain0), B, n_steps, nb['n_acc'].ptr,
                                nb['n_unc'].ptr, P_(tr.get('acc')), P_(tr.get('fwd')), P_(tr.get('rev')), P_(tr.get('e_new')),
                                P_(tr.get('log_u')), c.stream)

        stream = self._pcg_stream(chain0, n_global) if self.device_rng else None
        done, tr = False, {}
        if stream is not None:
            x_save, e_save = Tensor(x.shape), Tensor((B, ))
            c.lib.vms_memcpy_d2d(x_save.ptr, x.ptr, x.nbytes, c.stream)
            if valid:
                c.lib.vms_memcpy_d2d(e_save.ptr, e.ptr, e.nbytes, c.stream)
            acc_before = Tensor((1, ), np.uint64)
            c.lib.vms_memcpy_d2d(acc_before.ptr, nb['n_acc'].ptr, 8, c.stream)
            if not (configs_dev is not None and not self.check_uncertain):
                nb['n_unc'].fill_zero()
            tr = traces()
            launch(None, stream, tr)
            if configs_dev is not None and not self.check_uncertain:
                done = True  # device-resident loops read `uncertain()` themselves, once per loop
            elif int(nb['n_unc'].numpy()[0]) == 0:
                done = True
            else:  # repeat the call on the NumPy stream from the saved state
                self.host_stream_reruns += 1
                c.lib.vms_memcpy_d2d(x.ptr, x_save.ptr, x.nbytes, c.stream)
                if valid:
                    c.lib.vms_memcpy_d2d(e.ptr, e_save.ptr, e.nbytes, c.stream)
                c.lib.vms_memcpy_d2d(nb['n_acc'].ptr, acc_before.ptr, 8, c.stream)
            if done:
                self._rng.bit_generator.advance(n_steps * n_global)
        if not done:
            h = np.empty((n_steps, B))
            self._host_uniform_logs(h, chain0, n_global)  # mcmc.py:119, step by step
            log_u = Tensor.from_numpy(h)
            tr = traces(log_u)
            launch(log_u, None, tr)
        self._step0 += n_steps
        self._num_trials += B * n_steps
        if configs_dev is not None:
            self._nb_keep = keep  # the launch may still be running: its inputs stay alive until the next call
            return x, e
        total = int(nb['n_acc'].numpy()[0])
        self._num_acc += float(total - nb['folded'])
        nb['folded'] = total
        self._last_trace = {k: t.numpy() for k, t in tr.items()}
        return x.numpy().reshape(configs.shape), e.numpy()

    def run(self, configs, energies=None, n_steps=1):
        if self._fused_plan() is not None:
            return self.run_fused(configs, energies=energies, n_steps=n_steps)
        if self._nb_plan() is not None:
            return self.run_nb(configs, energies=energies, n_steps=n_steps)
        return self.run_device(configs, energies=energies, n_steps=n_steps)
