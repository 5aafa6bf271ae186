"""Multi-GPU plumbing (one process per GPU, `torch.distributed` over NCCL / NVLink; gloo on CPU for the tests).

The hot path shards naturally: configuration batches, MC chains and reference sites are independent, so
ranks own contiguous row blocks and exchange nothing on the data path.  Data-parallel TRAINING has exactly one exchange
step per iteration -- the sum of the flat gradient (44,396 floats for C2) -- done in place on the plan's gradient buffer
and followed by the Adam kernel with grad_scale = 1 / world_size (loss is a batch MEAN, losses.py:253).

torch is imported only here and only when WORLD_SIZE > 1 (or explicitly): the single-GPU product path has no torch.
"""
import os

import numpy as np


def env_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def shard_rows(n_global, rank, world):
    """Contiguous block of rows owned by `rank`: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(int(n_global), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_row_seed(seed, row0):
    """Per-shard NumPy generator keyed by the GLOBAL index of the shard's first row, so that synthetic inputs (and
    therefore results) do not depend on the number of ranks."""
    return np.random.default_rng([int(seed), int(row0)])


class Group(object):
    """Thin wrapper over torch.distributed for the three things the path needs: barrier, max-reduce of a timing,
    in-place sum-allreduce of a device (or host) float32 buffer."""

    def __init__(self, backend=None):
        self.rank, self.local_rank, self.world = env_world()
        self.torch = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            if backend is None:
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            self.backend = backend
            if backend == 'nccl':
                torch.cuda.set_device(self.local_rank)
            if not dist.is_initialized():
                os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
                os.environ.setdefault('MASTER_PORT', '29512')
                dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max(self, value):
        if self.world == 1:
            return float(value)
        dev = 'cuda' if self.backend == 'nccl' else 'cpu'
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value):
        if self.world == 1:
            return float(value)
        dev = 'cuda' if self.backend == 'nccl' else 'cpu'
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    # -- the one data-path collective: flat gradient sum
    def wrap_device_buffer(self, ptr, n, stream):
        """torch view of `n` float32 at device pointer `ptr` (no copy) + torch handle of the library's stream, so the
        NCCL call is ordered after the kernels that produced the gradient without a host sync."""
        torch = self.torch

        class _CAI(object):
            __cuda_array_interface__ = {'shape': (int(n), ), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}

        t = torch.as_tensor(_CAI(), device='cuda:%d' % self.local_rank)
        ext = torch.cuda.ExternalStream(int(stream), device='cuda:%d' % self.local_rank)
        return t, ext

    def allreduce_sum_(self, tensor, ext_stream=None, host_sync=None):
        """NCCL sum-allreduce of a wrapped device buffer.  This is the FALLBACK exchange (the product path is
        `PeerExchange`); it brackets the collective with host synchronisation of both streams: enqueueing NCCL on the
        library's own non-blocking stream worked at 2 ranks but never returned at 4 on the test box, and a
        fallback must above all terminate."""
        if self.world == 1:
            return
        if host_sync is not None:
            host_sync()                                   # the library stream has produced the gradient
        self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
        if self.backend == 'nccl':
            self.torch.cuda.current_stream().synchronize()    # ... and NCCL has reduced it before the library reads it

    def allreduce_sum_numpy_(self, array):
        """Host (gloo) variant used by the CPU tests of the sharding logic."""
        if self.world == 1:
            return array
        t = self.torch.from_numpy(array)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return array

    def close(self):
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


class PeerExchange(object):
    """The fused gradient-allreduce + Adam step over NVLink peer memory (`csrc/peer.cu`, `vms_peer_allreduce_adam`).

    Every rank allocates one buffer (two gradient slots + flags), the CUDA IPC handles travel through
    `torch.distributed.all_gather_object` (plumbing), and from then on a training step's exchange is a single kernel per
    rank with no host involvement.  Construction raises if peer mapping is unavailable; callers fall back to
    `Group.allreduce_sum_` (NCCL) + `vms_adam_step`."""

    def __init__(self, group, n_params):
        import ctypes as C
        from . import _abi
        self.C, self.group, self.n = C, group, int(n_params)
        self.ctx = _abi.ctx()
        lib = self.ctx.lib
        self.world, self.rank = group.world, group.rank
        if self.world > 8:
            raise NotImplementedError('PeerExchange: at most 8 ranks (one NVSwitch node)')
        nbytes = int(lib.vms_peer_buffer_bytes(self.n))
        self.buf = _abi.Tensor((nbytes // 4, ))
        lib.vms_memset(self.buf.ptr, 0, nbytes, self.ctx.stream)
        self.ctx.synchronize()
        handle = (C.c_ubyte * 64)()
        lib.vms_ipc_get_handle(self.buf.ptr, handle)
        handles = [None] * self.world
        if self.world > 1:
            group.dist.all_gather_object(handles, bytes(handle))
        self.opened = []
        bases = (C.c_void_p * self.world)()
        error = None
        try:
            for r in range(self.world):
                if r == self.rank:
                    bases[r] = self.buf.ptr
                else:
                    p = C.c_void_p()
                    lib.vms_ipc_open_handle((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p))
                    self.opened.append(p.value)
                    bases[r] = p.value
        except Exception as e:  # no peer mapping from this rank: every rank must learn it (no one-sided exit)
            error = e
        self.bases = bases
        self.step = 0
        # collective agreement doubles as the barrier: every buffer is zeroed and mapped before anyone signals
        n_ok = group.sum(0.0 if error is not None else 1.0)
        if n_ok != self.world:
            for p in self.opened:
                lib.vms_ipc_close_handle(p)
            self.opened = []
            raise RuntimeError('PeerExchange: peer mapping failed on %d of %d ranks%s' %
                               (self.world - int(n_ok), self.world, '' if error is None else ' (this rank: %s)' % error))

    def next_slot(self):
        """Device pointer the NEXT step's local gradient must be written to."""
        return self.buf.ptr + 4 * self.n * ((self.step + 1) & 1)

    def allreduce_adam(self, fused, opt, grad_out=None):
        """Sum the ranks' gradients of this step (slot `next_slot()` of every buffer), scale by 1 / world, Adam-update
        `fused.theta / m / v` in place.  All ranks must call it once per step."""
        self.step += 1
        fused.t += 1
        if self.check_every and self.step % self.check_every == 0:
            self.check()
        self.ctx.lib.vms_peer_allreduce_adam(self.world, self.rank, self.bases, self.n, self.step, 1.0 / self.world,
                                             fused.theta.ptr, fused.m.ptr, fused.v.ptr, fused.t, opt.learning_rate,
                                             opt.beta_1, opt.beta_2, opt.epsilon, None if grad_out is None else grad_out.ptr,
                                             self.ctx.stream)

    check_every = 256  # steps between host-side failure checks inside allreduce_adam (0: never)

    def timed_out(self):
        """True if an exchange of the job gave up waiting for a peer (the kernel's bound, VMS_PEER_TIMEOUT_MS): this rank
        timed out itself (flag 32 + rank) or a peer poisoned the job (flag 48).  No parameter update was applied from that
        step on."""
        import numpy as np
        self.ctx.synchronize()
        flags = self.buf.numpy().view(np.uint64)[self.n:self.n + 64]   # the flags follow the 2 n gradient floats
        return bool(flags[32 + self.rank] != 0 or flags[48] != 0)

    def check(self):
        """Raises if the exchange failed (a peer stalled longer than the bound); called every `check_every` steps."""
        if self.timed_out():
            raise RuntimeError('PeerExchange: a rank stopped answering within VMS_PEER_TIMEOUT_MS; parameter updates were '
                               'suspended on every replica from that step on (restart from the last checkpoint or use the '
                               'NCCL exchange)')

    def close(self):
        self.ctx.synchronize()
        self.group.barrier()
        for p in self.opened:
            self.ctx.lib.vms_ipc_close_handle(p)
        self.opened = []
