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

    def allreduce_sum_(self, tensor, ext_stream=None):
        if self.world == 1:
            return
        if ext_stream is not None:
            with self.torch.cuda.stream(ext_stream):
                self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
        else:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)

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
