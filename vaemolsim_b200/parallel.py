"""Multi-GPU plumbing: one process per GPU; rendezvous / barrier / scalar reductions over plain TCP sockets, the gradient
exchange as a CUDA-IPC peer-memory kernel (`PeerExchange`), NCCL through ctypes as the fallback.  No torch on the product
path (`Group(backend='gloo')` keeps a torch.distributed implementation for the CPU tests).

The hot path shards naturally: configuration batches, MC chains and reference sites are independent, so
ranks own contiguous row blocks and exchange nothing on the data path.  Data-parallel TRAINING has exactly one exchange
step per iteration -- the sum of the flat gradient (44,396 floats for C2) -- done in place on the plan's gradient buffer
and followed by the Adam kernel with grad_scale = 1 / world_size (loss is a batch MEAN, losses.py:253).

torch is imported only when a torch backend is requested explicitly (tests).
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
    """Process group of one node: rendezvous, barrier, max / sum of a scalar, all-gather of small byte strings (the CUDA
    IPC handles of `PeerExchange`), and the FALLBACK device all-reduce (NCCL through ctypes, no torch).

    Default backend 'socket': rank 0 listens on MASTER_ADDR : (VMS_RDZV_PORT or MASTER_PORT + 1000) -- torchrun's own
    store owns MASTER_PORT -- and relays; everything is plain TCP + pickle, so the product path imports neither torch nor
    torch.distributed (north_star: "no PyTorch").  backend='gloo' (or 'torch-nccl') keeps the torch.distributed
    implementation: it is what the CPU tests of the sharding logic run on (tests/test_parallel.py)."""

    def __init__(self, backend=None):
        self.rank, self.local_rank, self.world = env_world()
        self.torch = None
        self.backend = backend or os.environ.get('VMS_GROUP_BACKEND', 'socket')
        self._conns, self._sock, self._nccl = [], None, None
        if self.world == 1:
            return
        if self.backend == 'socket':
            self._init_socket()
            return
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        tb = 'nccl' if self.backend == 'torch-nccl' else 'gloo'
        self._torch_backend = tb
        if tb == 'nccl':
            torch.cuda.set_device(self.local_rank)
        if not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29512')
            dist.init_process_group(backend=tb, rank=self.rank, world_size=self.world)

    # ------------------------------------------------------------------ socket backend
    def _init_socket(self):
        import socket
        import time
        addr = os.environ.get('MASTER_ADDR', '127.0.0.1')
        port = int(os.environ.get('VMS_RDZV_PORT', 0)) or (int(os.environ.get('MASTER_PORT', '29512')) + 1000 - 1024) % 64000 + 1024
        deadline = time.time() + float(os.environ.get('VMS_RDZV_TIMEOUT', '120'))
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            while True:
                try:
                    srv.bind((addr, port))
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.2)
            srv.listen(self.world)
            srv.settimeout(max(1.0, deadline - time.time()))
            conns = {}
            while len(conns) < self.world - 1:
                cn, _ = srv.accept()
                cn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                cn.settimeout(None)
                r = int.from_bytes(self._recv_exact(cn, 4), 'little')
                conns[r] = cn
            self._sock = srv
            self._conns = [conns[r] for r in range(1, self.world)]
        else:
            while True:
                try:
                    cn = socket.create_connection((addr, port), timeout=5.0)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.1)
            cn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            cn.settimeout(None)
            cn.sendall(int(self.rank).to_bytes(4, 'little'))
            self._conns = [cn]

    @staticmethod
    def _recv_exact(cn, n):
        buf = bytearray()
        while len(buf) < n:
            chunk = cn.recv(n - len(buf))
            if not chunk:
                raise RuntimeError('Group: a peer closed its connection')
            buf += chunk
        return bytes(buf)

    def _send_msg(self, cn, payload):
        cn.sendall(len(payload).to_bytes(8, 'little') + payload)

    def _recv_msg(self, cn):
        return self._recv_exact(cn, int.from_bytes(self._recv_exact(cn, 8), 'little'))

    def all_gather_bytes(self, payload):
        """List of every rank's byte string, in rank order, on every rank."""
        payload = bytes(payload)
        if self.world == 1:
            return [payload]
        if self.backend != 'socket':
            out = [None] * self.world
            self.dist.all_gather_object(out, payload)
            return out
        import pickle
        if self.rank == 0:
            parts = [payload] + [self._recv_msg(cn) for cn in self._conns]
            blob = pickle.dumps(parts)
            for cn in self._conns:
                self._send_msg(cn, blob)
            return parts
        self._send_msg(self._conns[0], payload)
        return pickle.loads(self._recv_msg(self._conns[0]))

    def _gather_floats(self, value):
        import struct
        return [struct.unpack('<d', b)[0] for b in self.all_gather_bytes(struct.pack('<d', float(value)))]

    def barrier(self):
        if self.world == 1:
            return
        if self.backend == 'socket':
            self.all_gather_bytes(b'')
        else:
            self.dist.barrier()

    def max(self, value):
        if self.world == 1:
            return float(value)
        if self.backend == 'socket':
            return max(self._gather_floats(value))
        dev = 'cuda' if self._torch_backend == 'nccl' else 'cpu'
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value):
        if self.world == 1:
            return float(value)
        if self.backend == 'socket':
            return float(sum(self._gather_floats(value)))  # rank order: the same value on every rank
        dev = 'cuda' if self._torch_backend == 'nccl' else 'cpu'
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    # -- the one data-path collective: flat gradient sum (FALLBACK; the product path is PeerExchange)
    def nccl(self):
        """NCCL communicator over the node's ranks, created on first use (ctypes on libnccl; the unique id travels over
        this group)."""
        if self._nccl is None:
            self._nccl = NcclComm(self)
        return self._nccl

    def allreduce_sum_device_(self, ptr, n, stream, host_sync=None):
        """In-place NCCL sum-allreduce of `n` float32 at device pointer `ptr`, enqueued on the library's stream (ordered
        after the kernels that produced the gradient, no host synchronisation needed)."""
        if self.world == 1:
            return
        self.nccl().allreduce_sum_f32(ptr, n, stream)

    def allreduce_sum_numpy_(self, array):
        """Host variant used by the CPU tests of the sharding logic: sum in rank order (deterministic)."""
        if self.world == 1:
            return array
        if self.backend == 'socket':
            parts = self.all_gather_bytes(np.ascontiguousarray(array).tobytes())
            acc = np.zeros_like(array)
            for b in parts:
                acc += np.frombuffer(b, dtype=array.dtype).reshape(array.shape)
            array[...] = acc
            return array
        t = self.torch.from_numpy(array)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return array

    def close(self):
        if self._nccl is not None:
            self._nccl.close()
            self._nccl = None
        if self.world > 1 and self.backend == 'socket':
            try:
                self.barrier()
            except Exception:
                pass
            for cn in self._conns:
                cn.close()
            if self._sock is not None:
                self._sock.close()
            self._conns, self._sock = [], None
        elif self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


class NcclComm(object):
    """ncclAllReduce through ctypes (libnccl.so.2 of the `nvidia-nccl` wheel, or VMS_NCCL_LIB): the fallback exchange when
    CUDA IPC peer mapping is unavailable.  No torch involved: rank 0 creates the unique id, `Group.all_gather_bytes`
    distributes it."""

    def __init__(self, group):
        import ctypes as C
        from . import _abi
        self.C = C
        path = os.environ.get('VMS_NCCL_LIB')
        if not path:
            try:
                import importlib.util
                spec = importlib.util.find_spec('nvidia.nccl')
                base = list(spec.submodule_search_locations)[0]
                path = os.path.join(base, 'lib', 'libnccl.so.2')
            except Exception:
                path = 'libnccl.so.2'
        self.lib = C.CDLL(path)

        class UniqueId(C.Structure):
            _fields_ = [('internal', C.c_char * 128)]

        self.lib.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
        self.lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.ncclCommDestroy.argtypes = [C.c_void_p]
        self.lib.ncclGetErrorString.restype = C.c_char_p
        _abi.ctx()  # the device of this rank is current
        uid = UniqueId()
        if group.rank == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        # (string_at, not the c_char array's value: that one stops at the first NUL byte of the 128-byte id)
        raw = group.all_gather_bytes(C.string_at(C.byref(uid), 128) if group.rank == 0 else b'')[0]
        uid = UniqueId()
        C.memmove(C.byref(uid), raw.ljust(128, b'\0'), 128)
        self.comm = C.c_void_p()
        self._check(self.lib.ncclCommInitRank(C.byref(self.comm), group.world, uid, group.rank))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError('NCCL: %s' % self.lib.ncclGetErrorString(rc).decode())

    def allreduce_sum_f32(self, ptr, n, stream):
        self._check(self.lib.ncclAllReduce(ptr, ptr, int(n), 7, 0, self.comm, stream))  # ncclFloat32 = 7, ncclSum = 0

    def close(self):
        if getattr(self, 'comm', None):
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None


class PeerExchange(object):
    """The fused gradient-allreduce + Adam step over NVLink peer memory (`csrc/peer.cu`, `vms_peer_allreduce_adam`).

    Every rank allocates one buffer (two gradient slots + flags), the CUDA IPC handles travel through
    `Group.all_gather_bytes` (plumbing), and from then on a training step's exchange is a single kernel per
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
        handles = group.all_gather_bytes(bytes(handle))
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
        from . import _abi
        _abi.bump_param_epoch()
        if self.check_every and self.step % self.check_every == 0:
            self.check()
        self.ctx.lib.vms_peer_allreduce_adam(self.world, self.rank, self.bases, self.n, self.step, 1.0 / self.world,
                                             fused.theta.ptr, fused.m.ptr, fused.v.ptr, fused.t, opt.learning_rate,
                                             opt.beta_1, opt.beta_2, opt.epsilon, None if grad_out is None else grad_out.ptr,
                                             self.ctx.stream)

    def train_step(self, fused, x, eps, n, opt, scalars_ptr=None):
        """One data-parallel training step in one C call (`vms_elbo_train_step_peer`): forward + backward on this rank's
        batch (device tensors / pointers x, eps with n rows), the NVLink gradient exchange and Adam -- two launches when the
        whole-step tensor-core kernel serves the batch (its finish kernel does the exchange), otherwise the plan's own
        launches + `vms_peer_allreduce_adam`.  All ranks must call it once per step."""
        self.step += 1
        fused.t += 1
        from . import _abi
        _abi.bump_param_epoch()
        if self.check_every and self.step % self.check_every == 0:
            self.check()
        xp = x.ptr if hasattr(x, 'ptr') else x
        ep = eps.ptr if hasattr(eps, 'ptr') else eps
        self.ctx.lib.vms_elbo_train_step_peer(fused.handle, fused.theta.ptr, xp, ep, int(n),
                                              fused.scalars.ptr if scalars_ptr is None else scalars_ptr, fused.m.ptr,
                                              fused.v.ptr, fused.t, opt.learning_rate, opt.beta_1, opt.beta_2, opt.epsilon,
                                              self.world, self.rank, self.bases, self.step, self.ctx.stream)

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
