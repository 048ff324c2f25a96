"""Parallelism building blocks on top of the collective API.

The reference is a communication library and implements no model-parallel
strategy itself; it supplies the primitives they are built from.
This module packages those primitives the way training / serving code consumes
them on a B200 node:

* `init_from_env`        one rank per process (torchrun): pick the backend, build the world
* `TensorGroup`          torch-tensor front end: all_reduce / all_gather / reduce_scatter /
                         broadcast / all_to_all / send / recv on CUDA tensors (zero-copy when the
                         tensor already lives in the symmetric heap, staged otherwise)
* `GradBucket`           DP / ZeRO gradient bucket living in the heap: fill, then one all-reduce
                         (or reduce_scatter + all_gather) per bucket, stream ordered
* `RowParallelLinear`    TP row-parallel linear whose GEMM epilogue IS the reduce-scatter
                         (tcgen05 plugin), sequence-parallel output
* `ring_exchange`        neighbour send/recv step for context-parallel / ring-attention schedules
* `process_group`        `torch.distributed` backend "accl" (import accl_b200.parallel.process_group): the standard
                         ProcessGroup interface incl. DistributedDataParallel on top of this library
* `strategies`           ZeRO-1 optimizer step, column-parallel linear, pipeline hand-off, MoE dispatch / combine,
                         Ulysses sequence <-> head re-sharding (accl_b200/parallel/strategies.py)
"""
import os

import torch

from .. import _C
from ..core import Accl, Buffer, SUM, cuda_rank, socket_rank
from ..utils.dtypes import to_accl


def init_from_env(backend=None, **kw):
    """RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT -> an initialised `Accl`.

    backend: "cuda" (default when a GPU and the CUDA backend are available) or "emulator".
    Extra keyword arguments go to `initialize` (eager / rendezvous geometry)."""
    if backend is None:
        backend = "cuda" if (torch.cuda.is_available() and _C.with_cuda and _C.cuda_driver_available()) else "emulator"
    init_keys = ("n_egr_rx_bufs", "egr_rx_buf_size", "max_egr_size", "max_rndzv_size")
    init_kw = {k: kw.pop(k) for k in list(kw) if k in init_keys}
    if backend == "cuda":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        if "engine" not in kw and os.environ.get("ACCL_PG_ENGINE") is not None:
            kw["engine"] = os.environ["ACCL_PG_ENGINE"] not in ("", "0")
        if "max_ctas" not in kw and os.environ.get("ACCL_MAX_CTAS"):
            kw["max_ctas"] = int(os.environ["ACCL_MAX_CTAS"])  # channel cap (default 128; the planner sizes every call)
        if "heap_mb" not in kw and os.environ.get("ACCL_HEAP_MB"):
            kw["heap_mb"] = int(os.environ["ACCL_HEAP_MB"])   # symmetric heap per rank (default 1024)
        acc = cuda_rank(**kw)
        init_kw.setdefault("n_egr_rx_bufs", 4)
        init_kw.setdefault("egr_rx_buf_size", 64 << 10)
        init_kw.setdefault("max_egr_size", 64 << 10)
        init_kw.setdefault("max_rndzv_size", 1 << 30)
    else:
        acc = socket_rank(**kw)
    acc.initialize(**init_kw)
    if backend != "cuda" and os.environ.get("ACCL_EMU_ONE_HOP", "0") not in ("", "0"):
        acc.set_one_hop_schedules(True)   # emulator: the B200 backend's schedules instead of the reference's rings / trees
    return acc


class TensorGroup:
    """torch.distributed-flavoured calls on top of one `Accl` (device-resident tensors)."""

    def __init__(self, accl: Accl, comm_id=0, scratch_bytes=64 << 20):
        self.accl = accl
        self.comm_id = comm_id
        self.world = len(accl.get_comm_group(comm_id))
        self.rank = accl.get_comm_rank(comm_id)
        self._scratch = {}
        self._wrapped = {}  # heap-resident torch tensors seen so far -> Buffer views
        self._scratch_bytes = scratch_bytes
        # CUDA: operands are device resident and calls are stream ordered.  Emulator: the tensor side of a
        # buffer is its host mirror, so calls are blocking and sync to / from the engine's memory themselves.
        self._res = bool(accl.is_cuda)
        self._async = bool(accl.is_cuda)
        # engine mode: asynchronous point-to-point calls do not hold the stream (they may stay parked until the peer posts
        # its side), so a staged receive has to be waited for on the host before its staging buffer is read
        self._engine = bool(accl.is_cuda) and "mode=engine" in accl.describe()
        self._pending = []  # asynchronous requests whose return code has not been looked at yet

    # -- request bookkeeping ---------------------------------------------------
    def _track(self, req):
        """Keep an asynchronous request until it has completed, then check its return code: a device-side timeout,
        tag mismatch or protocol error must not go unnoticed while training continues on garbage."""
        if req is None or not hasattr(req, "test"):
            return req
        self._pending.append(req)
        self.check(block=len(self._pending) > 64)
        return req

    def check(self, block=False):
        """Raise if any finished call reported an error; with block=True wait for everything outstanding."""
        keep = []
        for r in self._pending:
            if block:
                r.wait()
            if block or r.test():
                rc = r.retcode()
                if rc:
                    self._pending = [q for q in self._pending if q is not r]
                    raise RuntimeError(f"accl call failed on rank {self.rank}: {_C.error_to_string(rc)} (0x{rc:x})")
            else:
                keep.append(r)
        self._pending = keep

    def _t(self, buf):
        """The torch tensor aliasing a buffer on this backend."""
        return buf.dev if self._res else buf.host

    # -- heap-resident tensors -------------------------------------------
    def empty(self, *shape, dtype=torch.float32):
        """A tensor allocated in the symmetric heap: zero-copy operand of every collective."""
        n = 1
        for s in shape:
            n *= s
        buf = self.accl.create_buffer(n, dtype)
        t = self._t(buf).view(*shape)
        t._accl_buffer = buf
        return t

    def _buffer_of(self, t: torch.Tensor, role):
        buf = getattr(t, "_accl_buffer", None)
        if buf is not None and buf.length == t.numel():
            return buf, False
        if self._res and t.is_contiguous() and self.accl.heap_contains(t):
            # memory from the heap pool (Accl.heap_mem_pool): zero-copy, the view is cached by address
            key = (t.data_ptr(), t.numel(), t.dtype)
            buf = self._wrapped.get(key)
            if buf is None:
                if len(self._wrapped) > 256:
                    self._wrapped.clear()
                buf = self._wrapped[key] = self.accl.wrap_device(t)
            return buf, False
        # stage through ONE cached heap buffer per (role, dtype); it grows to the largest tensor seen (calls carry
        # their own element counts, so a longer buffer is fine).  Reuse is safe: staging copies and collectives are
        # ordered on the caller's stream.
        key = (role, t.dtype)
        cur = self._scratch.get(key)
        if cur is None or cur.length < t.numel():
            self._scratch.pop(key, None)
            cur = None  # the old allocation returns to the heap (same order on every rank) before the new one is made
            self._scratch[key] = self.accl.create_buffer(t.numel(), t.dtype)
        return self._scratch[key], True

    def _chunks(self, t: torch.Tensor):
        """Element-wise collectives on a tensor outside the heap that is larger than the staging budget: contiguous
        views of at most `scratch_bytes` each (None when the tensor can go in one piece)."""
        if getattr(t, "_accl_buffer", None) is not None or t.numel() * t.element_size() <= self._scratch_bytes:
            return None
        if self._res and self.accl.heap_contains(t):
            return None
        step = max(1, self._scratch_bytes // t.element_size())
        flat = t.view(-1) if t.is_contiguous() else None
        if flat is None:
            return None
        return [flat[o:o + step] for o in range(0, t.numel(), step)]

    def _run(self, fn, src: torch.Tensor, dst: torch.Tensor, src_elems, dst_elems):
        sb, s_staged = self._buffer_of(src, "s")
        db, d_staged = (sb, s_staged) if dst is src else self._buffer_of(dst, "d")
        if s_staged:
            self._t(sb)[:src.numel()].copy_(src.reshape(-1))
        req = self._track(fn(sb, db))
        if d_staged:
            dst.copy_(self._t(db)[:dst.numel()].view_as(dst))  # (works for non-contiguous destinations too)
        return req

    # -- collectives ---------------------------------------------------------
    def all_reduce(self, t: torch.Tensor, op=SUM):
        parts = self._chunks(t)
        if parts:
            for p in parts:
                req = self.all_reduce(p, op)
            return req
        n = t.numel()
        return self._run(lambda s, d: self.accl.allreduce(s, d, n, op, self.comm_id, self._res, self._res, run_async=self._async), t, t, n, n)

    def broadcast(self, t: torch.Tensor, root=0):
        parts = self._chunks(t)
        if parts:
            for p in parts:
                req = self.broadcast(p, root)
            return req
        n = t.numel()
        return self._run(lambda s, d: self.accl.bcast(s, n, root, self.comm_id, self._res, self._res, run_async=self._async), t, t, n, n)

    def all_gather_into_tensor(self, out: torch.Tensor, inp: torch.Tensor):
        n = inp.numel()
        return self._run(lambda s, d: self.accl.allgather(s, d, n, self.comm_id, self._res, self._res, run_async=self._async), inp, out, n, n * self.world)

    def reduce_scatter_tensor(self, out: torch.Tensor, inp: torch.Tensor, op=SUM):
        n = out.numel()
        return self._run(lambda s, d: self.accl.reduce_scatter(s, d, n, op, self.comm_id, self._res, self._res, run_async=self._async), inp, out, n * self.world, n)

    def all_to_all_single(self, out: torch.Tensor, inp: torch.Tensor):
        n = inp.numel() // self.world
        return self._run(lambda s, d: self.accl.alltoall(s, d, n, self.comm_id, self._res, self._res, run_async=self._async), inp, out, n * self.world, n * self.world)

    def send(self, t: torch.Tensor, dst, tag=0):
        sb, staged = self._buffer_of(t, "s")
        if staged:
            self._t(sb)[:t.numel()].copy_(t.reshape(-1))
        # always asynchronous: a blocking rendezvous send would wait for the peer's recv
        return self._track(self.accl.send(sb, t.numel(), dst, tag, self.comm_id, self._res, run_async=True))

    def recv(self, t: torch.Tensor, src, tag=0):
        db, staged = self._buffer_of(t, "d")
        req = self._track(self.accl.recv(db, t.numel(), src, tag, self.comm_id, self._res, run_async=self._async))
        if staged:
            if self._engine and req is not None:
                req.wait()
            t.copy_(self._t(db)[:t.numel()].view_as(t))
        return req

    def barrier(self):
        self.accl.barrier(self.comm_id)
        self.check(block=True)


class GradBucket:
    """A flat gradient bucket in the symmetric heap (DP all-reduce or ZeRO reduce_scatter)."""

    def __init__(self, group: TensorGroup, numel, dtype=torch.bfloat16):
        self.group = group
        self.numel = (numel + group.world - 1) // group.world * group.world
        self.flat = group.empty(self.numel, dtype=dtype)
        self.flat.zero_()

    def views(self, shapes):
        """Carve parameter-shaped gradient views out of the bucket."""
        out, off = [], 0
        for shp in shapes:
            n = 1
            for s in shp:
                n *= s
            out.append(self.flat[off:off + n].view(*shp))
            off += n
        return out

    def all_reduce(self, average=True):
        req = self.group.all_reduce(self.flat)
        if average:
            self.flat.div_(self.group.world)
        return req

    def reduce_scatter(self):
        shard = self.flat.view(self.group.world, -1)[self.group.rank]
        buf = self.flat._accl_buffer
        n = self.numel // self.group.world
        out = self.group.empty(n, dtype=self.flat.dtype)
        g = self.group
        g._track(g.accl.reduce_scatter(buf, out._accl_buffer, n, SUM, g.comm_id, g._res, g._res, run_async=g._async))
        del shard
        return out


class RowParallelLinear(torch.nn.Module):
    """y = reduce_scatter_over_ranks(x_r @ W_r^T): Megatron row-parallel linear with a sequence-parallel
    output shard.  On the CUDA backend the GEMM (tcgen05) emits its tiles straight into the owner rank's
    shard over NVLink (accl_b200.ops.gemm_reduce_scatter); elsewhere falls back to matmul + reduce_scatter."""

    def __init__(self, group: TensorGroup, in_features_per_rank, out_features, dtype=torch.bfloat16, device=None):
        super().__init__()
        self.group = group
        dev = device if device is not None else (torch.device("cuda", group.accl.cuda_device) if group.accl.is_cuda else "cpu")
        self.weight = torch.nn.Parameter(torch.empty(out_features, in_features_per_rank, dtype=dtype, device=dev))
        torch.nn.init.normal_(self.weight, std=0.02)
        self._out = None

    @torch.no_grad()
    def forward(self, x):
        from ..ops import gemm_reduce_scatter
        m = x.shape[0]
        P = self.group.world
        fused_ok = (self.group.accl.is_cuda and x.dtype == torch.bfloat16 and m % (128 * P) == 0 and
                    self.weight.shape[0] % 256 == 0 and self.weight.shape[1] % 64 == 0)
        if fused_ok:
            if self._out is None or self._out.length != m // P * self.weight.shape[0]:
                self._out = self.group.accl.create_buffer(m // P * self.weight.shape[0], torch.bfloat16)
            gemm_reduce_scatter(self.group.accl, x.contiguous(), self.weight.data, self._out)
            return self._out.dev.view(m // P, self.weight.shape[0])
        full = x @ self.weight.t()
        out = torch.empty(m // P, self.weight.shape[0], dtype=full.dtype, device=full.device)
        self.group.reduce_scatter_tensor(out, full)
        return out


def ring_exchange(group: TensorGroup, send_t: torch.Tensor, recv_t: torch.Tensor, tag=0):
    """One step of a ring schedule (context parallelism / ring attention): pass a block to the next
    rank while receiving the previous rank's block.  Even ranks send first (legal for both protocols)."""
    nxt, prv = (group.rank + 1) % group.world, (group.rank - 1) % group.world
    if group.rank % 2 == 0:
        s = group.send(send_t, nxt, tag)
        r = group.recv(recv_t, prv, tag)
    else:
        r = group.recv(recv_t, prv, tag)
        s = group.send(send_t, nxt, tag)
    return s, r
