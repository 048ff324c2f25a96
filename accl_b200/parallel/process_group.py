"""`torch.distributed` backend "accl": the collectives of this library behind the standard ProcessGroup
interface, so existing `torch.distributed` code (including DistributedDataParallel) runs on it unchanged.

    import accl_b200.parallel.process_group          # registers the backend
    dist.init_process_group("accl", init_method="tcp://127.0.0.1:29500", rank=r, world_size=w)
    dist.all_reduce(t)

One rank per process.  The engine is chosen like `init_from_env`: the CUDA backend when a GPU and the sm_100a
extension are available (tensors must then be CUDA tensors; calls are stream ordered and their Work objects
complete immediately, like NCCL's), otherwise the CPU emulator over loopback TCP (calls are blocking).
SURVEY 7.2 step 9 ("optional ProcessGroup shim for drop-in comparison").
"""
import os

import torch
import torch.distributed as dist
from torch._C._distributed_c10d import (AllgatherOptions, AllreduceCoalescedOptions, AllreduceOptions, AllToAllOptions,
                                        BarrierOptions, BroadcastOptions, GatherOptions, ReduceOptions,
                                        ReduceScatterOptions, ScatterOptions, _create_work_from_future)
from torch.futures import Future

from ..core import MAX, SUM
from . import TensorGroup, init_from_env


def _done(result):
    fut = Future()
    fut.set_result(result)
    return _create_work_from_future(fut)


class AcclProcessGroup(dist.ProcessGroup):
    def __init__(self, rank, world_size, accl):
        super().__init__(rank, world_size)
        self._rank, self._world = rank, world_size
        self.accl = accl
        self.group = TensorGroup(accl)

    # -- identity ------------------------------------------------------------
    def getBackendName(self):
        return "accl"

    def size(self):
        return self._world

    def __repr__(self):
        return f"AcclProcessGroup(rank={self._rank}, world={self._world}, {self.accl.describe()})"

    # -- helpers ----------------------------------------------------------------
    @staticmethod
    def _op(opts):
        op = getattr(opts, "reduceOp", dist.ReduceOp.SUM)
        if op == dist.ReduceOp.SUM:
            return SUM, None
        if op == dist.ReduceOp.MAX:
            return MAX, None
        if op == dist.ReduceOp.AVG:
            return SUM, "avg"
        raise NotImplementedError(f"accl backend: reduce op {op} (SUM, MAX and AVG are supported)")

    def _finish(self, t, post):
        if post == "avg":
            t.div_(self._world)

    # -- collectives ----------------------------------------------------------------
    def allreduce(self, tensor_list, opts=AllreduceOptions()):
        fn, post = self._op(opts)
        for t in tensor_list:
            self.group.all_reduce(t, fn)
            self._finish(t, post)
        return _done(tensor_list)

    def allreduce_coalesced(self, tensor_list, opts=AllreduceCoalescedOptions()):
        return self.allreduce(tensor_list, opts)

    def broadcast(self, tensor_list, opts=BroadcastOptions()):
        for t in tensor_list:
            self.group.broadcast(t, opts.rootRank)
        return _done(tensor_list)

    def reduce(self, tensor_list, opts=ReduceOptions()):
        fn, post = self._op(opts)
        for t in tensor_list:
            n = t.numel()
            flat = t.reshape(-1) if t.is_contiguous() else t.contiguous().view(-1)
            sb, _ = self.group._buffer_of(flat, "s")
            db, _ = self.group._buffer_of(flat, "d")
            self.group._t(sb)[:n].copy_(flat)
            self.accl.reduce(sb, db, n, opts.rootRank, fn, self.group.comm_id, self.group._res, self.group._res,
                             run_async=self.group._async)
            if self._rank == opts.rootRank:
                t.copy_(self.group._t(db)[:n].view_as(t))
                self._finish(t, post)
        return _done(tensor_list)

    def _allgather_base(self, output_tensor, input_tensor, opts=AllgatherOptions()):
        self.group.all_gather_into_tensor(output_tensor, input_tensor.contiguous())
        return _done(output_tensor)

    def allgather(self, output_tensors, input_tensor, opts=AllgatherOptions()):
        for outs, inp in zip(output_tensors, input_tensor):
            flat = torch.empty(self._world * inp.numel(), dtype=inp.dtype, device=inp.device)
            self.group.all_gather_into_tensor(flat, inp.contiguous())
            for o, chunk in zip(outs, flat.chunk(self._world)):
                o.copy_(chunk.view_as(o))
        return _done(output_tensors)

    def allgather_into_tensor_coalesced(self, output_tensor_list, input_tensor_list, opts=AllgatherOptions()):
        for o, i in zip(output_tensor_list, input_tensor_list):
            self._allgather_base(o, i, opts)
        return _done(output_tensor_list)

    def _reduce_scatter_base(self, output_tensor, input_tensor, opts=ReduceScatterOptions()):
        fn, post = self._op(opts)
        self.group.reduce_scatter_tensor(output_tensor, input_tensor.contiguous(), fn)
        self._finish(output_tensor, post)
        return _done(output_tensor)

    def reduce_scatter(self, output_tensors, scatter_lists, opts=ReduceScatterOptions()):
        for out, parts in zip(output_tensors, scatter_lists):
            self._reduce_scatter_base(out, torch.cat([p.reshape(-1) for p in parts]), opts)
        return _done(output_tensors)

    def reduce_scatter_tensor_coalesced(self, output_tensors, input_tensors, opts=ReduceScatterOptions()):
        for o, i in zip(output_tensors, input_tensors):
            self._reduce_scatter_base(o, i, opts)
        return _done(output_tensors)

    def alltoall_base(self, output_buffer, input_buffer, output_split_sizes, input_split_sizes, opts=AllToAllOptions()):
        if output_split_sizes or input_split_sizes:
            raise NotImplementedError("accl backend: all_to_all_single with uneven splits")
        self.group.all_to_all_single(output_buffer, input_buffer.contiguous())
        return _done(output_buffer)

    def alltoall(self, output_tensor_list, input_tensor_list, opts=AllToAllOptions()):
        inp = torch.cat([t.reshape(-1) for t in input_tensor_list])
        out = torch.empty_like(inp)
        self.group.all_to_all_single(out, inp)
        for o, chunk in zip(output_tensor_list, out.chunk(self._world)):
            o.copy_(chunk.view_as(o))
        return _done(output_tensor_list)

    def gather(self, output_tensors, input_tensors, opts=GatherOptions()):
        inp = input_tensors[0].contiguous()
        n = inp.numel()
        sb, _ = self.group._buffer_of(inp.view(-1), "s")
        db = self.accl.create_buffer(n * self._world, inp.dtype)
        self.group._t(sb)[:n].copy_(inp.view(-1))
        self.accl.gather(sb, db, n, opts.rootRank, self.group.comm_id, self.group._res, self.group._res,
                         run_async=self.group._async)
        if self._rank == opts.rootRank:
            for o, chunk in zip(output_tensors[0], self.group._t(db).chunk(self._world)):
                o.copy_(chunk.view_as(o))
        return _done(output_tensors)

    def scatter(self, output_tensors, input_tensors, opts=ScatterOptions()):
        out = output_tensors[0]
        n = out.numel()
        sb = self.accl.create_buffer(n * self._world, out.dtype)
        db, _ = self.group._buffer_of(out.reshape(-1), "d")
        if self._rank == opts.rootRank:
            self.group._t(sb).copy_(torch.cat([t.reshape(-1) for t in input_tensors[0]]))
        self.accl.scatter(sb, db, n, opts.rootRank, self.group.comm_id, self.group._res, self.group._res,
                          run_async=self.group._async)
        out.copy_(self.group._t(db)[:n].view_as(out))
        return _done(output_tensors)

    def barrier(self, opts=BarrierOptions()):
        self.group.barrier()
        return _done(None)

    # -- point to point ----------------------------------------------------------------
    def send(self, tensors, dst, tag=0):
        reqs = [self.group.send(t.contiguous(), dst, tag) for t in tensors]
        for r in reqs:      # asynchronous issue (a rendezvous send completes when the peer has posted its recv)
            r.wait()
        return _done(None)

    def recv(self, tensors, src, tag=0):
        for t in tensors:
            self.group.recv(t, src, tag)
        return _done(tensors)


def _create(store, rank, world_size, timeout):
    os.environ.setdefault("RANK", str(rank))
    os.environ.setdefault("WORLD_SIZE", str(world_size))
    accl = init_from_env(None)
    assert accl.rank == rank and accl.world == world_size, "accl backend: RANK / WORLD_SIZE disagree with init_process_group"
    return AcclProcessGroup(rank, world_size, accl)


dist.Backend.register_backend("accl", _create, devices=["cpu", "cuda"])
