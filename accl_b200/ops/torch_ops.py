"""`torch.ops.accl_b200.*`: the collectives as registered torch custom ops, dispatched to the calling
thread's default `TensorGroup` (torch surface / drop-in comparison against
`torch.distributed`).  In-place on the output tensor, stream ordered on CUDA.

    from accl_b200.parallel import TensorGroup
    from accl_b200.ops import torch_ops
    torch_ops.set_default_group(TensorGroup(accl))
    torch.ops.accl_b200.all_reduce(t, "sum")
    torch.ops.accl_b200.all_gather(out, shard)
"""
import threading

import torch

from ..core import MAX, SUM

_tls = threading.local()
_OPS = {"sum": SUM, "max": MAX}


def set_default_group(group):
    """Bind the `TensorGroup` used by torch.ops.accl_b200.* on this thread (one rank per thread or process)."""
    _tls.group = group


def default_group():
    g = getattr(_tls, "group", None)
    if g is None:
        raise RuntimeError("accl_b200: no default group on this thread; call torch_ops.set_default_group(TensorGroup(accl))")
    return g


@torch.library.custom_op("accl_b200::all_reduce", mutates_args=("t",))
def all_reduce(t: torch.Tensor, op: str = "sum") -> None:
    default_group().all_reduce(t, _OPS[op])


@torch.library.custom_op("accl_b200::broadcast", mutates_args=("t",))
def broadcast(t: torch.Tensor, root: int = 0) -> None:
    default_group().broadcast(t, root)


@torch.library.custom_op("accl_b200::all_gather", mutates_args=("out",))
def all_gather(out: torch.Tensor, inp: torch.Tensor) -> None:
    default_group().all_gather_into_tensor(out, inp)


@torch.library.custom_op("accl_b200::reduce_scatter", mutates_args=("out",))
def reduce_scatter(out: torch.Tensor, inp: torch.Tensor, op: str = "sum") -> None:
    default_group().reduce_scatter_tensor(out, inp, _OPS[op])


@torch.library.custom_op("accl_b200::all_to_all", mutates_args=("out",))
def all_to_all(out: torch.Tensor, inp: torch.Tensor) -> None:
    default_group().all_to_all_single(out, inp)


@torch.library.custom_op("accl_b200::barrier", mutates_args=())
def barrier() -> None:
    default_group().barrier()
