"""Fused compute+collective plugins (CUDA backend).

* gemm_reduce_scatter: row-parallel linear layer — every rank multiplies its
  K-slice on the tcgen05 tensor cores and each finished tile is added straight
  into the owner rank's output shard over NVLink (no intermediate C, no
  separate reduce_scatter kernel).
* vadd_allreduce: the vector-add plugin; the kernel computes x + y chunk by
  chunk and hands every finished chunk to the persistent engine itself
  (device API, all_reduce_async) while it computes the next one.
* vadd_put / stream_pull: the reference's vadd_put example — x + 1 pushed into
  a stream of the destination rank while computing; the consumer drains it.
"""
import torch

from .. import _C
from ..core import Accl, Buffer


def _stream_handle(accl):
    h = torch.cuda.current_stream(accl.cuda_device).cuda_stream
    return h if h else 1  # cudaStreamLegacy


def gemm_reduce_scatter(accl: Accl, a: torch.Tensor, w: torch.Tensor, out: Buffer = None, variant: int = 0,
                        out_dtype=torch.bfloat16) -> Buffer:
    """out[M/P, N] (in the symmetric heap) = reduce_scatter over ranks of a[M, K_r] @ w[N, K_r]^T.

    `a` and `w` are this rank's K-slices (bf16, contiguous, on this rank's GPU).
    Needs M % (128 * world) == 0, N % 256 == 0, K_r % 64 == 0.  Stream ordered.
    The shard is bf16 (partials added in bf16 by the TMA unit) or, with an fp32 `out` buffer / out_dtype, fp32
    (the P partial products are accumulated without intermediate rounding).  variant: 0 automatic, 1 one CTA per
    128 x 256 tile, 2 CTA pair (tcgen05 cta_group::2) per 256 x 256 tile (needs M % (256 * world) == 0).
    """
    assert a.is_cuda and w.is_cuda and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.is_contiguous() and w.is_contiguous() and a.shape[1] == w.shape[1]
    M, K = a.shape
    N = w.shape[0]
    P = accl.world
    if out is None:
        out = accl.create_buffer(M // P * N, out_dtype)
    assert out.length >= M // P * N
    _C.gemm_reduce_scatter(accl.impl, a.data_ptr(), w.data_ptr(), out.impl, M, N, K, _stream_handle(accl), variant)
    return out


def vadd_allreduce(accl: Accl, x: Buffer, y: Buffer, out: Buffer, tmp: Buffer = None, count: int = None,
                   chunk_elems: int = 0):
    """out = allreduce_sum(x + y) (fp32), produced and reduced in chunks of `chunk_elems` (0: 1 Mi elements).
    Returns a 1-element int32 CUDA tensor that receives the engine's status word (0 = success) once the kernel
    has finished."""
    count = x.length if count is None else count
    if tmp is None:
        tmp = accl.create_buffer(count, torch.float32)
    status = torch.full((1,), -1, dtype=torch.int32, device=torch.device("cuda", accl.cuda_device))
    _C.vadd_allreduce(accl.impl, x.impl, y.impl, tmp.impl, out.impl, count, status.data_ptr(), _stream_handle(accl),
                      chunk_elems)
    status._keep = tmp
    return status


def vadd_put(accl: Accl, src: Buffer, count: int, dst_rank: int, stream_id: int = 9):
    """The reference's vadd_put user kernel: src + 1 (fp32) is pushed tile by tile into stream `stream_id` of rank
    `dst_rank` while it is being computed (device::Data::push).  Returns the status tensor."""
    status = torch.full((1,), -1, dtype=torch.int32, device=torch.device("cuda", accl.cuda_device))
    _C.vadd_put(accl.impl, src.impl, count, dst_rank, stream_id, status.data_ptr(), _stream_handle(accl))
    return status


def stream_pull(accl: Accl, dst: Buffer, count: int, stream_id: int = 9):
    """Drain `count` fp32 of stream `stream_id` of this rank into `dst` (device::Data::pull).  Returns the status tensor."""
    status = torch.full((1,), -1, dtype=torch.int32, device=torch.device("cuda", accl.cuda_device))
    _C.stream_pull(accl.impl, dst.impl, count, stream_id, status.data_ptr(), _stream_handle(accl))
    return status


def stream_loopback(accl: Accl, count: int, add_one: bool = False, scratch: Buffer = None):
    """Run the loopback user kernel on this rank's stream port: pull `count` fp32 words that a previous
    `copy_to_stream` / `recv_to_stream` / `RES_STREAM` op produced, optionally add one, push them back for a
    following `copy_from_stream` / `send_from_stream` / `OP0_STREAM` op (reference
    kernels/plugins/loopback, test/host/hls_simulator/test.cpp:153).  Returns the status tensor."""
    if scratch is None:
        scratch = accl.create_buffer(count, torch.float32)
    status = torch.full((1,), -1, dtype=torch.int32, device=torch.device("cuda", accl.cuda_device))
    _C.stream_loopback(accl.impl, scratch.impl, count, add_one, status.data_ptr(), _stream_handle(accl))
    status._keep = scratch
    return status
