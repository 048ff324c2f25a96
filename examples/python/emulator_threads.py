"""Four emulator ranks as threads: every collective once, through the Python API.

    python examples/python/emulator_threads.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import accl_b200 as A  # noqa: E402


def body(accl, rank, world):
    n = 1000
    src, dst = accl.create_buffer(n), accl.create_buffer(n * world)
    src.host[:] = rank + 1
    accl.allreduce(src, dst, n, A.SUM)
    assert torch.all(dst.host[:n] == world * (world + 1) / 2)
    accl.allgather(src, dst, n)
    assert torch.equal(dst.host[::n], torch.arange(1, world + 1, dtype=torch.float32))
    accl.bcast(src, n, root=2)
    assert torch.all(src.host == 3)
    accl.reduce(src, dst, n, root=0, func=A.MAX)
    req = accl.allreduce(src, dst, n, A.SUM, compress_dtype=torch.float16, run_async=True)   # fp16 on the wire
    req.wait()
    dst.sync_from_device()
    ns = req.duration_ns()
    req.free()
    accl.barrier()
    return f"rank {rank}: ok, last all-reduce took {ns / 1e3:.0f} us in the engine"


if __name__ == "__main__":
    for line in A.run_ranks(4, body, dict(n_egr_rx_bufs=16, egr_rx_buf_size=1024, max_egr_size=1024, max_rndzv_size=1 << 20)):
        print(line)
