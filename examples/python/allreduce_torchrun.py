"""One process per B200 (torchrun): zero-copy heap tensors, stream-ordered all-reduce, torch custom ops.

    torchrun --nproc-per-node 8 examples/python/allreduce_torchrun.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import accl_b200 as A  # noqa: E402
from accl_b200.ops import torch_ops  # noqa: E402
from accl_b200.parallel import TensorGroup, init_from_env  # noqa: E402

accl = init_from_env("cuda", heap_mb=2048, max_ctas=128)
group = TensorGroup(accl)
torch_ops.set_default_group(group)

x = group.empty(64 << 20, dtype=torch.bfloat16)          # lives in the symmetric heap: no staging copy
x.fill_(1.0)
torch.ops.accl_b200.all_reduce(x, "sum")                 # stream ordered on torch's current stream
torch.cuda.synchronize()
assert float(x[0]) == group.world

src, dst = accl.create_buffer(1 << 26, torch.float32), accl.create_buffer(1 << 26, torch.float32)
src.dev.fill_(float(group.rank))
req = accl.allreduce(src, dst, 1 << 26, A.SUM, from_fpga=True, to_fpga=True, run_async=True)
req.wait()
if group.rank == 0:
    nbytes = 4 << 26
    us = req.duration_ns() / 1e3
    print(f"{group.world} GPUs: 256 MiB all-reduce in {us:.0f} us = {nbytes / us * 1e-3 * 2 * (group.world - 1) / group.world:.0f} GB/s bus bandwidth")
req.free()
accl.barrier()
accl.deinit()
