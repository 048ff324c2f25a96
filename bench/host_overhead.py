"""Host-side cost of issuing one call (no GPU wait in the loop): where the small-message time goes."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402

a = A.cuda_rank(0, 1, 0, heap_mb=256)
a.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=16 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)
n = 1024
s, d = a.create_buffer(n), a.create_buffer(n)
s.dev.fill_(1.0)
N = 5000


def loop(fn, name):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(N):
        fn()
    host = (time.perf_counter() - t) / N * 1e6
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t) / N * 1e6
    print(f"{name:46s} host issue {host:6.2f} us/call   (incl. drain {tot:6.2f})", flush=True)


x = torch.ones(n, device="cuda")
loop(lambda: x.add_(1.0), "torch x.add_(1)  [reference: one torch kernel]")
loop(lambda: torch.cuda.current_stream(0).cuda_stream, "torch.cuda.current_stream().cuda_stream")
loop(lambda: a.impl.nop(True).free(), "_C nop async + free")
loop(lambda: a.nop(run_async=True).free(), "Accl.nop async + free (python wrapper)")
si, di = s.impl, d.impl
loop(lambda: a.impl.allreduce(si, di, n, A.SUM, 0, True, True, A.DataType.none, True).free(), "_C allreduce async + free")
loop(lambda: a.allreduce(s, d, n, A.SUM, from_fpga=True, to_fpga=True, run_async=True).free(), "Accl.allreduce async + free")
loop(lambda: a.allreduce(s, d, n, A.SUM, from_fpga=True, to_fpga=True).free(), "Accl.allreduce blocking")
