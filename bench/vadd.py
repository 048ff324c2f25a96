"""Vector-add plugin -> all-reduce (BASELINE config #4): a compute kernel that issues the collective itself
through the device API, vs the unfused baseline (torch add kernel + NCCL all_reduce).

  python -m torch.distributed.run --nproc-per-node N ... bench/vadd.py [--min-log2 12 --max-log2 26]

Fused: one `k_plugin_vadd_allreduce` launch per step.  The kernel produces x + y chunk by chunk; the CTA that
completes a chunk takes a ticket in the resident engine's command ring (`accl::device::Command::all_reduce_async`)
and the kernel goes on computing while the engine's control CTA plans the call and its worker CTAs reduce the
chunk over NVLink — no host on the path after the launch, compute and communication overlap.  Device-timed with
CUDA events, max over ranks.  Note: the engine mode needs `CUDA_DEVICE_MAX_CONNECTIONS >= 2`; the launcher sets 32.
"""
import argparse
import faulthandler
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402
from accl_b200.ops import vadd_allreduce  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=12)
    ap.add_argument("--max-log2", type=int, default=26)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--max-ctas", type=int, default=64)
    ap.add_argument("--engine-workers", type=int, default=48)
    ap.add_argument("--chunk-kb", type=int, default=0, help="bytes of x + y handed to the engine per device-issued all-reduce")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", 90)), exit=True)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nmax = (1 << a.max_log2) // 4
    acc = A.cuda_rank(rank, world, local, heap_mb=max(768, (16 * nmax >> 20) + 512), max_ctas=a.max_ctas, engine=True,
                      engine_workers=a.engine_workers)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=64 << 10, max_egr_size=1 << 20, max_rndzv_size=1 << 30)
    if rank == 0:
        print("#", acc.describe(), flush=True)
    x, y, tmp, out = (acc.create_buffer(nmax, torch.float32) for _ in range(4))
    x.dev.fill_(1.0)
    y.dev.fill_(float(rank))
    tx, ty = torch.ones(nmax, device="cuda"), torch.full((nmax,), float(rank), device="cuda")

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    rows = []
    for lg in range(a.min_log2, a.max_log2 + 1, a.step):
        n = (1 << lg) // 4
        statuses = []

        def fused():
            statuses.append(vadd_allreduce(acc, x, y, out, tmp, n, chunk_elems=(a.chunk_kb << 10) // 4))
            if len(statuses) > 8:
                statuses.pop(0)

        def baseline():
            z = tx[:n] + ty[:n]
            if world > 1:
                dist.all_reduce(z)
            return z

        fused()
        torch.cuda.synchronize()
        expect = float(world + sum(range(world)))
        ok = bool(torch.all(out.dev[:n] == expect)) and int(statuses[-1].item()) == 0
        iters = a.iters if lg <= 22 else max(5, a.iters // 5)
        ms_f = timed(fused, iters)
        ms_b = timed(baseline, iters)
        row = dict(op="vadd_allreduce", bytes=n * 4, world=world, fused_us=ms_f * 1e3, torch_nccl_us=ms_b * 1e3,
                   speedup=ms_b / ms_f, correct=ok)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0 and a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "a") as fh:
            for r in rows:
                fh.write(json.dumps(r) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
