"""NCCL baseline sweep (the bar to beat, BASELINE.json north_star).

Run under torchrun:  python -m torch.distributed.run --nproc-per-node N \
    --master-addr 127.0.0.1 --master-port 29511 bench/nccl_baseline.py [--out csv]

Device-timed with CUDA events, max over ranks, NCCL-tests bus-bandwidth
convention (allreduce 2(P-1)/P, allgather / reduce_scatter (P-1)/P).
Mirrors the reference's sweep benchmark (test/host/xrt/src/bench.cpp:25-61)
with NCCL in the role of the library under test.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def busbw_factor(op, p):
    if p == 1:
        return 1.0
    return {"allreduce": 2.0 * (p - 1) / p, "allgather": (p - 1) / p,
            "reduce_scatter": (p - 1) / p, "bcast": 1.0, "reduce": 1.0,
            "alltoall": (p - 1) / p, "sendrecv": 1.0}[op]


def time_op(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--ops", default="allreduce,allgather,reduce_scatter")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    p = dist.get_world_size()
    dt = getattr(torch, args.dtype)
    esz = torch.empty((), dtype=dt).element_size()
    rows = []
    for op in args.ops.split(","):
        for lg in range(args.min_log2, args.max_log2 + 1, args.step):
            nbytes = 1 << lg
            n = nbytes // esz
            iters = 200 if nbytes <= (1 << 20) else (20 if nbytes <= (1 << 26) else 5)
            if op == "allreduce":
                x = torch.ones(n, dtype=dt, device="cuda")
                fn = lambda: dist.all_reduce(x)
            elif op == "allgather":
                x = torch.ones(n // p, dtype=dt, device="cuda")
                y = torch.empty(n // p * p, dtype=dt, device="cuda")
                fn = lambda: dist.all_gather_into_tensor(y, x)
            elif op == "reduce_scatter":
                x = torch.ones(n // p * p, dtype=dt, device="cuda")
                y = torch.empty(n // p, dtype=dt, device="cuda")
                fn = lambda: dist.reduce_scatter_tensor(y, x)
            elif op == "bcast":
                x = torch.ones(n, dtype=dt, device="cuda")
                fn = lambda: dist.broadcast(x, 0)
            elif op == "reduce":
                x = torch.ones(n, dtype=dt, device="cuda")
                fn = lambda: dist.reduce(x, 0)
            else:
                continue
            ms = time_op(fn, iters)
            alg = nbytes / ms * 1e-6
            row = dict(impl="nccl", op=op, bytes=nbytes, dtype=args.dtype, world=p,
                       us=ms * 1e3, algbw_GBps=alg, busbw_GBps=alg * busbw_factor(op, p))
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if rank == 0 and args.out:
        import csv
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w", newline="") as f:
            w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
            w.writeheader()
            w.writerows(rows)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
