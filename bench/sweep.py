"""Collective sweep: accl_b200 vs NCCL in the same process, same sizes.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --out gpurun_out/sweep.csv

Device-timed (CUDA events), max over ranks, NCCL-tests bus-bandwidth factors.
Counterpart of the reference's ACCLSweepBenchmark (test/host/xrt/src/bench.cpp:25-61)
plus its parse_bench_results.py comparison — against NCCL instead of an alpha-beta model
(the alpha-beta model lives in accl_b200/models/cost_model.py).
"""
import argparse
import csv
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402


def factor(op, p):
    if p == 1:
        return 1.0
    return {"allreduce": 2.0 * (p - 1) / p, "allgather": (p - 1) / p, "reduce_scatter": (p - 1) / p,
            "bcast": 1.0, "reduce": 1.0, "scatter": (p - 1) / p, "gather": (p - 1) / p, "alltoall": (p - 1) / p,
            "sendrecv": 1.0}[op]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="allreduce,allgather,reduce_scatter")
    ap.add_argument("--min-log2", type=int, default=10)
    ap.add_argument("--max-log2", type=int, default=30)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--max-ctas", type=int, default=128)
    ap.add_argument("--egr-kb", type=int, default=4096, help="eager threshold in KiB: messages up to it use the one-way protocols")
    ap.add_argument("--slot-kb", type=int, default=64, help="eager slot (RX buffer) size in KiB")
    ap.add_argument("--batches", type=int, default=5, help="device-timed batches per point; the median is reported")
    ap.add_argument("--graph", action="store_true", help="add CUDA-graph replay columns (ours in direct mode, and NCCL)")
    ap.add_argument("--tune", default="", help="name=value,... runtime knobs, identical on every rank")
    ap.add_argument("--engine-workers", type=int, default=0)
    ap.add_argument("--compress", default="", help="wire dtype (float16 / bfloat16 / float8_e4m3): all-reduce / all-gather / "
                    "reduce-scatter run with compress_dtype; the NCCL column stays the uncompressed call")
    ap.add_argument("--oneshot-kb", type=int, default=2048)
    ap.add_argument("--nvls-min-ranks", type=int, default=3)
    ap.add_argument("--engine", action="store_true")
    ap.add_argument("--no-nccl", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = getattr(torch, args.dtype)
    esz = torch.empty((), dtype=dt).element_size()
    max_bytes = 1 << args.max_log2
    acc = A.cuda_rank(rank, world, local, heap_mb=(3 * max_bytes >> 20) + 768, max_ctas=args.max_ctas, engine=args.engine,
                      oneshot_kb=args.oneshot_kb, nvls_min_ranks=args.nvls_min_ranks, engine_workers=args.engine_workers)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=args.slot_kb << 10, max_egr_size=max(args.egr_kb, args.slot_kb) << 10,
                   max_rndzv_size=1 << 30)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        acc.set_tuning(k, int(v))
    if rank == 0:
        print("#", acc.describe(), flush=True)
    big_s = acc.create_buffer(max_bytes // esz, dt)
    big_d = acc.create_buffer(max_bytes // esz, dt)
    big_s.dev.fill_(1.0)
    nx = torch.ones(max_bytes // esz, dtype=dt, device="cuda") if not args.no_nccl and world > 1 else None
    ny = torch.empty(max_bytes // esz, dtype=dt, device="cuda") if nx is not None else None

    def timed(fn, iters, batches=None):
        """Median over `batches` device-timed batches (max over ranks each): a shared box shows
        occasional multi-ms hiccups, for NCCL and for us alike; the median keeps one from deciding a row."""
        batches = batches or args.batches
        for _ in range(3):
            fn()
        per_batch = max(2, iters // batches)
        out = []
        for _ in range(batches):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(per_batch):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / per_batch
            if world > 1:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            out.append(ms)
        return sorted(out)[len(out) // 2]

    def device_us(call, reps=15):
        """engine-measured duration of the call itself (completion record: first CTA in -> last CTA out, or
        fetched -> retired in engine mode) — the reference's PERFCNT / get_duration; median, max over ranks"""
        durs = []
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            r = call()
            r.wait()
            durs.append(r.duration_ns() * 1e-3)
            r.free()
        us = sorted(durs)[len(durs) // 2]
        if world > 1:
            t = torch.tensor([us], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t.item())
        return us

    def graph_us(fn, per_graph=20, replays=5):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn()
            st.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(per_graph):
                    fn()
            g.replay()
            st.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for _ in range(replays):
                g.replay()
            b.record(st)
            st.synchronize()
        ms = a.elapsed_time(b) / (replays * per_graph)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms * 1e3

    rows = []
    for op in args.ops.split(","):
        for lg in range(args.min_log2, args.max_log2 + 1, args.step):
            nbytes = 1 << lg
            n = nbytes // esz          # message size in the NCCL-tests sense (total for AG/RS)
            per = n // world           # per-rank block for AG / RS / scatter / gather
            iters = 100 if nbytes <= (1 << 20) else (40 if nbytes <= (1 << 26) else 15)
            kw = dict(from_fpga=True, to_fpga=True, run_async=True)
            if args.compress and op in ("allreduce", "allgather", "reduce_scatter"):
                kw["compress_dtype"] = args.compress if args.compress.startswith("float8") else getattr(torch, args.compress)
            s, d = big_s, big_d
            if op == "allreduce":
                c = lambda: acc.allreduce(s, d, n, A.SUM, **kw)
                g = (lambda: dist.all_reduce(nx[:n])) if nx is not None else None
            elif op == "allgather":
                if per == 0: continue
                c = lambda: acc.allgather(s, d, per, **kw)
                g = (lambda: dist.all_gather_into_tensor(ny[:per * world], nx[:per])) if nx is not None else None
            elif op == "reduce_scatter":
                if per == 0: continue
                c = lambda: acc.reduce_scatter(s, d, per, A.SUM, **kw)
                g = (lambda: dist.reduce_scatter_tensor(ny[:per], nx[:per * world])) if nx is not None else None
            elif op == "bcast":
                c = lambda: acc.bcast(s, n, 0, **kw)
                g = (lambda: dist.broadcast(nx[:n], 0)) if nx is not None else None
            elif op == "reduce":
                c = lambda: acc.reduce(s, d, n, 0, A.SUM, **kw)
                g = (lambda: dist.reduce(nx[:n], 0)) if nx is not None else None
            elif op == "scatter":
                if per == 0: continue
                c = lambda: acc.scatter(s, d, per, 0, **kw)
                sl = [nx[i * per:(i + 1) * per] for i in range(world)] if nx is not None and rank == 0 else None
                g = (lambda: dist.scatter(ny[:per], sl, src=0)) if nx is not None else None
            elif op == "gather":
                if per == 0: continue
                c = lambda: acc.gather(s, d, per, 0, **kw)
                gl = [ny[i * per:(i + 1) * per] for i in range(world)] if nx is not None and rank == 0 else None
                g = (lambda: dist.gather(nx[:per], gl, dst=0)) if nx is not None else None
            elif op == "alltoall":
                if per == 0: continue
                c = lambda: acc.alltoall(s, d, per, **kw)
                g = (lambda: dist.all_to_all_single(ny[:per * world], nx[:per * world])) if nx is not None else None
            else:
                continue
            f = lambda: c().free()  # noqa: E731
            ms = timed(f, iters)
            row = dict(op=op, bytes=nbytes, dtype=args.dtype, wire=args.compress or args.dtype, world=world,
                       mode="engine" if args.engine else "direct",
                       accl_us=ms * 1e3, accl_busbw=nbytes / ms * 1e-6 * factor(op, world),
                       accl_device_us=device_us(c) if nbytes <= (64 << 20) else None)
            if args.graph and not args.engine and nbytes <= (4 << 20):
                row["accl_graph_us"] = graph_us(f)
            if g is not None:
                ms_n = timed(g, iters)
                row.update(nccl_us=ms_n * 1e3, nccl_busbw=nbytes / ms_n * 1e-6 * factor(op, world), speedup=ms_n / ms)
                if args.graph and nbytes <= (4 << 20):
                    row["nccl_graph_us"] = graph_us(g)
                    if "accl_graph_us" in row:
                        row["speedup_graph"] = row["nccl_graph_us"] / row["accl_graph_us"]
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
                if args.out:  # survive a timeout: every row is on disk as soon as it is measured
                    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
                    with open(args.out + ".jsonl", "a") as jf:
                        jf.write(json.dumps(row) + "\n")
    if rank == 0 and args.out and rows:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        keys = sorted({k for r in rows for k in r}, key=lambda k: (k not in ("op", "bytes"), k))
        with open(args.out, "w", newline="") as fh:
            w = csv.DictWriter(fh, fieldnames=keys)
            w.writeheader()
            w.writerows(rows)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
