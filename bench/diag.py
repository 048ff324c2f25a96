"""Where does a small call's time go?  Per (op, size): host-loop CUDA-event time of back-to-back asynchronous
calls, the device-measured duration of the same call (completion record, `get_duration` — the reference's
PERFCNT, test/host/xrt/include/fixture.hpp:134-152) and NCCL through torch.distributed, for the direct-launch
and the persistent-engine execution modes.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 bench/diag.py
"""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="nop,allreduce,allgather,reduce_scatter")
    ap.add_argument("--sizes", default="1024,16384,65536,262144,1048576,4194304")
    ap.add_argument("--modes", default="direct,engine")
    ap.add_argument("--max-ctas", type=int, default=64)
    ap.add_argument("--egr-kb", type=int, default=64, help="eager threshold (one-way protocols up to it)")
    ap.add_argument("--slot-kb", type=int, default=64, help="eager slot size")
    ap.add_argument("--nvls-min-ranks", type=int, default=3)
    ap.add_argument("--big-mb", type=int, default=0, help="also time one all-reduce of this size (MiB)")
    ap.add_argument("--graph", action="store_true", help="also time CUDA-graph replays (direct mode and NCCL)")
    ap.add_argument("--engine-workers", type=int, default=32)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = [int(s) for s in args.sizes.split(",")]
    maxb = max(max(sizes), args.big_mb << 20)

    def graph_time(fn, per_graph=20, replays=6):
        """us per call when `per_graph` calls are captured once and replayed (launch-bound loops belong in a graph)"""
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

    def ev_time(fn, iters):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms * 1e3

    rows = []
    fh = None
    if rank == 0 and args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        fh = open(args.out, "w")

    def emit(row):
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
            if fh:
                fh.write(json.dumps(row) + "\n")
                fh.flush()

    port = int(os.environ.get("MASTER_PORT", 29500)) + 137
    for mi, mode in enumerate(args.modes.split(",")):
        acc = A.cuda_rank(rank, world, local, port=port + 11 * mi, heap_mb=(3 * maxb >> 20) + 256, max_ctas=args.max_ctas,
                          engine=(mode == "engine"), nvls_min_ranks=args.nvls_min_ranks, engine_workers=args.engine_workers)
        acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=args.slot_kb << 10, max_egr_size=max(args.egr_kb, args.slot_kb) << 10, max_rndzv_size=1 << 30)
        if rank == 0:
            print("#", acc.describe(), flush=True)
        n_max = maxb // 4
        s = acc.create_buffer(n_max, torch.float32)
        d = acc.create_buffer(n_max, torch.float32)
        s.dev.fill_(1.0)
        nx = torch.ones(n_max, device="cuda")
        ny = torch.empty(n_max, device="cuda")
        kw = dict(from_fpga=True, to_fpga=True)
        for op in args.ops.split(","):
            for nbytes in ([0] if op == "nop" else sizes):
                n = nbytes // 4
                per = n // world
                if op == "nop":
                    call = lambda **k: acc.nop(**k)  # noqa: E731
                    ref = None
                elif op == "allreduce":
                    call = lambda **k: acc.allreduce(s, d, n, A.SUM, **kw, **k)  # noqa: E731
                    ref = lambda: dist.all_reduce(nx[:n])  # noqa: E731
                elif op == "allgather":
                    call = lambda **k: acc.allgather(s, d, per, **kw, **k)  # noqa: E731
                    ref = lambda: dist.all_gather_into_tensor(ny[:per * world], nx[:per])  # noqa: E731
                elif op == "reduce_scatter":
                    call = lambda **k: acc.reduce_scatter(s, d, per, A.SUM, **kw, **k)  # noqa: E731
                    ref = lambda: dist.reduce_scatter_tensor(ny[:per], nx[:per * world])  # noqa: E731
                else:
                    continue
                iters = 200 if nbytes <= (1 << 20) else 50
                us_async = ev_time(lambda: call(run_async=True).free(), iters)
                durs = []
                for _ in range(40):
                    if world > 1:
                        dist.barrier()
                    r = call(run_async=True)
                    r.wait()
                    durs.append(r.duration_ns() * 1e-3)
                    r.free()
                row = dict(mode=mode, op=op, bytes=nbytes, world=world, us_event_async=round(us_async, 2),
                           us_device_median=round(statistics.median(durs), 2), us_device_min=round(min(durs), 2))
                if args.graph and mode == "direct" and op != "nop":
                    row["us_graph"] = round(graph_time(lambda: call(run_async=True).free()), 2)
                if ref is not None and world > 1 and mi == 0:
                    row["us_nccl"] = round(ev_time(ref, iters), 2)
                    if args.graph:
                        row["us_nccl_graph"] = round(graph_time(ref), 2)
                emit(row)
        if args.big_mb:
            n = (args.big_mb << 20) // 4
            us = ev_time(lambda: acc.allreduce(s, d, n, A.SUM, run_async=True, **kw).free(), 10)
            f = 2.0 * (world - 1) / world if world > 1 else 1.0
            row = dict(mode=mode, op="allreduce", bytes=args.big_mb << 20, world=world, us_event_async=round(us, 1),
                       busbw=round((args.big_mb << 20) / us * 1e-3 * f, 1))
            if world > 1 and mi == 0:
                un = ev_time(lambda: dist.all_reduce(nx[:n]), 10)
                row.update(us_nccl=round(un, 1), nccl_busbw=round((args.big_mb << 20) / un * 1e-3 * f, 1))
            emit(row)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        del s, d
        acc.deinit()
    if fh:
        fh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
