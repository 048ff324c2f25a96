"""Deterministic sequence of collectives for an `ncu` capture on >= 2 GPUs (one kernel per call, direct launch).

  ncu --replay-mode application --target-processes all --clock-control none -k regex:k_call \\
      --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,... \\
      --csv --log-file gpurun_out/ncu_coll.csv \\
      python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 bench/ncu_target.py

Application replay re-runs the whole job per pass, so both ranks stay in lockstep (kernel replay would re-run one
rank's kernel against flags its peer already raised).  The list of calls, in launch order, is written next to the
capture (`--plan`), bench/ncu_summary.py joins the two.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plan", default="gpurun_out/ncu_coll_plan.json")
    ap.add_argument("--big-mb", type=int, default=256)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    big = a.big_mb << 20
    acc = A.cuda_rank(rank, world, local, heap_mb=(3 * big >> 20) + 768, max_ctas=128)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=64 << 10, max_egr_size=4 << 20, max_rndzv_size=1 << 30)
    n = big // 4
    s, d = acc.create_buffer(n), acc.create_buffer(n)
    s.dev.normal_()
    kw = dict(from_fpga=True, to_fpga=True)
    P = world
    calls = []

    def add(name, nbytes, fn):
        calls.append((name, nbytes, fn))

    for nb in (1 << 10, 64 << 10, 1 << 20, 16 << 20, big):
        add("allreduce", nb, lambda nb=nb: acc.allreduce(s, d, nb // 4, A.SUM, **kw))
    for nb in (64 << 10, 16 << 20, big):
        add("allgather", nb, lambda nb=nb: acc.allgather(s, d, nb // 4 // P, **kw))
        add("reduce_scatter", nb, lambda nb=nb: acc.reduce_scatter(s, d, nb // 4 // P, A.SUM, **kw))
    add("bcast", 16 << 20, lambda: acc.bcast(s, (16 << 20) // 4, 0, **kw))
    add("reduce", 16 << 20, lambda: acc.reduce(s, d, (16 << 20) // 4, 0, A.SUM, **kw))
    add("allreduce bf16 wire", big, lambda: acc.allreduce(s, d, n, A.SUM, compress_dtype=torch.bfloat16, **kw))
    plan = []
    for rep in range(2):  # first round warms up (and is captured too: same kernels)
        for name, nb, fn in calls:
            fn()
            plan.append(dict(op=name, bytes=nb, rep=rep))
    torch.cuda.synchronize()
    acc.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(a.plan) or ".", exist_ok=True)
        json.dump(dict(world=world, calls=plan, describe=acc.describe()), open(a.plan, "w"))
    acc.deinit()


if __name__ == "__main__":
    main()
