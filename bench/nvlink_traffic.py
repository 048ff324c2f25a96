"""NVLink byte counters vs algorithmic bytes: `nvidia-smi nvlink -gt d` is read before and after a batch of calls, the
delta (sum over the 18 links of this rank's GPU, Tx and Rx) is compared with what the algorithm has to move.
Evidence that the fused paths move what they claim and nothing more (north-star: "bus bandwidth against 900 GB/s per
direction").

  torchrun --nproc-per-node N bench/nvlink_traffic.py --out gpurun_out/nvlink_traffic.jsonl
"""
import argparse
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402
from accl_b200.ops import gemm_reduce_scatter  # noqa: E402


def counters(gpu):
    """(tx_bytes, rx_bytes) summed over the links of one GPU; None when the tool does not report them"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu)], capture_output=True, text=True, timeout=20).stdout
    except Exception:  # noqa: BLE001
        return None
    tx = rx = 0
    found = False
    for m in re.finditer(r"Data (Tx|Rx):\s*([0-9]+)\s*(KiB|MiB|GiB|KB|MB|GB|B)?", out):
        mult = {"KiB": 1 << 10, "MiB": 1 << 20, "GiB": 1 << 30, "KB": 1000, "MB": 10 ** 6, "GB": 10 ** 9, "B": 1, None: 1 << 10}[m.group(3)]
        v = int(m.group(2)) * mult
        if m.group(1) == "Tx":
            tx += v
        else:
            rx += v
        found = True
    return (tx, rx) if found else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = a.mb << 20
    n = nbytes // 4
    acc = A.cuda_rank(rank, world, local, heap_mb=(4 * nbytes >> 20) + 1024, max_ctas=128)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=64 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)
    s, d = acc.create_buffer(n), acc.create_buffer(n)
    s.dev.normal_()
    kw = dict(from_fpga=True, to_fpga=True, run_async=True)
    P = world
    per = n // P
    M, N, K = 8192, 8192, 2048
    x = (torch.randn(M, K, device="cuda") * 0.25).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.25).bfloat16()
    shard = acc.create_buffer(M // P * N, torch.bfloat16)
    nvls = "nvls=yes" in acc.describe() and P >= 3
    cases = [
        ("allreduce fp32 (NVLS two-shot)" if nvls else "allreduce fp32 (peer two-shot)",
         lambda: acc.allreduce(s, d, n, A.SUM, **kw).free(),
         nbytes * (1 + 1 / P) if nvls else 2 * nbytes * (P - 1) / P, nbytes * (1 + 1 / P) if nvls else 2 * nbytes * (P - 1) / P),
        ("reduce_scatter fp32 (peer pull)", lambda: acc.reduce_scatter(s, d, per, A.SUM, **kw).free(),
         per * 4 * (P - 1), per * 4 * (P - 1)),
        ("allgather fp32 (peer push)", lambda: acc.allgather(s, d, per, **kw).free(), per * 4 * (P - 1), per * 4 * (P - 1)),
        ("allreduce fp32, bf16 wire (compressed two-shot)", lambda: acc.allreduce(s, d, n, A.SUM, compress_dtype=torch.bfloat16, **kw).free(),
         nbytes / 2 * 2 * (P - 1) / P, nbytes / 2 * 2 * (P - 1) / P),
        ("gemm -> reduce_scatter bf16 8192x8192x2048 (TMA reduce-add into the owners' shards)",
         lambda: gemm_reduce_scatter(acc, x, w, shard), M * N * 2 * (P - 1) / P, M * N * 2 * (P - 1) / P),
    ]
    for name, fn, exp_tx, exp_rx in cases:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        c0 = counters(local) if rank == 0 else None
        dist.barrier()
        for _ in range(a.iters):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        c1 = counters(local) if rank == 0 else None
        dist.barrier()
        if rank == 0:
            row = dict(case=name, world=world, iters=a.iters, expected_tx_bytes_per_call=exp_tx, expected_rx_bytes_per_call=exp_rx)
            if c0 and c1:
                row.update(measured_tx_bytes_per_call=(c1[0] - c0[0]) / a.iters, measured_rx_bytes_per_call=(c1[1] - c0[1]) / a.iters,
                           tx_ratio=(c1[0] - c0[0]) / a.iters / exp_tx, rx_ratio=(c1[1] - c0[1]) / a.iters / exp_rx)
            else:
                row["note"] = "nvidia-smi nvlink -gt d reported no counters"
            print(json.dumps(row), flush=True)
            if a.out:
                os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
                with open(a.out, "a") as fh:
                    fh.write(json.dumps(row) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
