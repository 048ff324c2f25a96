"""GEMM -> reduce-scatter: fused tcgen05 plugin vs cuBLAS + NCCL (BASELINE config #5).

  python bench/gemm_rs.py                       # 1 GPU: fused kernel vs cuBLAS GEMM (tensor-core efficiency)
  python -m torch.distributed.run --nproc-per-node N ... bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192

Per rank: A[M, K] and W[N, K] bf16 (its K-slice); result: C = sum_r A_r W_r^T reduce-scattered along M.
Reports device-timed ms (CUDA events, max over ranks), TFLOP/s per GPU and the fraction of the
roofline max(FLOPs / measured cuBLAS peak, NVLink bytes / 770 GB/s).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402
from accl_b200.ops import gemm_reduce_scatter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gm", dest="m", type=int, default=8192)
    ap.add_argument("--gn", dest="n", type=int, default=8192)
    ap.add_argument("--gk", dest="k", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 single CTA per tile, 2 CTA pair (cta_group::2)")
    ap.add_argument("--f32", action="store_true", help="fp32 output shard (fp32 accumulation of the partial products)")
    ap.add_argument("--shapes", default="", help='several runs in one process: "MxNxK[:f32][:v1|:v2],..." (K per rank)')
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    runs = []
    if a.shapes:
        for spec in a.shapes.split(","):
            dims, *opts = spec.split(":")
            m, n, k = (int(v) for v in dims.lower().split("x"))
            runs.append((m, n, k, "f32" in opts, 1 if "v1" in opts else 2 if "v2" in opts else a.variant))
    else:
        runs.append((a.m, a.n, a.k, a.f32, a.variant))
    shard_mb = sum((m // world * n * (4 if f32 else 2) >> 20) + 1 for m, n, _, f32, _ in runs)
    acc = A.cuda_rank(rank, world, local, heap_mb=max(512, shard_mb + 384), max_ctas=32)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=16 << 10, max_egr_size=16 << 10, max_rndzv_size=1 << 30)
    for M, N, K, f32, variant in runs:
        a.f32, a.variant = f32, variant
        one(a, acc, rank, world, M, N, K)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def one(a, acc, rank, world, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.25).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.25).bfloat16()
    out = acc.create_buffer(M // world * N, torch.float32 if a.f32 else torch.bfloat16)
    ref_out = torch.empty(M // world, N, dtype=torch.bfloat16, device="cuda")

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

    def fused():
        gemm_reduce_scatter(acc, x, w, out, variant=a.variant)

    def baseline():
        c = x @ w.t()
        if world > 1:
            dist.reduce_scatter_tensor(ref_out, c)
        else:
            ref_out.copy_(c)

    def gemm_only():
        return x @ w.t()

    ms_f = timed(fused, a.iters)
    ms_b = timed(baseline, a.iters)
    ms_g = timed(gemm_only, a.iters)
    flops = 2.0 * M * N * K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    nvlink_bytes = M * N * 2 * (world - 1) / world  # partial tiles leaving this GPU
    if a.f32:
        nvlink_bytes *= 2
    link = 900e9  # nominal NVLink 5 per direction (BASELINE.json); the measured peer-copy rate is ~680-770 GB/s
    t_roof = max(flops / (peak_tf * 1e12), nvlink_bytes / link) * 1e3
    err = None
    err32 = None
    if a.check:
        fused()
        baseline()
        torch.cuda.synchronize()
        got = out.dev.view(M // world, N).float()
        err = float((got - ref_out.float()).abs().max())
        # against an fp32 reference of the same op: fp32 matmul of the bf16 operands, summed over ranks in fp32
        rows = slice(rank * (M // world), (rank + 1) * (M // world))
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        full = x.float() @ w.float().t()
        torch.backends.cuda.matmul.allow_tf32 = prev
        if world > 1:
            dist.all_reduce(full)
        err32 = float((got - full[rows]).abs().max())
        ref32 = float((ref_out.float() - full[rows]).abs().max())
    if rank == 0:
        row = dict(op="gemm_reduce_scatter", m=M, n=N, k_per_rank=K, world=world, fused_ms=ms_f, cublas_nccl_ms=ms_b,
                   cublas_gemm_only_ms=ms_g, fused_tflops=flops / ms_f * 1e-9, cublas_tflops=flops / ms_g * 1e-9,
                   speedup_vs_cublas_nccl=ms_b / ms_f, roofline_ms=t_roof, frac_of_roofline=t_roof / ms_f,
                   peak_source="MEASURED_PEAKS.json bf16_tflops" if peaks else "fallback 1590", link_GBps=900,
                   variant=a.variant, out_dtype="fp32" if a.f32 else "bf16", max_abs_err_vs_cublas_nccl=err,
                   max_abs_err_vs_fp32_ref=err32, cublas_nccl_err_vs_fp32_ref=ref32 if a.check else None)
        print(json.dumps(row), flush=True)
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            with open(a.out, "a") as fh:
                fh.write(json.dumps(row) + "\n")
    del x, w, ref_out
    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
