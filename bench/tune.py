"""Knob sweep for the large-message algorithms: every combination is one `set_tuning` away (no rebuild), so a
whole grid costs one job.  Device-timed (CUDA events, median of batches, max over ranks), result checked once per
combination against a float64 reference on a slice.

  torchrun --nproc-per-node 8 bench/tune.py --what allreduce --mb 256 --out gpurun_out/tune_ar.jsonl

what = allreduce : nvls_ctas x nvls_unroll x hybrid_16ths   (NVLS two-shot; hybrid adds the peer two-shot body)
       reduce    : reduce_push on / off, channels
       bcast     : bcast_flags on / off, channels
"""
import argparse
import itertools
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="allreduce")
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="a coarser grid")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = getattr(torch, args.dtype)
    esz = torch.empty((), dtype=dt).element_size()
    nbytes = args.mb << 20
    n = nbytes // esz
    acc = A.cuda_rank(rank, world, local, heap_mb=(3 * nbytes >> 20) + 768, max_ctas=128)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=64 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)
    if rank == 0:
        print("#", acc.describe(), flush=True)
    s, d = acc.create_buffer(n, dt), acc.create_buffer(n, dt)
    g = torch.Generator(device="cuda").manual_seed(11 + rank)
    s.dev.copy_((torch.rand(n, device="cuda", generator=g) * 2 - 1).to(dt))
    probe = slice(n // 3, n // 3 + (1 << 20))
    ref_sum = s.dev[probe].double()
    dist.all_reduce(ref_sum)
    ref_b = s.dev[probe].double()
    dist.broadcast(ref_b, 0)
    kw = dict(from_fpga=True, to_fpga=True, run_async=True)
    tol = {torch.float32: 1e-4, torch.bfloat16: 0.15, torch.float16: 0.03}[dt]

    def timed(fn):
        for _ in range(2):
            fn()
        out = []
        for _ in range(args.batches):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / args.iters], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out.append(float(t.item()))
        return sorted(out)[len(out) // 2]

    fh = open(args.out, "a") if rank == 0 and args.out else None

    def emit(row):
        if rank == 0:
            print(json.dumps(row), flush=True)
            if fh:
                fh.write(json.dumps(row) + "\n")
                fh.flush()

    whats = args.what.split(",")
    if "allreduce" in whats:
        ref = ref_sum
        ctas = [16, 24, 32, 48, 64] if args.quick else [16, 24, 32, 48, 64, 80, 96, 128]
        unroll = [2, 4, 8] if args.quick else [2, 4, 8, 16]
        hyb = [0] if args.quick else [0, 2, 4]
        f = 2.0 * (world - 1) / world
        for c, u, h in itertools.product(ctas, unroll, hyb):
            for k, v in (("nvls_ctas", c), ("nvls_unroll", u), ("hybrid_16ths", h)):
                acc.set_tuning(k, v)
            d.dev.zero_()
            ms = timed(lambda: acc.allreduce(s, d, n, A.SUM, **kw).free())
            err = float((d.dev[probe].double() - ref).abs().max())
            emit(dict(what="allreduce", mb=args.mb, dtype=args.dtype, world=world, nvls_ctas=c, nvls_unroll=u, hybrid_16ths=h,
                      us=ms * 1e3, busbw=nbytes / ms * 1e-6 * f, ok=err <= tol * world, err=err))
        x = s.dev.clone()
        ms = timed(lambda: dist.all_reduce(x))
        del x
        emit(dict(what="allreduce", mb=args.mb, dtype=args.dtype, world=world, impl="nccl", us=ms * 1e3, busbw=nbytes / ms * 1e-6 * f))
        for k, v in (("nvls_ctas", 32), ("nvls_unroll", 4), ("hybrid_16ths", 0)):
            acc.set_tuning(k, v)
    if "reduce" in whats:
        ref = ref_sum
        for push, c in itertools.product([0, 1, 2], [64, 128]):
            acc.set_tuning("reduce_push", push)
            acc.set_tuning("max_ctas", c)
            d.dev.zero_()
            ms = timed(lambda: acc.reduce(s, d, n, 0, A.SUM, **kw).free())
            err = float((d.dev[probe].double() - ref).abs().max()) if rank == 0 else 0.0
            emit(dict(what="reduce", mb=args.mb, dtype=args.dtype, world=world, reduce_push=push, max_ctas=c, us=ms * 1e3,
                      busbw=nbytes / ms * 1e-6, ok=err <= tol * world, err=err))
        x = s.dev.clone()
        ms = timed(lambda: dist.reduce(x, 0))
        del x
        emit(dict(what="reduce", mb=args.mb, dtype=args.dtype, world=world, impl="nccl", us=ms * 1e3, busbw=nbytes / ms * 1e-6))
        acc.set_tuning("reduce_push", 0)
        acc.set_tuning("max_ctas", 128)
    if "bcast" in whats:
        ref = ref_b
        for fl, c in itertools.product([0, 1, 2], [64, 128]):
            acc.set_tuning("bcast_flags", fl)
            acc.set_tuning("max_ctas", c)
            if rank != 0:
                s.dev[probe].zero_()
            torch.cuda.synchronize()
            dist.barrier()
            ms = timed(lambda: acc.bcast(s, n, 0, **kw).free())
            err = float((s.dev[probe].double() - ref).abs().max())
            emit(dict(what="bcast", mb=args.mb, dtype=args.dtype, world=world, bcast_flags=fl, max_ctas=c, us=ms * 1e3,
                      busbw=nbytes / ms * 1e-6, ok=err <= tol, err=err))
        ms = timed(lambda: dist.broadcast(s.dev, 0))
        emit(dict(what="bcast", mb=args.mb, dtype=args.dtype, world=world, impl="nccl", us=ms * 1e3, busbw=nbytes / ms * 1e-6))
    if fh:
        fh.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
