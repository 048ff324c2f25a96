"""DistributedDataParallel step time on the `accl` torch.distributed backend vs NCCL (same model, same data, same box).

  torchrun --nproc-per-node N bench/ddp.py --backend accl [--engine]
  torchrun --nproc-per-node N bench/ddp.py --backend nccl

Model: a stack of `--layers` Linear(`--width`, `--width`) + GELU, fp32 parameters (the gradient all-reduce buckets are
what is being compared), AdamW step included.  Device-timed with CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="accl", choices=["accl", "nccl"])
    ap.add_argument("--engine", action="store_true")
    ap.add_argument("--heap-buckets", action="store_true", help="accl: allocate DDP's gradient buckets in the symmetric heap (zero-copy all-reduce)")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--width", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if a.backend == "accl":
        os.environ["ACCL_PG_ENGINE"] = "1" if a.engine else "0"
        os.environ.setdefault("ACCL_HEAP_MB", "4096")
        import accl_b200.parallel.process_group  # noqa: F401  (registers the backend)
    dist.init_process_group(a.backend, init_method="env://", rank=rank, world_size=world, **({"device_id": dev} if a.backend == "nccl" else {}))
    torch.manual_seed(0)
    layers = []
    for _ in range(a.layers):
        layers += [torch.nn.Linear(a.width, a.width), torch.nn.GELU()]
    model = torch.nn.Sequential(*layers).to(dev)
    import contextlib
    def ctx():
        if a.backend == "accl" and a.heap_buckets:
            return torch.cuda.use_mem_pool(accl_b200.parallel.process_group.heap_mem_pool())
        return contextlib.nullcontext()

    with ctx():
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local] if a.backend == "nccl" else None, bucket_cap_mb=64,
                                                        gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4)
    x = torch.randn(a.batch, a.width, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))

    def step():
        opt.zero_grad(set_to_none=True)
        ddp(x).square().mean().backward()
        opt.step()

    # DDP rebuilds its buckets (in gradient-ready order) at the start of the second forward: keep the warm-up forwards
    # inside the pool so that the rebuilt buckets are heap-resident too
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        with ctx():
            out = ddp(x)
        out.square().mean().backward()
        opt.step()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        step()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / a.steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # gradients must agree across ranks after the step
    g = next(model.parameters()).grad.clone()
    chk = g.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX)
    ok = bool(torch.allclose(chk, g))
    if rank == 0:
        nparam = sum(p.numel() for p in model.parameters())
        row = dict(bench="ddp_step", backend=a.backend + ("+engine" if a.engine and a.backend == "accl" else "") + ("+heap_buckets" if a.heap_buckets else ""), world=world, params=nparam,
                   grad_mb=nparam * 4 / 2 ** 20, ms_per_step=float(t.item()), grads_agree=ok)
        if a.backend == "accl":
            g = accl_b200.parallel.process_group._primary["pg"].group
            row["staged_mb"] = sum(b.nbytes for b in g._scratch.values()) / 2 ** 20   # 0 when every bucket was a zero-copy operand
            row["zero_copy_tensors"] = len(g._wrapped)
        print(json.dumps(row), flush=True)
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            with open(a.out, "a") as fh:
                fh.write(json.dumps(row) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
