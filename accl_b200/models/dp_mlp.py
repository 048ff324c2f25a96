"""Data-parallel training of a small MLP on top of the collective API — the end-to-end consumer the
library is built for, runnable on either backend:

  python -m accl_b200.models.emulator -n 4 -- python -m accl_b200.models.dp_mlp            # CPU emulator, 4 processes
  torchrun --nproc-per-node 8 -m accl_b200.models.dp_mlp --backend cuda                    # 8 x B200

Each rank draws its own shard of a synthetic regression problem; gradients are written into one flat bucket
in engine memory and either all-reduced (`--zero 0`, DDP) or reduce-scattered / all-gathered around a sharded
optimizer step (`--zero 1`, ZeRO-1).  The script checks that every rank ends with bit-identical parameters and
that the loss went down, and prints one JSON line on rank 0.
"""
import argparse
import json

import torch

from ..parallel import GradBucket, TensorGroup, init_from_env
from ..parallel.strategies import ZeroOptimizer


def make_model(d_in, d_hidden, device, seed=0):
    g = torch.Generator().manual_seed(seed)  # identical initial weights on every rank
    w1 = (torch.randn(d_hidden, d_in, generator=g) / d_in ** 0.5).to(device).requires_grad_()
    b1 = torch.zeros(d_hidden, device=device, requires_grad=True)
    w2 = (torch.randn(1, d_hidden, generator=g) / d_hidden ** 0.5).to(device).requires_grad_()
    return [w1, b1, w2]


def forward(params, x):
    w1, b1, w2 = params
    return torch.tanh(x @ w1.t() + b1) @ w2.t()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default=None, choices=[None, "cuda", "emulator"])
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--batch", type=int, default=64, help="per-rank batch")
    ap.add_argument("--zero", type=int, default=0, choices=[0, 1])
    ap.add_argument("--lr", type=float, default=0.2)
    a = ap.parse_args(argv)
    acc = init_from_env(a.backend, **({} if a.backend == "cuda" else
                                      dict(max_egr_size=1024, egr_rx_buf_size=1024, n_egr_rx_bufs=16, max_rndzv_size=1 << 20)))
    group = TensorGroup(acc)
    rank, world = group.rank, group.world
    device = torch.device("cuda", acc.cuda_device) if acc.is_cuda else torch.device("cpu")
    d_in, d_hidden = 16, 32
    params = make_model(d_in, d_hidden, device)
    true_w = torch.linspace(-1, 1, d_in, device=device)
    gen = torch.Generator().manual_seed(1000 + rank)          # a different data shard per rank
    shapes = [tuple(p.shape) for p in params]
    if a.zero:
        opt = ZeroOptimizer(group, [p.data for p in params], lr=a.lr)
        grad_views = opt.grad_views
    else:
        bucket = GradBucket(group, sum(p.numel() for p in params), dtype=torch.float32)
        grad_views = bucket.views(shapes)
    losses = []
    for _ in range(a.steps):
        x = torch.randn(a.batch, d_in, generator=gen).to(device)
        y = torch.sin(x @ true_w).unsqueeze(1)
        loss = torch.nn.functional.mse_loss(forward(params, x), y)
        grads = torch.autograd.grad(loss, params)
        for v, g in zip(grad_views, grads):
            v.copy_(g)
        if a.zero:
            opt.step()                                         # reduce_scatter -> shard update -> all_gather
        else:
            bucket.all_reduce(average=True)                    # one all-reduce for the whole model
            with torch.no_grad():
                for p, v in zip(params, grad_views):
                    p -= a.lr * v
        lt = loss.detach().reshape(1).clone()
        group.all_reduce(lt)
        losses.append(float(lt.item()) / world)
    # every rank must hold the same model: compare a checksum
    flat = torch.cat([p.detach().reshape(-1) for p in params]).double()
    chk = torch.stack([flat.sum(), (flat * flat).sum()]).float().cpu()
    mx, mn = chk.clone(), chk.clone()
    group.all_reduce(mx, op=__import__("accl_b200").MAX)
    neg = -mn
    group.all_reduce(neg, op=__import__("accl_b200").MAX)
    same = bool(torch.equal(mx, -neg))
    ok = same and losses[-1] < 0.7 * losses[0]
    group.barrier()
    if rank == 0:
        print(json.dumps({"model": "dp_mlp", "world": world, "zero": a.zero, "backend": "cuda" if acc.is_cuda else "emulator",
                          "loss_first": losses[0], "loss_last": losses[-1], "replicas_identical": same, "ok": ok}), flush=True)
    acc.deinit()
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
