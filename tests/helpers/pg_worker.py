"""Worker of tests/test_emulator_multiprocess.py::test_torch_distributed_backend (one process per rank)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import accl_b200.parallel.process_group  # noqa
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("accl", init_method=f"tcp://127.0.0.1:{os.environ['PG_PORT']}", rank=rank, world_size=world)
t = torch.full((1000,), float(rank + 1))
dist.all_reduce(t)
assert torch.all(t == world * (world + 1) / 2), t[:4]
dist.all_reduce(t, op=dist.ReduceOp.MAX)
b = torch.full((33,), float(rank)); dist.broadcast(b, src=1); assert torch.all(b == 1)
out = torch.empty(4 * world); dist.all_gather_into_tensor(out, torch.full((4,), float(rank)))
assert torch.equal(out[::4], torch.arange(world, dtype=torch.float32))
lst = [torch.empty(3) for _ in range(world)]; dist.all_gather(lst, torch.full((3,), float(rank)))
assert all(torch.all(lst[q] == q) for q in range(world))
rs = torch.empty(5); dist.reduce_scatter_tensor(rs, torch.ones(5 * world)); assert torch.all(rs == world)
a2a = torch.empty(2 * world); dist.all_to_all_single(a2a, torch.arange(2 * world, dtype=torch.float32) + 100 * rank)
assert torch.equal(a2a[::2], torch.tensor([100.0 * q + 2 * rank for q in range(world)]))
r = torch.full((7,), float(rank)); dist.reduce(r, dst=0, op=dist.ReduceOp.SUM)
if rank == 0: assert torch.all(r == sum(range(world)))
if rank == 0: dist.send(torch.arange(10, dtype=torch.float32), dst=1)
if rank == 1:
    x = torch.empty(10); dist.recv(x, src=0); assert torch.equal(x, torch.arange(10, dtype=torch.float32))
dist.barrier()
# DistributedDataParallel on top
torch.manual_seed(0)
model = torch.nn.Linear(8, 4)
ddp = torch.nn.parallel.DistributedDataParallel(model)
xs = torch.randn(16, 8, generator=torch.Generator().manual_seed(rank))
ddp(xs).sum().backward()
g = model.weight.grad.clone()
chk = g.clone(); dist.all_reduce(chk, op=dist.ReduceOp.MAX)
assert torch.allclose(chk, g), "DDP gradients differ across ranks"
dist.barrier()
print(f"pg rank {rank}: ok", flush=True)
dist.destroy_process_group()
