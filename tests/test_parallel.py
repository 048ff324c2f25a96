"""Torch-facing layer on the CPU emulator: TensorGroup, GradBucket, ring_exchange and the registered
`torch.ops.accl_b200.*` custom ops (same code paths the CUDA backend uses, minus stream ordering)."""
import torch

import accl_b200 as A
from accl_b200.ops import torch_ops
from accl_b200.parallel import GradBucket, TensorGroup, ring_exchange

CFG = dict(n_egr_rx_bufs=16, egr_rx_buf_size=1024, max_egr_size=1024, max_rndzv_size=1 << 20)
W = 3


def test_tensor_group_collectives():
    def fn(a, r, w):
        g = TensorGroup(a)
        t = torch.full((1000,), float(r + 1))                       # ordinary tensor: staged
        g.all_reduce(t)
        assert torch.all(t == sum(range(1, w + 1)))
        h = g.empty(64)                                             # engine-backed tensor: no staging
        h.fill_(float(r))
        g.all_reduce(h, A.MAX)
        assert torch.all(h == w - 1)
        out = torch.empty(50 * w)
        g.all_gather_into_tensor(out, torch.full((50,), float(r)))
        assert torch.equal(out.view(w, 50)[:, 0], torch.arange(w, dtype=torch.float32))
        rs = torch.empty(40)
        g.reduce_scatter_tensor(rs, torch.arange(40 * w, dtype=torch.float32))
        assert torch.equal(rs, w * torch.arange(40 * r, 40 * (r + 1), dtype=torch.float32))
        b = torch.full((33,), float(r))
        g.broadcast(b, root=1)
        assert torch.all(b == 1.0)
        x = torch.arange(w * 7, dtype=torch.float32) + 100 * r
        y = torch.empty(w * 7)
        g.all_to_all_single(y, x)
        for q in range(w):
            assert torch.equal(y[q * 7:(q + 1) * 7], torch.arange(r * 7, (r + 1) * 7, dtype=torch.float32) + 100 * q)
        g.barrier()
    A.run_ranks(W, fn, CFG)


def test_ring_exchange_and_grad_bucket():
    def fn(a, r, w):
        g = TensorGroup(a)
        send, recv = torch.full((500,), float(r)), torch.empty(500)   # 2000 B > eager threshold: rendezvous
        s, _ = ring_exchange(g, send, recv, tag=4)
        s.wait()
        assert torch.all(recv == float((r - 1) % w))
        bucket = GradBucket(g, 10 + 6, dtype=torch.float32)
        gw, gb = bucket.views([(2, 5), (6,)])
        gw.fill_(float(r))
        gb.fill_(1.0)
        bucket.all_reduce(average=True)
        assert torch.allclose(gw, torch.full((2, 5), sum(range(w)) / w)) and torch.allclose(gb, torch.ones(6))
        gw.fill_(float(r + 1))
        shard = bucket.reduce_scatter()
        assert shard.numel() == bucket.numel // w
    A.run_ranks(W, fn, CFG)


def test_registered_torch_ops():
    def fn(a, r, w):
        torch_ops.set_default_group(TensorGroup(a))
        t = torch.full((10,), float(r))
        torch.ops.accl_b200.all_reduce(t, "sum")
        assert torch.all(t == sum(range(w)))
        torch.ops.accl_b200.all_reduce(t, "max")
        out = torch.empty(4 * w)
        torch.ops.accl_b200.all_gather(out, torch.full((4,), float(r)))
        assert torch.equal(out[::4], torch.arange(w, dtype=torch.float32))
        rs = torch.empty(4)
        torch.ops.accl_b200.reduce_scatter(rs, torch.ones(4 * w), "sum")
        assert torch.all(rs == w)
        b = torch.full((5,), float(r))
        torch.ops.accl_b200.broadcast(b, 2)
        assert torch.all(b == 2.0)
        torch.ops.accl_b200.barrier()
    A.run_ranks(W, fn, CFG)
