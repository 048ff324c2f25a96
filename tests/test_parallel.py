"""Torch-facing layer on the CPU emulator: TensorGroup, GradBucket, ring_exchange and the registered
`torch.ops.accl_b200.*` custom ops (same code paths the CUDA backend uses, minus stream ordering)."""
import pytest
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


def test_staging_is_bounded_large_tensors_go_in_chunks():
    """Tensors outside the heap are staged through one buffer per (role, dtype); beyond `scratch_bytes`
    all_reduce / broadcast run chunk by chunk (DDP's 250 MB parameter broadcast must not exhaust the heap)."""
    def fn(a, r, w):
        g = TensorGroup(a, scratch_bytes=4096)                       # 1024 floats per chunk
        t = torch.arange(5000, dtype=torch.float32) * (r + 1)
        g.all_reduce(t)
        assert torch.equal(t, torch.arange(5000, dtype=torch.float32) * sum(range(1, w + 1)))
        b = torch.full((3000,), float(r))
        g.broadcast(b, root=2)
        assert torch.all(b == 2.0)
        for n in (10, 700, 20, 1024):                                # different sizes share (and grow) one staging buffer
            u = torch.full((n,), 1.0)
            g.all_reduce(u)
            assert torch.all(u == w)
        assert len(g._scratch) == 1 and max(len(v) for v in g._scratch.values()) == 1024
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


@pytest.mark.parametrize("one_hop", [False, True], ids=["rings", "one_hop"])
def test_zero_optimizer_matches_plain_sgd(one_hop):
    from accl_b200.parallel.strategies import ZeroOptimizer
    torch.manual_seed(0)
    w0, b0 = torch.randn(4, 5), torch.randn(7)
    grads = [[torch.randn(4, 5, generator=torch.Generator().manual_seed(10 * s + r)) for r in range(W)] for s in range(3)]

    def fn(a, r, w):
        a.set_one_hop_schedules(one_hop)   # reference-style rings / trees, or the B200 backend's schedules
        g = TensorGroup(a)
        pw, pb = w0.clone(), b0.clone()
        opt = ZeroOptimizer(g, [pw, pb], lr=0.5, momentum=0.9)
        for s in range(3):
            opt.grad_views[0].copy_(grads[s][r])
            opt.grad_views[1].fill_(float(r))
            opt.step()
        return pw, pb

    res = A.run_ranks(W, fn, CFG)
    # reference: plain momentum SGD on the averaged gradients
    pw, pb, vw, vb = w0.clone(), b0.clone(), torch.zeros(4, 5), torch.zeros(7)
    for s in range(3):
        gw, gb = sum(grads[s]) / W, torch.full((7,), sum(range(W)) / W)
        vw, vb = 0.9 * vw + gw, 0.9 * vb + gb
        pw, pb = pw - 0.5 * vw, pb - 0.5 * vb
    for w_, b_ in res:
        assert torch.allclose(w_, pw, atol=1e-5) and torch.allclose(b_, pb, atol=1e-5)


@pytest.mark.parametrize("one_hop", [False, True], ids=["rings", "one_hop"])
def test_column_parallel_pipeline_moe_ulysses(one_hop):
    from accl_b200.parallel.strategies import (ColumnParallelLinear, moe_combine, moe_dispatch, pipeline_recv,
                                               pipeline_send, ulysses_head_to_seq, ulysses_seq_to_head)
    S, H, D = 6 * W, 2 * W, 4
    full = torch.arange(S * H * D, dtype=torch.float32).view(S, H, D)

    def fn(a, r, w):
        a.set_one_hop_schedules(one_hop)
        g = TensorGroup(a)
        # column-parallel linear with gathered output == the unsharded GEMM
        lin = ColumnParallelLinear(g, 8, 3, gather_output=True)
        with torch.no_grad():
            lin.weight.copy_(torch.arange(24, dtype=torch.float32).view(3, 8) + 100 * r)
        x = torch.ones(5, 8)
        y = lin(x)
        ref = torch.cat([x @ (torch.arange(24, dtype=torch.float32).view(3, 8) + 100 * q).t() for q in range(w)], dim=-1)
        assert torch.equal(y, ref)
        # pipeline hand-off r -> r + 1 with micro-batch tags
        act = torch.full((40,), float(r))
        got = torch.empty(40)
        reqs = []
        if r + 1 < w:
            reqs.append(pipeline_send(g, act, r + 1, microbatch=3))
        if r > 0:
            pipeline_recv(g, got, r - 1, microbatch=3)
            assert torch.all(got == float(r - 1))
        for q in reqs:
            q.wait()
        # MoE dispatch / combine round trip
        cap, hid = 3, 4
        routed = torch.stack([torch.full((cap, hid), float(10 * r + q)) for q in range(w)])
        recv = moe_dispatch(g, routed)
        for q in range(w):
            assert torch.all(recv[q] == float(10 * q + r))
        back = moe_combine(g, recv * 2)
        assert torch.equal(back, routed * 2)
        # Ulysses: sequence shard -> head shard -> back
        mine = full[r * (S // w):(r + 1) * (S // w)].contiguous()
        heads = ulysses_seq_to_head(g, mine)
        assert torch.equal(heads, full[:, r * (H // w):(r + 1) * (H // w)])
        assert torch.equal(ulysses_head_to_seq(g, heads.contiguous()), mine)
    A.run_ranks(W, fn, CFG)
