"""GPU tests: persistent engine mode, device-issued calls and the fused plugins."""
import pytest
import torch

import accl_b200 as A
from accl_b200 import SUM
from accl_b200.ops import gemm_reduce_scatter, vadd_allreduce

pytestmark = pytest.mark.gpu
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
CFG = dict(n_egr_rx_bufs=4, egr_rx_buf_size=16 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)


def devices(world):
    return [r % max(NGPU, 1) for r in range(world)]


def close(a, b, rtol, atol):
    return torch.allclose(a.cpu().double(), b.cpu().double(), rtol=rtol, atol=atol)


@pytest.mark.parametrize("variant,out_dtype", [(0, torch.bfloat16), (1, torch.bfloat16), (2, torch.bfloat16), (2, torch.float32),
                                               (1, torch.float32)])
@pytest.mark.parametrize("world", [1, 2])
def test_gemm_reduce_scatter_matches_fp32_reference(world, variant, out_dtype):
    """Both tcgen05 kernels (one CTA per 128 x 256 tile; CTA pair, cta_group::2, per 256 x 256 tile), bf16 and fp32 shards,
    against an fp32 matmul of the same bf16 operands summed over ranks in fp32."""
    M, N, K = 512 * world, 512, 320   # several tiles per rank, a K that is not a multiple of the stage depth

    def fn(a, r, w):
        g = torch.Generator().manual_seed(100 + r)
        x = (torch.randn(M, K, generator=g) * 0.5).bfloat16().cuda(a.cuda_device)
        wt = (torch.randn(N, K, generator=g) * 0.5).bfloat16().cuda(a.cuda_device)
        out = gemm_reduce_scatter(a, x, wt, variant=variant, out_dtype=out_dtype)
        torch.cuda.current_stream().synchronize()
        assert out.dev.dtype == out_dtype
        return out.dev.view(M // w, N).float().cpu(), x.float().cpu(), wt.float().cpu()

    res = A.run_cuda_ranks(devices(world), fn, CFG, heap_mb=64, max_ctas=4)
    full = sum(x @ wt.t() for _, x, wt in res)  # fp32 reference of the same op
    for r, (shard, _, _) in enumerate(res):
        ref = full[r * (M // world):(r + 1) * (M // world)]
        if out_dtype == torch.float32:
            assert close(shard, ref, 1e-4, 1e-3), (r, (shard - ref).abs().max())    # fp32 accumulation end to end
        else:
            assert close(shard, ref, 2e-2, 2e-1 * world), (r, (shard - ref).abs().max())  # bf16 output, bf16 adds of `world` partials


def test_engine_mode_collectives():
    n = 5000

    def fn(a, r, w):
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.host[:] = torch.arange(n, dtype=torch.float32) + r
        for _ in range(3):
            a.allreduce(s, d, n, SUM)
        ref = sum(torch.arange(n, dtype=torch.float32) + q for q in range(w))
        assert torch.equal(d.host, ref)
        big = 1 << 18
        bs, bd = a.create_buffer(big), a.create_buffer(big)
        bs.host[:] = float(r + 1)
        a.allreduce(bs, bd, big, SUM)
        assert float(bd.host[-1]) == sum(range(1, w + 1))
        req = a.nop()
        assert req.retcode() == 0
        a.barrier()
        return a.describe()

    out = A.run_cuda_ranks(devices(2), fn, CFG, heap_mb=64, max_ctas=4, engine=True)
    assert "mode=engine" in out[0]


def test_vadd_plugin_issues_allreduce_from_the_device():
    n = 1 << 16

    def fn(a, r, w):
        x, y, out = a.create_buffer(n), a.create_buffer(n), a.create_buffer(n)
        x.dev.fill_(float(r + 1))
        y.dev.copy_(torch.arange(n, dtype=torch.float32, device=x.dev.device))
        status = vadd_allreduce(a, x, y, out)
        torch.cuda.current_stream().synchronize()
        assert int(status.item()) == 0
        ref = torch.arange(n, dtype=torch.float32) * w + sum(range(1, w + 1))
        assert torch.equal(out.dev.cpu(), ref)

    A.run_cuda_ranks(devices(2), fn, CFG, heap_mb=64, max_ctas=4, engine=True)


def test_tensor_group_and_row_parallel_linear():
    from accl_b200.parallel import TensorGroup, RowParallelLinear, ring_exchange
    M, K, N = 512, 128, 256

    def fn(a, r, w):
        g = TensorGroup(a)
        dev = torch.device("cuda", a.cuda_device)
        t = torch.full((1000,), float(r + 1), device=dev)          # ordinary torch tensor: staged through the heap
        g.all_reduce(t)
        assert torch.all(t == sum(range(1, w + 1)))
        h = g.empty(64, dtype=torch.float32)                        # heap tensor: zero-copy
        h.fill_(float(r))
        g.all_reduce(h)
        assert torch.all(h == sum(range(w)))
        out = torch.empty(64 * w, device=dev)
        g.all_gather_into_tensor(out, torch.full((64,), float(r), device=dev))
        assert torch.equal(out.view(w, 64)[:, 0].cpu(), torch.arange(w, dtype=torch.float32))
        a_blk, b_blk = torch.full((32,), float(r), device=dev), torch.empty(32, device=dev)
        ring_exchange(g, a_blk, b_blk, tag=4)
        assert torch.all(b_blk == float((r - 1) % w))
        lin = RowParallelLinear(g, K, N)
        gen = torch.Generator().manual_seed(5 + r)
        x = (torch.randn(M * w, K, generator=gen) * 0.5).bfloat16().to(dev)
        y = lin(x)
        torch.cuda.current_stream().synchronize()
        return y.float().cpu(), x.float().cpu(), lin.weight.data.float().cpu()

    res = A.run_cuda_ranks(devices(2), fn, CFG, heap_mb=64, max_ctas=4)
    full = sum(x @ wt.t() for _, x, wt in res)
    rows = full.shape[0] // 2
    for r, (y, _, _) in enumerate(res):
        assert close(y, full[r * rows:(r + 1) * rows], 2e-2, 3e-1)


def test_user_kernel_in_the_stream_path():
    """send -> recv_to_stream -> user kernel (device::Data pull / +1 / push) -> send_from_stream -> recv
    (reference test/host/hls_simulator/test.cpp:153 `test_loopback`)."""
    from accl_b200 import DataType
    from accl_b200.ops import stream_loopback
    n = 2048

    def fn(a, r, w):
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.host[:] = torch.arange(n, dtype=torch.float32) + 100 * r
        nxt, prv = (r + 1) % w, (r - 1) % w
        req = a.send(s, n, nxt, tag=3, run_async=True)
        a.recv_to_stream(DataType.float32, n, prv, tag=3)
        req.wait()
        st = stream_loopback(a, n, add_one=True)
        req = a.send_from_stream(DataType.float32, n, nxt, tag=4, run_async=True)
        a.recv(d, n, prv, tag=4)
        req.wait()
        torch.cuda.synchronize()
        assert int(st.item()) == 0
        # data made two hops: it comes from the rank two places behind me, plus one
        src = (r - 2) % w
        assert torch.equal(d.host, torch.arange(n, dtype=torch.float32) + 100 * src + 1)
    A.run_cuda_ranks([0, 0], fn, dict(n_egr_rx_bufs=4, egr_rx_buf_size=16 << 10, max_egr_size=16 << 10,
                                          max_rndzv_size=1 << 26), heap_mb=64, max_ctas=4)


def test_heap_mem_pool_tensors_are_zero_copy_operands():
    """torch tensors allocated under `use_mem_pool(accl.heap_mem_pool())` live in the symmetric heap: TensorGroup takes
    them without staging (this is how DDP's gradient buckets become zero-copy on the "accl" backend)."""
    from accl_b200.parallel import TensorGroup
    n = 1 << 18

    def fn(a, r, w):
        dev = torch.device("cuda", a.cuda_device)
        g = TensorGroup(a)
        pool = a.heap_mem_pool()
        with torch.cuda.use_mem_pool(pool):
            t = torch.full((n,), float(r + 1), device=dev)
            u = torch.arange(n, device=dev, dtype=torch.float32) + r
        plain = torch.full((n,), float(r + 1), device=dev)
        assert a.heap_contains(t) and a.heap_contains(u) and not a.heap_contains(plain)
        buf, staged = g._buffer_of(t, "s")
        assert not staged and buf.dev.data_ptr() == t.data_ptr()
        g.all_reduce(t)
        g.all_reduce(u)
        g.all_reduce(plain)
        torch.cuda.current_stream().synchronize()
        g.check(block=True)
        s = float(sum(range(1, w + 1)))
        assert torch.all(t == s) and torch.all(plain == s)
        assert torch.equal(u.cpu(), torch.arange(n, dtype=torch.float32) * w + sum(range(w)))
        assert not g._scratch or all(k[0] != "s" or v.length == n for k, v in g._scratch.items())  # only `plain` was staged
        del t, u
        return True

    assert all(A.run_cuda_ranks(devices(2), fn, CFG, heap_mb=64, max_ctas=4))
