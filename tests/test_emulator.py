"""Host-API test matrix on the CPU emulator backend (no GPU needed).

Mirrors the reference's functional suite (test/host/xrt/src/test.cpp:30-1159):
one test per primitive/collective, parametrised over roots, reduce functions,
segmentation edge cases and compression, run in an eager configuration and in
a rendezvous configuration (the reference only exercises rendezvous on RDMA
hardware; here both protocols run everywhere).  Ranks are threads of one
process sharing an in-process fabric.
"""
import numpy as np
import pytest
import torch

import accl_b200 as A
from accl_b200 import DataType, MAX, SUM

EAGER = dict(n_egr_rx_bufs=16, egr_rx_buf_size=1024, max_egr_size=1 << 20, max_rndzv_size=1 << 24)
RNDZV = dict(n_egr_rx_bufs=16, egr_rx_buf_size=64, max_egr_size=64, max_rndzv_size=32 * 1024)
PROTOCOLS = [pytest.param(EAGER, id="eager"), pytest.param(RNDZV, id="rndzv")]
WORLDS = [2, 3, 4]
COUNT = 300  # > 64 B so the rendezvous configuration really uses rendezvous; not a multiple of any world size


def data(count, rank, dtype=torch.float32, salt=0):
    g = torch.Generator().manual_seed(1234 + 17 * rank + salt)
    if dtype in (torch.int32, torch.int64):
        return torch.randint(-1000, 1000, (count,), generator=g, dtype=dtype)
    return (torch.rand(count, generator=g, dtype=torch.float32) * 8 - 4).to(dtype)


def reduce_ref(world, count, func, dtype=torch.float32, salt=0):
    xs = [data(count, r, dtype, salt).to(torch.float64 if dtype.is_floating_point else torch.int64) for r in range(world)]
    out = xs[0].clone()
    for x in xs[1:]:
        out = out + x if func == SUM else torch.maximum(out, x)
    return out


def close(a, b, rtol=1e-5, atol=1e-5):
    return torch.allclose(a.to(torch.float64), b.to(torch.float64), rtol=rtol, atol=atol)


# ------------------------------------------------------------ single rank
def test_copy_and_combine_single_rank():
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, 0)
        a.copy(s, d, COUNT)
        assert torch.equal(s.host, d.host)
        x, y, z = a.create_buffer(COUNT), a.create_buffer(COUNT), a.create_buffer(COUNT)
        x.host[:] = data(COUNT, 1)
        y.host[:] = data(COUNT, 2)
        a.combine(COUNT, SUM, x, y, z)
        assert close(z.host, x.host + y.host)
        a.combine(COUNT, MAX, x, y, z)
        assert torch.equal(z.host, torch.maximum(x.host, y.host))
    A.run_ranks(1, fn)


def test_copy_stream_roundtrip():
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, 0)
        a.copy_to_stream(s, COUNT)      # engine -> user stream (looped back)
        a.copy_from_stream(d, COUNT)    # user stream -> memory
        assert torch.equal(s.host, d.host)
        a.copy_to_stream(s, COUNT)
        a.copy_from_to_stream(DataType.float32, COUNT)
        a.copy_from_stream(d, COUNT)
        assert torch.equal(s.host, d.host)
    A.run_ranks(1, fn)


@pytest.mark.parametrize("kinds", [("p2p", "device"), ("host_only", "device"), ("device", "host_only"),
                                   ("host_only", "host_only"), ("p2p", "p2p")])
def test_copy_between_buffer_kinds(kinds):
    def fn(a, r, w):
        ks = [getattr(A.BufferKind, k) for k in kinds]
        s, d = a.create_buffer(COUNT, kind=ks[0]), a.create_buffer(COUNT, kind=ks[1])
        s.host[:] = data(COUNT, 3)
        a.copy(s, d, COUNT)
        assert torch.equal(s.host, d.host)
    A.run_ranks(1, fn)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float64, torch.int32, torch.int64])
@pytest.mark.parametrize("func", [SUM, MAX])
def test_combine_dtypes(dtype, func):
    def fn(a, r, w):
        x, y, z = (a.create_buffer(COUNT, dtype) for _ in range(3))
        x.host[:] = data(COUNT, 1, dtype)
        y.host[:] = data(COUNT, 2, dtype)
        a.combine(COUNT, func, x, y, z)
        ref = (x.host.float() + y.host.float()).to(dtype) if (func == SUM and dtype.is_floating_point and dtype != torch.float64) \
            else (x.host + y.host if func == SUM else torch.maximum(x.host, y.host))
        assert torch.equal(z.host, ref)
    A.run_ranks(1, fn)


def test_wrapped_host_array():
    def fn(a, r, w):
        src = np.arange(COUNT, dtype=np.float32)
        dst = np.zeros(COUNT, dtype=np.float32)
        s, d = a.wrap(src), a.wrap(dst)
        a.copy(s, d, COUNT)
        assert np.array_equal(src, dst)
    A.run_ranks(1, fn)


def test_nop_perf_counter_and_requests():
    def fn(a, r, w):
        req = a.nop()
        assert req.test()
        assert req.retcode() == 0
        assert 0 < req.duration_ns() < 50_000_000
        req.free()
        h = a.nop(run_async=True)
        assert h.wait_for(2000)
        h.free()
    A.run_ranks(1, fn)


def test_threshold_validation_and_reinit_guard():
    def fn(a, r, w):
        with pytest.raises(RuntimeError, match="EAGER_THRESHOLD_INVALID"):
            a.set_max_eager_msg_size(8)          # below the RX buffer size
        with pytest.raises(RuntimeError, match="RENDEZVOUS_THRESHOLD_INVALID"):
            a.set_max_rendezvous_msg_size(16)    # not above the eager threshold
        with pytest.raises(RuntimeError, match="appears configured"):
            a.initialize()
        with pytest.raises(ValueError):
            a.stream_put(a.create_buffer(4), 4, 0, 3)   # stream ids 0-8 are reserved
        assert "rank 0" in a.dump_communicator()
        assert "Spare RX Buffer 0" in a.dump_eager_rx_buffers()
        assert "exchange mem" in a.dump_exchange_memory()
    A.run_ranks(1, fn)


# ---------------------------------------------------------- point to point
@pytest.mark.parametrize("cfg", PROTOCOLS)
def test_sendrecv_basic(cfg):
    def fn(a, r, w):
        buf = a.create_buffer(COUNT)
        if r == 0:
            buf.host[:] = data(COUNT, 0)
            a.send(buf, COUNT, 1, tag=0)
        else:
            a.recv(buf, COUNT, 0, tag=0)
            assert torch.equal(buf.host, data(COUNT, 0))
    A.run_ranks(2, fn, cfg)


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
def test_sendrecv_ring(cfg, world):
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, r)
        nxt, prv = (r + 1) % w, (r - 1) % w
        # even ranks send first, odd ranks receive first: legal for both protocols
        if r % 2 == 0:
            req = a.send(s, COUNT, nxt, tag=7, run_async=True)
            a.recv(d, COUNT, prv, tag=7)
            req.wait()
        else:
            req = a.send(s, COUNT, nxt, tag=7, run_async=True)
            a.recv(d, COUNT, prv, tag=7)
            req.wait()
        assert req.retcode() == 0
        assert torch.equal(d.host, data(COUNT, prv))
    A.run_ranks(world, fn, cfg)


@pytest.mark.parametrize("k", [1, 2])
@pytest.mark.parametrize("delta", [-1, 0, 1])
def test_segmentation(k, delta):
    seg = EAGER["egr_rx_buf_size"] // 4
    count = k * seg + delta

    def fn(a, r, w):
        buf = a.create_buffer(count)
        if r == 0:
            buf.host[:] = data(count, 0)
            a.send(buf, count, 1, tag=1)
        else:
            a.recv(buf, count, 0, tag=1)
            assert torch.equal(buf.host, data(count, 0))
    A.run_ranks(2, fn, EAGER)


def test_tag_any_and_ordering():
    def fn(a, r, w):
        if r == 0:
            for i in range(4):
                b = a.create_buffer(16)
                b.host[:] = float(i)
                a.send(b, 16, 1, tag=10 + i)
        else:
            for i in range(4):  # TAG_ANY receives in send order
                b = a.create_buffer(16)
                a.recv(b, 16, 0)
                assert torch.all(b.host == float(i))
    A.run_ranks(2, fn, EAGER)


def test_sendrecv_stream_and_stream_put():
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, r)
        nxt, prv = (r + 1) % w, (r - 1) % w
        # memory -> network -> (stream, looped back) -> memory
        req = a.send(s, COUNT, nxt, tag=9, run_async=True)
        a.recv_to_stream(DataType.float32, COUNT, prv, tag=9)
        req.wait()
        a.copy_from_stream(d, COUNT)
        assert torch.equal(d.host, data(COUNT, prv))
        # memory -> stream -> network -> memory
        a.copy_to_stream(s, COUNT)
        req = a.send_from_stream(DataType.float32, COUNT, nxt, tag=11, run_async=True)
        a.recv(d, COUNT, prv, tag=11)
        req.wait()
        assert torch.equal(d.host, data(COUNT, prv))
        # one-sided put into the peer's stream 9, no matching recv
        a.barrier()
        a.stream_put(s, COUNT, nxt, 9)
        a.copy_from_stream(d, COUNT)
        assert torch.equal(d.host, data(COUNT, prv))
    A.run_ranks(2, fn, EAGER)


@pytest.mark.parametrize("wire", [torch.float16, torch.bfloat16])
def test_sendrecv_compressed(wire):
    def fn(a, r, w):
        buf = a.create_buffer(COUNT)
        if r == 0:
            buf.host[:] = data(COUNT, 0)
            a.send(buf, COUNT, 1, tag=3, compress_dtype=wire)
        else:
            a.recv(buf, COUNT, 0, tag=3, compress_dtype=wire)
            assert close(buf.host, data(COUNT, 0).to(wire).float(), 0, 0)
    A.run_ranks(2, fn, EAGER)


def test_sendrecv_fp8_block_scaled():
    def fn(a, r, w):
        buf = a.create_buffer(COUNT)
        if r == 0:
            buf.host[:] = data(COUNT, 0)
            a.send(buf, COUNT, 1, tag=3, compress_dtype=DataType.float8_e4m3)
        else:
            a.recv(buf, COUNT, 0, tag=3, compress_dtype=DataType.float8_e4m3)
            ref = data(COUNT, 0)
            # e4m3 has 3 mantissa bits: relative error <= 2^-4 of the block maximum
            blocks = ref.abs().view(-1)[: COUNT // 32 * 32].view(-1, 32).amax(dim=1).repeat_interleave(32)
            err = (buf.host - ref).abs()[: blocks.numel()]
            assert torch.all(err <= blocks * 2.0 ** -4 + 1e-6)
    A.run_ranks(2, fn, EAGER)


# -------------------------------------------------------------- collectives
@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
def test_bcast_every_root(cfg, world):
    def fn(a, r, w):
        for root in range(w):
            buf = a.create_buffer(COUNT)
            if r == root:
                buf.host[:] = data(COUNT, root, salt=root)
            a.bcast(buf, COUNT, root)
            assert torch.equal(buf.host, data(COUNT, root, salt=root))
    A.run_ranks(world, fn, cfg)


def test_bcast_binomial_tree_rendezvous():
    # 5 ranks > flat-tree limit (3): exercises the binomial tree
    def fn(a, r, w):
        for root in (0, 3):
            buf = a.create_buffer(COUNT)
            if r == root:
                buf.host[:] = data(COUNT, root)
            a.bcast(buf, COUNT, root)
            assert torch.equal(buf.host, data(COUNT, root))
    A.run_ranks(5, fn, RNDZV)


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
def test_scatter_gather_every_root(cfg, world):
    def fn(a, r, w):
        for root in range(w):
            send = a.create_buffer(COUNT * w)
            recv = a.create_buffer(COUNT)
            full = data(COUNT * w, root, salt=5)
            if r == root:
                send.host[:] = full
            a.scatter(send, recv, COUNT, root)
            assert torch.equal(recv.host, full[r * COUNT:(r + 1) * COUNT])
            out = a.create_buffer(COUNT * w)
            a.gather(recv, out, COUNT, root)
            if r == root:
                assert torch.equal(out.host, full)
    A.run_ranks(world, fn, cfg)


def test_gather_fanin_throttle():
    # large blocks: the rendezvous root admits only 2 peers at a time
    def fn(a, r, w):
        n = 12000  # 48 KB > 32 KB fan-in threshold
        s = a.create_buffer(n)
        s.host[:] = float(r + 1)
        out = a.create_buffer(n * w)
        a.gather(s, out, n, 1)
        if r == 1:
            assert torch.equal(out.host.view(w, n)[:, 0], torch.arange(1, w + 1, dtype=torch.float32))
    A.run_ranks(4, fn, dict(n_egr_rx_bufs=16, egr_rx_buf_size=64, max_egr_size=64, max_rndzv_size=1 << 20))


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
def test_allgather(cfg, world):
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT * w)
        s.host[:] = data(COUNT, r)
        a.allgather(s, d, COUNT)
        assert torch.equal(d.host, torch.cat([data(COUNT, q) for q in range(w)]))
    A.run_ranks(world, fn, cfg)


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("func", [SUM, MAX])
def test_reduce_every_root(cfg, world, func):
    def fn(a, r, w):
        for root in range(w):
            s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
            s.host[:] = data(COUNT, r, salt=root)
            a.reduce(s, d, COUNT, root, func)
            if r == root:
                assert close(d.host, reduce_ref(w, COUNT, func, salt=root), 1e-5, 1e-5)
    A.run_ranks(world, fn, cfg)


def test_reduce_binary_tree_rendezvous():
    # 6 ranks and 40 KB: beyond both flat-tree limits, chunked through the scratch buffers
    def fn(a, r, w):
        n = 10000
        for root in (0, 4):
            s, d = a.create_buffer(n), a.create_buffer(n)
            s.host[:] = data(n, r, salt=root)
            a.reduce(s, d, n, root, SUM)
            if r == root:
                assert close(d.host, reduce_ref(w, n, SUM, salt=root), 1e-5, 1e-4)
    A.run_ranks(6, fn, RNDZV)


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("func", [SUM, MAX])
def test_reduce_scatter(cfg, world, func):
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT * w), a.create_buffer(COUNT)
        s.host[:] = data(COUNT * w, r)
        a.reduce_scatter(s, d, COUNT, func)
        assert close(d.host, reduce_ref(w, COUNT * w, func)[r * COUNT:(r + 1) * COUNT], 1e-5, 1e-5)
    A.run_ranks(world, fn, cfg)


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("func", [SUM, MAX])
def test_allreduce(cfg, world, func):
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, r)
        a.allreduce(s, d, COUNT, func)
        assert close(d.host, reduce_ref(w, COUNT, func), 1e-5, 1e-5)
    A.run_ranks(world, fn, cfg)


@pytest.mark.parametrize("count", [1, 3, 5, 257])
def test_allreduce_awkward_counts(count):
    def fn(a, r, w):
        s, d = a.create_buffer(count), a.create_buffer(count)
        s.host[:] = data(count, r)
        a.allreduce(s, d, count, SUM)
        assert close(d.host, reduce_ref(w, count, SUM))
    A.run_ranks(4, fn, EAGER)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float64, torch.int32, torch.int64])
def test_allreduce_dtypes(dtype):
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT, dtype), a.create_buffer(COUNT, dtype)
        s.host[:] = data(COUNT, r, dtype)
        a.allreduce(s, d, COUNT, SUM)
        ref = reduce_ref(w, COUNT, SUM, dtype)
        tol = {torch.float16: 2e-2, torch.bfloat16: 1e-1}.get(dtype, 1e-9)
        assert close(d.host, ref, tol, tol)
    A.run_ranks(3, fn, EAGER)


@pytest.mark.parametrize("op", ["bcast", "reduce", "allreduce", "reduce_scatter", "allgather", "scatter", "gather"])
def test_collectives_compressed_fp16_wire(op):
    wire = torch.float16

    def fn(a, r, w):
        tol = dict(rtol=5e-3, atol=5e-2)  # the reference's tolerances for compressed runs
        if op == "bcast":
            b = a.create_buffer(COUNT)
            if r == 0:
                b.host[:] = data(COUNT, 0)
            a.bcast(b, COUNT, 0, compress_dtype=wire)
            assert close(b.host, data(COUNT, 0), **tol)
        elif op == "reduce":
            s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
            s.host[:] = data(COUNT, r)
            a.reduce(s, d, COUNT, 0, SUM, compress_dtype=wire)
            if r == 0:
                assert close(d.host, reduce_ref(w, COUNT, SUM), **tol)
        elif op == "allreduce":
            s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
            s.host[:] = data(COUNT, r)
            a.allreduce(s, d, COUNT, SUM, compress_dtype=wire)
            assert close(d.host, reduce_ref(w, COUNT, SUM), **tol)
        elif op == "reduce_scatter":
            s, d = a.create_buffer(COUNT * w), a.create_buffer(COUNT)
            s.host[:] = data(COUNT * w, r)
            a.reduce_scatter(s, d, COUNT, SUM, compress_dtype=wire)
            assert close(d.host, reduce_ref(w, COUNT * w, SUM)[r * COUNT:(r + 1) * COUNT], **tol)
        elif op == "allgather":
            s, d = a.create_buffer(COUNT), a.create_buffer(COUNT * w)
            s.host[:] = data(COUNT, r)
            a.allgather(s, d, COUNT, compress_dtype=wire)
            assert close(d.host, torch.cat([data(COUNT, q) for q in range(w)]), **tol)
        elif op == "scatter":
            s, d = a.create_buffer(COUNT * w), a.create_buffer(COUNT)
            if r == 0:
                s.host[:] = data(COUNT * w, 0)
            a.scatter(s, d, COUNT, 0, compress_dtype=wire)
            assert close(d.host, data(COUNT * w, 0)[r * COUNT:(r + 1) * COUNT], **tol)
        else:
            s, d = a.create_buffer(COUNT), a.create_buffer(COUNT * w)
            s.host[:] = data(COUNT, r)
            a.gather(s, d, COUNT, 0, compress_dtype=wire)
            if r == 0:
                assert close(d.host, torch.cat([data(COUNT, q) for q in range(w)]), **tol)
    A.run_ranks(3, fn, EAGER)


def test_mixed_dtype_operands():
    # fp32 source, fp16 result buffer: the narrower type is the "compressed" operand
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT, torch.float32), a.create_buffer(COUNT, torch.float16)
        s.host[:] = data(COUNT, r)
        a.allreduce(s, d, COUNT, SUM)
        assert close(d.host, reduce_ref(w, COUNT, SUM), 5e-3, 5e-2)
    A.run_ranks(2, fn, EAGER)


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", WORLDS)
def test_alltoall(cfg, world):
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT * w), a.create_buffer(COUNT * w)
        s.host[:] = data(COUNT * w, r)
        a.alltoall(s, d, COUNT)
        ref = torch.cat([data(COUNT * w, q)[r * COUNT:(r + 1) * COUNT] for q in range(w)])
        assert torch.equal(d.host, ref)
    A.run_ranks(world, fn, cfg)


@pytest.mark.parametrize("world", [1, 2, 4])
def test_barrier(world):
    import time
    stamps = {}

    def fn(a, r, w):
        if r == w - 1:
            time.sleep(0.2)
        a.barrier()
        stamps[r] = time.time()
        a.barrier()
    A.run_ranks(world, fn, EAGER)
    assert max(stamps.values()) - min(stamps.values()) < 0.15


def test_reduce_stream_variants():
    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, r)
        ref = reduce_ref(w, COUNT, SUM)
        # stream -> memory
        a.copy_to_stream(s, COUNT)
        a.reduce_stream2mem(DataType.float32, d, COUNT, 0, SUM)
        if r == 0:
            assert close(d.host, ref)
        # memory -> stream
        a.reduce_mem2stream(s, DataType.float32, COUNT, 0, SUM)
        if r == 0:
            a.copy_from_stream(d, COUNT)
            assert close(d.host, ref)
        # stream -> stream
        a.copy_to_stream(s, COUNT)
        a.reduce_stream2stream(DataType.float32, DataType.float32, COUNT, 0, SUM)
        if r == 0:
            a.copy_from_stream(d, COUNT)
            assert close(d.host, ref)
    A.run_ranks(2, fn, EAGER)


# --------------------------------------------------------- communicators
@pytest.mark.parametrize("cfg", PROTOCOLS)
def test_allgather_on_split_communicators(cfg):
    def fn(a, r, w):
        ranks = A.Accl.generate_ranks(w)
        half = w // 2
        group = list(range(half)) if r < half else list(range(half, w))
        comm = a.create_communicator([ranks[g] for g in group], group.index(r))
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT * len(group))
        s.host[:] = data(COUNT, r)
        a.allgather(s, d, COUNT, comm_id=comm)
        assert torch.equal(d.host, torch.cat([data(COUNT, q) for q in group]))
        assert a.get_comm_rank(comm) == group.index(r)
        assert len(a.get_comm_group(comm)) == len(group)
    A.run_ranks(4, fn, cfg)


def test_multicomm_subgroup():
    # 3-of-4 subgroup: p2p and allreduce inside it while rank 3 stays out
    def fn(a, r, w):
        ranks = A.Accl.generate_ranks(w)
        group = [0, 1, 2]
        if r not in group:
            a.barrier()
            return
        comm = a.create_communicator([ranks[g] for g in group], group.index(r))
        me = group.index(r)
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, r)
        if me == 0:
            a.send(s, COUNT, 1, tag=2, comm_id=comm)
        elif me == 1:
            a.recv(d, COUNT, 0, tag=2, comm_id=comm)
            assert torch.equal(d.host, data(COUNT, 0))
        a.allreduce(s, d, COUNT, SUM, comm_id=comm)
        assert close(d.host, reduce_ref(3, COUNT, SUM))
        # the global communicator still works and has its own sequence space
        a.barrier()
        assert "outbound_seq" in a.dump_communicator()
    A.run_ranks(4, fn, EAGER)


def test_rendezvous_send_parks_in_retry_queue():
    # a rendezvous send issued before the matching recv is posted must not
    # block the engine: a later call on the same rank still completes
    def fn(a, r, w):
        big, small = a.create_buffer(4000), a.create_buffer(8)
        if r == 0:
            big.host[:] = data(4000, 0)
            req = a.send(big, 4000, 1, tag=1, run_async=True)
            a.nop()                                  # overtakes the parked send
            assert not req.test() or True            # (may already be done if rank 1 was fast)
            req.wait()
            assert req.retcode() == 0
        else:
            import time
            time.sleep(0.1)
            a.recv(big, 4000, 0, tag=1)
            assert torch.equal(big.host, data(4000, 0))
    A.run_ranks(2, fn, RNDZV)


def test_device_issued_call_through_second_port():
    # the client arbiter: a "kernel" issues a 15-word command itself
    def fn(a, r, w):
        s, d = a.create_buffer(64), a.create_buffer(64)
        s.host[:] = data(64, 0)
        s.sync_to_device()
        words = [1, 64, 0, 0, 0, A.TAG_ANY, a.get_arithmetic_config_addr(DataType.float32, DataType.float32), 0, 0,
                 s.address & 0xFFFFFFFF, s.address >> 32, 0, 0, d.address & 0xFFFFFFFF, d.address >> 32]
        assert a.emu_device_call(words) == 0
        d.sync_from_device()
        assert torch.equal(s.host, d.host)
    A.run_ranks(1, fn)


# ------------------------------------------------------------ stress (reference test/host/xrt/src/stress.cpp:24-33)
@pytest.mark.parametrize("cfg", PROTOCOLS)
def test_stress_sendrecv_ring(cfg):
    iters = 400

    def fn(a, r, w):
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        nxt, prv = (r + 1) % w, (r - 1) % w
        for i in range(iters):
            s.host[:] = data(COUNT, r, salt=i & 7)
            if r % 2 == 0:
                a.send(s, COUNT, nxt, tag=i & 0xFF)
                a.recv(d, COUNT, prv, tag=i & 0xFF)
            else:
                a.recv(d, COUNT, prv, tag=i & 0xFF)
                a.send(s, COUNT, nxt, tag=i & 0xFF)
            if i % 50 == 0:
                assert torch.equal(d.host, data(COUNT, prv, salt=i & 7))
        assert torch.equal(d.host, data(COUNT, prv, salt=(iters - 1) & 7))
    A.run_ranks(4, fn, cfg)


def test_call_trace_file(tmp_path):
    """ACCL_TRACE=<prefix>: Chrome trace-event JSON with engine-measured durations."""
    import json
    import os
    import subprocess
    import sys
    prefix = str(tmp_path / "trace_")
    code = ("import accl_b200 as A\n"
            "def fn(a, r, w):\n"
            "    s, d = a.create_buffer(256), a.create_buffer(256)\n"
            "    s.host[:] = r\n"
            "    a.allreduce(s, d, 256, A.SUM)\n"
            "    q = a.bcast(s, 256, 0, run_async=True); q.wait(); q.free()\n"
            "A.run_ranks(2, fn)\n")
    env = dict(os.environ, ACCL_TRACE=prefix, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env.pop("RANK", None)
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=120)
    ev = json.load(open(prefix + "0.json"))["traceEvents"]
    names = [(e["pid"], e["name"]) for e in ev]
    for r in (0, 1):
        assert (r, "allreduce") in names and (r, "bcast") in names
    done = [e for e in ev if e["name"] in ("allreduce", "bcast")]
    assert all(e["args"]["completed"] and e["args"]["retcode"] == 0 and e["args"]["engine_ns"] > 0 for e in done)


# ------------------------------------------------------------ device-side API through the BFM hooks
# (reference test/host/hls_simulator/test.cpp:54-250, run against an emulator without kernel loopback)
OP_SEND, OP_RECV, OP_REDUCE = 3, 4, 8
OP0_STREAM, RES_STREAM = 1, 2


def _desc(a, scenario, count, root_src_dst=0, function=0, tag=A.TAG_ANY, stream_flags=0, addr0=0, addr2=0):
    """The 15-word descriptor a kernel emits with accl::device::Command::start_call."""
    return [scenario, count, 0, root_src_dst, function, tag, a.get_arithmetic_config_addr(DataType.float32, DataType.float32),
            0, stream_flags, addr0 & 0xFFFFFFFF, addr0 >> 32, 0, 0, addr2 & 0xFFFFFFFF, addr2 >> 32]


def _f32(t):
    return t.numpy().astype(np.float32).tobytes()


def test_bfm_vadd_put():
    """A 'user kernel' adds 1 to its input and stream_puts the result into the next rank's stream 9, issuing
    the command itself (reference vadd_put.cpp:25-86, hls_simulator/test.cpp:54)."""
    n = 64

    def fn(a, r, w):
        a.emu_set_kernel_loopback(False)
        x = data(n, r)
        a.barrier()
        # kernel body: compute, push the data words, then the command (send, operand from stream, result to stream 9)
        a.emu_kernel_push(_f32(x + 1))
        assert a.emu_device_call(_desc(a, OP_SEND, n, root_src_dst=(r + 1) % w, tag=9, stream_flags=OP0_STREAM | RES_STREAM)) == 0
        got = np.frombuffer(a.emu_kernel_pull(9, n * 4), dtype=np.float32)
        assert np.array_equal(got, (data(n, (r - 1) % w) + 1).numpy())
    A.run_ranks(2, fn, EAGER)


def test_bfm_loopback_through_user_kernel():
    """send -> recv-to-stream -> user kernel loop -> send-from-stream -> recv (hls_simulator/test.cpp:153)."""
    n = 128

    def fn(a, r, w):
        a.emu_set_kernel_loopback(False)
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.host[:] = data(n, r)
        nxt, prv = (r + 1) % w, (r - 1) % w
        req = a.send(s, n, nxt, tag=5, run_async=True)
        a.recv_to_stream(DataType.float32, n, prv, tag=5)          # network -> stream <tag> (the kernel's input)
        req.wait()
        words = a.emu_kernel_pull(5, n * 4)                        # the user kernel: read, transform, write back
        a.emu_kernel_push((np.frombuffer(words, dtype=np.float32) * 2).tobytes())
        req = a.send_from_stream(DataType.float32, n, nxt, tag=6, run_async=True)
        a.recv(d, n, prv, tag=6)
        req.wait()
        assert torch.equal(d.host, data(n, (r - 2) % w) * 2)
    A.run_ranks(3, fn, EAGER)


def test_bfm_reduce_stream_to_stream():
    """Every rank's kernel feeds the reduction from its output stream; the root's kernel receives the result
    on its input stream (hls_simulator/test.cpp:199 `test_reduce_stream`)."""
    n = 96

    def fn(a, r, w):
        a.emu_set_kernel_loopback(False)
        a.emu_kernel_push(_f32(data(n, r)))
        a.reduce_stream2stream(DataType.float32, DataType.float32, n, 1, SUM)
        if r == 1:
            got = torch.from_numpy(np.frombuffer(a.emu_kernel_pull(0, n * 4), dtype=np.float32).copy())
            assert close(got, reduce_ref(w, n, SUM), 1e-5, 1e-5)
    A.run_ranks(3, fn, EAGER)


def test_rank_table_helpers(tmp_path):
    """generate_ranks / get_ips (reference accl_network_utils.cpp:394-449): IP list, JSON rank file, local default."""
    cfg = tmp_path / "ranks.json"
    cfg.write_text('{"ips": ["10.1.0.1", "10.1.0.2",\\n "10.1.0.3"], "other": [1, 2]}')
    assert A._C.get_ips(str(cfg)) == ["10.1.0.1", "10.1.0.2", "10.1.0.3"]
    ranks = A.Accl.generate_ranks(config_file=cfg, base_port=6000, max_segment_size=2048)
    assert [(r.ip, r.port, r.session_id, r.max_segment_size) for r in ranks] == [
        ("10.1.0.1", 6000, 0, 2048), ("10.1.0.2", 6001, 1, 2048), ("10.1.0.3", 6002, 2, 2048)]
    local = A.Accl.generate_ranks(4)
    assert [r.ip for r in local] == ["127.0.0.1"] * 4 and [r.port for r in local] == [5500, 5501, 5502, 5503]
    (tmp_path / "bad.json").write_text('{"nodes": []}')
    with pytest.raises(RuntimeError):
        A._C.get_ips(str(tmp_path / "bad.json"))

    # an explicit table drives initialize() like the synthetic one
    def fn(a, r, w):
        s, d = a.create_buffer(32), a.create_buffer(32)
        s.host[:] = r + 1
        a.allreduce(s, d, 32, SUM)
        assert torch.all(d.host == 3)
    accls = A.emulator_world(2)
    import threading
    errs = []

    def body(r):
        try:
            accls[r].initialize(A.Accl.generate_ranks(ips=["127.0.0.1", "127.0.0.1"]), r)
            fn(accls[r], r, 2)
        except BaseException as e:  # noqa: BLE001
            errs.append(repr(e))
    ts = [threading.Thread(target=body, args=(r,)) for r in range(2)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errs, errs


# ------------------------------------------------------------ failure detection
def test_receive_timeout_and_error_decoding():
    """A recv nobody answers fails with RECEIVE_TIMEOUT_ERROR after `set_timeout`; the exception names the bits."""
    def fn(a, r, w):
        a.set_timeout(20000)                          # 20 ms in engine ticks (us)
        if r == 0:
            d = a.create_buffer(16)
            with pytest.raises(RuntimeError, match="RECEIVE_TIMEOUT_ERROR"):
                a.recv(d, 16, 1, tag=77)
        a.barrier()                                   # the engine is still usable afterwards
    A.run_ranks(2, fn, EAGER)


def test_wait_with_timeout_and_soft_reset_drains_parked_calls():
    """A rendezvous send without a receiver parks in the retry queue: wait(timeout) reports 'not yet', and
    soft_reset hands the waiter NOT_READY_ERROR instead of leaving it hanging (reference fw:2249-2261)."""
    def fn(a, r, w):
        if r == 0:
            big = a.create_buffer(4000)
            req = a.send(big, 4000, 1, tag=9, run_async=True)
            assert req.wait_for(50) is False          # still parked after 50 ms
            assert not req.test()
            a.impl.soft_reset()
            assert req.wait_for(2000) is True
            assert req.retcode() != 0
            assert "NOT_READY" in A._C.error_to_string(req.retcode())
    A.run_ranks(2, fn, RNDZV)


def test_bad_arguments_are_rejected():
    def fn(a, r, w):
        s, d = a.create_buffer(8), a.create_buffer(8)
        with pytest.raises(RuntimeError):
            a.bcast(s, 8, 7)                          # root outside the communicator
        with pytest.raises((RuntimeError, IndexError)):
            a.allreduce(s, d, 8, SUM, comm_id=5)      # communicator never created
        with pytest.raises((RuntimeError, IndexError, ValueError)):
            a.copy(s, d, 64)                          # count beyond the buffers
        a.barrier()
    A.run_ranks(2, fn, EAGER)


def test_same_membership_communicators_do_not_cross_match():
    """Two communicators over identical members have separate sequence spaces: a receive on one must not take a
    message of the other even if that one arrived first with the same (source, tag, sequence number).  Found by
    the sub-communicator property test (all ranks drawn into the 'sub'-communicator) under load."""
    import time

    def fn(a, r, w):
        c2 = a.create_communicator(a.get_comm_group(0), r)
        x, y = a.create_buffer(16), a.create_buffer(16)
        if r == 0:
            x.host[:] = 1.0
            y.host[:] = 2.0
            q1 = a.send(x, 16, 1, tag=5, comm_id=c2, run_async=True)
            q2 = a.send(y, 16, 1, tag=5, run_async=True)
            q1.wait()
            q2.wait()
        else:
            time.sleep(0.2)                      # both messages sit in the rx pool
            a.recv(y, 16, 0, tag=5)              # global communicator first, although its message arrived second
            a.recv(x, 16, 0, tag=5, comm_id=c2)
            assert torch.all(y.host == 2.0) and torch.all(x.host == 1.0)
    A.run_ranks(2, fn, EAGER)


def test_parked_sends_keep_issue_order_and_do_not_starve_under_blocking_receives():
    """Found by the point-to-point property test.  (1) Several rendezvous sends parked for the same peer and tag are
    served in issue order (non-overtaking).  (2) A rank blocked in an eager receive still lets its parked rendezvous
    sends go out: the peer needs them before it can send what the receive is waiting for."""
    cfg = dict(n_egr_rx_bufs=8, egr_rx_buf_size=1024, max_egr_size=1024, max_rndzv_size=2048)

    def fn(a, r, w):
        big = [a.create_buffer(1232) for _ in range(3)]
        small = a.create_buffer(64)
        if r == 0:
            for i, b in enumerate(big):
                b.host[:] = float(i + 1)
            reqs = [a.send(b, 1232, 1, tag=69, run_async=True) for b in big]   # three parked sends, same peer, same tag
            a.recv(small, 64, 1, tag=7)          # eager receive: rank 1 only sends this after it has all three
            assert torch.all(small.host == 9.0)
            for q in reqs:
                q.wait()
                assert q.retcode() == 0
        else:
            import time
            time.sleep(0.05)                     # rank 0 is already blocked in its receive
            for i, b in enumerate(big):
                a.recv(b, 1232, 0, tag=69)
                assert torch.all(b.host == float(i + 1)), (i, b.host[:2])
            small.host[:] = 9.0
            a.send(small, 64, 0, tag=7)
        a.barrier()
    for _ in range(5):
        A.run_ranks(2, fn, cfg)


def test_notes_of_collectives_and_point_to_point_do_not_cross_match():
    """A parked rendezvous send (tag t) and a rendezvous collective (TAG_ANY notes) wait for the same peer at the same
    time: the collective must not take the address note of the peer's receive, nor the send the collective's.
    Found by the mixed point-to-point / collective property test."""
    cfg = dict(n_egr_rx_bufs=8, egr_rx_buf_size=256, max_egr_size=1024, max_rndzv_size=1 << 20)
    n, m = 2460, 568

    def fn(a, r, w):
        a.set_timeout(30_000_000)
        big, got = a.create_buffer(n), a.create_buffer(n)
        s, d = a.create_buffer(m * w), a.create_buffer(m * w)
        s.host[:] = torch.arange(m * w, dtype=torch.float32) + 1000 * r
        req = None
        if r == 0:
            big.host[:] = 7.0
            req = a.send(big, n, 1, tag=72, run_async=True)    # parks: rank 1 posts the receive late
        if r == 1:
            import time
            time.sleep(0.05)
            a.recv(got, n, 0, tag=72)
            assert torch.all(got.host == 7.0)
        a.alltoall(s, d, m)                                      # rendezvous, address notes carry TAG_ANY
        for q in range(w):
            assert torch.equal(d.host[q * m:(q + 1) * m], torch.arange(r * m, (r + 1) * m, dtype=torch.float32) + 1000 * q)
        if req is not None:
            req.wait()
            assert req.retcode() == 0
        a.barrier()
    for _ in range(5):
        A.run_ranks(3, fn, cfg)
