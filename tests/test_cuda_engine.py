"""GPU tests of the persistent engine (csrc/src/cuda/engine.cu): every call is a command in the engine's device
ring, executed by the resident kernel as a resumable state machine — calls that cannot progress are parked with
their `step` and the next one is tried (reference: retry queue, ccl_offload_control.c:2264-2288, 2460-2478).
Covers: the whole collective matrix through the engine, rendezvous send/send-then-recv/recv (parked sends), tag
matching out of order, eager segments with credits, device-measured NOP latency (reference perf_counter test,
test/host/xrt/src/test.cpp:1137-1152), parked-call timeouts, device-issued calls while host calls are in flight."""
import pytest
import torch

import accl_b200 as A
from accl_b200 import MAX, SUM

pytestmark = pytest.mark.gpu
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
EAGER = dict(n_egr_rx_bufs=4, egr_rx_buf_size=16 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)
RNDZV = dict(n_egr_rx_bufs=4, egr_rx_buf_size=1024, max_egr_size=1024, max_rndzv_size=1 << 30)
PROTOCOLS = [pytest.param(EAGER, id="eager"), pytest.param(RNDZV, id="rndzv")]
COUNT = 5000


def devices(world):
    return [r % max(NGPU, 1) for r in range(world)]


def data(count, rank, dtype=torch.float32, salt=0):
    g = torch.Generator().manual_seed(555 + 17 * rank + salt)
    return (torch.rand(count, generator=g, dtype=torch.float32) * 8 - 4).to(dtype)


def ref_reduce(world, count, func, salt=0):
    xs = [data(count, r, salt=salt).double() for r in range(world)]
    out = xs[0].clone()
    for x in xs[1:]:
        out = out + x if func == SUM else torch.maximum(out, x)
    return out


def close(a, b, rtol=1e-5, atol=1e-4):
    return torch.allclose(a.cpu().double(), b.cpu().double(), rtol=rtol, atol=atol)


def run(world, fn, cfg=EAGER, **kw):
    opts = dict(heap_mb=128, max_ctas=8, engine=True, engine_workers=8)
    opts.update(kw)
    return A.run_cuda_ranks(devices(world), fn, cfg, **opts)


def test_nop_latency_is_device_measured():
    def fn(a, r, w):
        assert "mode=engine" in a.describe()
        medians = []
        for _ in range(3):   # a shared box can disturb one round; the claim is about the engine, not about the neighbours
            durs = []
            for _ in range(50):
                q = a.nop()
                assert q.retcode() == 0
                durs.append(q.duration_ns())
            durs.sort()
            medians.append(durs[len(durs) // 2])
            if 0 < medians[-1] <= 2000:
                break
        # fetched -> retired inside the resident kernel: the reference accepts 100 ns - 1 us on its 250 MHz soft CPU
        assert 0 < min(medians) <= 2000, medians
        return min(medians)
    print("engine NOP ns:", run(1, fn))


@pytest.mark.parametrize("cfg", PROTOCOLS)
@pytest.mark.parametrize("world", [2, 3])
def test_collective_matrix(cfg, world):
    def fn(a, r, w):
        n = COUNT
        for func in (SUM, MAX):
            s, d = a.create_buffer(n), a.create_buffer(n)
            s.host[:] = data(n, r)
            a.allreduce(s, d, n, func)
            assert close(d.host, ref_reduce(w, n, func)), "allreduce"
        s, d = a.create_buffer(n * w), a.create_buffer(n)
        s.host[:] = data(n * w, r)
        a.reduce_scatter(s, d, n, SUM)
        assert close(d.host, ref_reduce(w, n * w, SUM)[r * n:(r + 1) * n]), "reduce_scatter"
        g = a.create_buffer(n * w)
        a.allgather(d, g, n)
        assert close(g.host, ref_reduce(w, n * w, SUM)), "allgather"
        t = a.create_buffer(n * w)
        a.alltoall(s, t, n)
        assert torch.equal(t.host, torch.cat([data(n * w, q)[r * n:(r + 1) * n] for q in range(w)])), "alltoall"
        for root in range(w):
            b = a.create_buffer(n)
            if r == root:
                b.host[:] = data(n, root, salt=3)
            a.bcast(b, n, root)
            assert torch.equal(b.host, data(n, root, salt=3)), "bcast"
            recv = a.create_buffer(n)
            full = data(n * w, root, salt=9)
            if r == root:
                s.host[:] = full
            a.scatter(s, recv, n, root)
            assert torch.equal(recv.host, full[r * n:(r + 1) * n]), "scatter"
            out = a.create_buffer(n * w)
            a.gather(recv, out, n, root)
            if r == root:
                assert torch.equal(out.host, full), "gather"
            x, y = a.create_buffer(n), a.create_buffer(n)
            x.host[:] = data(n, r, salt=root)
            a.reduce(x, y, n, root, SUM)
            if r == root:
                assert close(y.host, ref_reduce(w, n, SUM, salt=root)), "reduce"
        a.barrier()
        c1, c2, c3 = a.create_buffer(n), a.create_buffer(n), a.create_buffer(n)
        c1.host[:] = data(n, 1)
        c2.host[:] = data(n, 2)
        a.copy(c1, c3, n)
        assert torch.equal(c1.host, c3.host)
        a.combine(n, SUM, c1, c2, c3)
        assert close(c3.host, c1.host.double() + c2.host.double())
    run(world, fn, cfg)


def test_large_rendezvous_collectives_run_as_moves():
    def fn(a, r, w):
        n = (1 << 20) + 3
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.dev.copy_(data(n, r).cuda(a.cuda_device))
        for _ in range(3):
            a.allreduce(s, d, n, SUM, from_fpga=True, to_fpga=True)
        torch.cuda.current_stream().synchronize()
        assert close(d.dev, ref_reduce(w, n, SUM))
        a.allreduce(s, s, n, SUM, from_fpga=True, to_fpga=True)  # in place
        torch.cuda.current_stream().synchronize()
        assert close(s.dev, ref_reduce(w, n, SUM))
    run(2, fn, RNDZV)


@pytest.mark.parametrize("cfg", PROTOCOLS)
def test_send_send_then_recv_recv(cfg):
    # both ranks send first: with rendezvous the sends cannot complete before the receives are posted — the engine
    # parks them (NOT_READY, step kept) and executes the receives that follow in its queue
    def fn(a, r, w):
        peer = 1 - r
        s, d = a.create_buffer(COUNT), a.create_buffer(COUNT)
        s.host[:] = data(COUNT, r)
        q1 = a.send(s, COUNT, peer, tag=5, run_async=True)
        q2 = a.recv(d, COUNT, peer, tag=5, run_async=True)
        q1.wait()
        q2.wait()
        assert q1.retcode() == 0 and q2.retcode() == 0
        d.sync_from_device()
        assert torch.equal(d.host, data(COUNT, peer))
        return a.cuda_debug_state()
    out = run(2, fn, cfg)
    if cfg is RNDZV:
        assert "parks=" in out[0]


def test_rendezvous_tag_matching_out_of_order():
    # sends issued in the opposite order of the receives: matched by tag through the address mailbox
    def fn(a, r, w):
        n = 4096
        if r == 0:
            s1, s2 = a.create_buffer(n), a.create_buffer(n)
            s1.host[:] = data(n, 0, salt=1)
            s2.host[:] = data(n, 0, salt=2)
            q = [a.send(s2, n, 1, tag=2, run_async=True), a.send(s1, n, 1, tag=1, run_async=True)]
        else:
            d1, d2 = a.create_buffer(n), a.create_buffer(n)
            q = [a.recv(d1, n, 0, tag=1, run_async=True), a.recv(d2, n, 0, tag=2, run_async=True)]
        for x in q:
            x.wait()
            assert x.retcode() == 0
        if r == 1:
            d1.sync_from_device()
            d2.sync_from_device()
            assert torch.equal(d1.host, data(n, 0, salt=1)) and torch.equal(d2.host, data(n, 0, salt=2))
    run(2, fn, RNDZV)


def test_many_outstanding_point_to_point_calls():
    # more rendezvous receives in flight than mailbox slots per pair, plus eager traffic in between
    def fn(a, r, w):
        n, k = 3000, 12
        bufs = [a.create_buffer(n) for _ in range(k)]
        reqs = []
        for i in range(k):
            if r == 0:
                bufs[i].host[:] = data(n, 0, salt=i)
                reqs.append(a.send(bufs[i], n, 1, tag=100 + i, run_async=True))
            else:
                reqs.append(a.recv(bufs[i], n, 0, tag=100 + i, run_async=True))
        small = a.create_buffer(16)
        small.host[:] = float(r + 1)
        res = a.create_buffer(16)
        a.allreduce(small, res, 16, SUM)
        assert float(res.host[0]) == 3.0
        for q in reqs:
            q.wait()
            assert q.retcode() == 0
        if r == 1:
            for i in range(k):
                bufs[i].sync_from_device()
                assert torch.equal(bufs[i].host, data(n, 0, salt=i)), i
    run(2, fn, RNDZV)


def test_eager_segments_park_on_credits():
    # 40 KB message through 4 slots of 4 KB: the sender runs out of credits until the receiver consumes
    def fn(a, r, w):
        n = 10000
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.host[:] = data(n, r)
        if r == 0:
            q = a.send(s, n, 1, tag=7, run_async=True)
        else:
            q = a.recv(d, n, 0, tag=7, run_async=True)
        q.wait()
        assert q.retcode() == 0
        if r == 1:
            d.sync_from_device()
            assert torch.equal(d.host, data(n, 0))
    run(2, fn, dict(n_egr_rx_bufs=4, egr_rx_buf_size=4096, max_egr_size=64 << 10, max_rndzv_size=1 << 30))


def test_parked_call_times_out():
    # a receive nobody answers: retired with RECEIVE_TIMEOUT_ERROR after the engine's wait budget, engine stays usable
    def fn(a, r, w):
        a.set_timeout(4000)  # x 32 us = 128 ms
        d = a.create_buffer(64)
        if r == 0:
            q = a.recv(d, 64, 1, tag=1, run_async=True)
            q.wait()
            assert q.retcode() & 0x800, hex(q.retcode())  # RECEIVE_TIMEOUT_ERROR (bit 11)
        a.set_timeout(1000000)
        s = a.create_buffer(8)
        s.host[:] = 1.0
        a.allreduce(s, s, 8, SUM)
        assert float(s.host[0]) == float(w)
    run(2, fn, RNDZV)


def test_device_issued_calls_next_to_host_calls():
    from accl_b200.ops import vadd_allreduce
    n = 300000

    def fn(a, r, w):
        x, y, out = a.create_buffer(n), a.create_buffer(n), a.create_buffer(n)
        x.dev.copy_(data(n, r).cuda(a.cuda_device))
        y.dev.fill_(1.0)
        st = vadd_allreduce(a, x, y, out, chunk_elems=65536)       # 5 chunks, each handed to the engine by the kernel
        s = a.create_buffer(100)
        s.host[:] = float(r)
        a.allreduce(s, s, 100, SUM)                                # a host call queued behind the plugin's
        torch.cuda.current_stream().synchronize()
        assert int(st.item()) == 0, hex(int(st.item()))
        assert close(out.dev, ref_reduce(w, n, SUM) + w)
        assert float(s.host[0]) == sum(range(w))
    run(2, fn, RNDZV)


@pytest.mark.parametrize("mode", ["direct", "engine"])
def test_torch_distributed_backend_on_gpus(mode):
    """`dist.init_process_group("accl")` with CUDA tensors, one process per rank (torchrun): the collectives,
    point to point, sub-groups and a DistributedDataParallel step (tests/helpers/pg_worker.py)."""
    import os
    import random
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = 2
    env = dict(os.environ, PG_PORT=str(random.randint(20000, 30000)), CUDA_DEVICE_MAX_CONNECTIONS="32",
               ACCL_PG_ENGINE="1" if mode == "engine" else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                        "127.0.0.1", "--master-port", str(random.randint(30001, 40000)),
                        os.path.join(root, "tests", "helpers", "pg_worker.py")], cwd=root, env=env, capture_output=True, text=True,
                       timeout=200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count(": ok") == world
