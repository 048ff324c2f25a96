"""Property tests (hypothesis) on the CPU emulator: random counts, world sizes, roots, reduce functions and —
most importantly — random eager / rendezvous geometry (rx buffer size, eager threshold, rendezvous segment size),
so that segmentation boundaries, protocol switch-over points and non-divisible counts are hit from every side.
Results are compared with a plain torch reference.  The reference's suite probes only k * segment +- 1
(test/host/xrt/src/test.cpp:345-393)."""
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import accl_b200 as A
from accl_b200 import MAX, SUM

OPS = ["sendrecv", "bcast", "scatter", "gather", "allgather", "reduce", "allreduce", "reduce_scatter", "alltoall"]


def data(n, r, salt):
    g = torch.Generator().manual_seed(977 * salt + 31 * r + 5)
    return torch.randint(-64, 64, (n,), generator=g).float()  # exactly representable: sums compare with ==


@st.composite
def geometry(draw):
    buf = draw(st.sampled_from([64, 128, 256, 1024, 4096]))
    max_egr = draw(st.sampled_from([64, 256, 1024, 4096, 1 << 20]))
    # the engine needs an eager message to fit the rx pool: keep threshold <= pool capacity
    n_bufs = draw(st.sampled_from([8, 16, 32]))
    max_egr = min(max_egr, buf * n_bufs // 2) if max_egr < (1 << 20) else max_egr
    max_egr = max(max_egr, buf)      # the engine rejects an eager threshold below one rx buffer
    rndzv = draw(st.sampled_from([256, 1024, 32 * 1024, 1 << 20]))
    rndzv = max(rndzv, 2 * max_egr)  # the engine rejects a rendezvous segment size below the eager threshold
    return dict(n_egr_rx_bufs=n_bufs, egr_rx_buf_size=buf, max_egr_size=max_egr, max_rndzv_size=rndzv)


def run_op(a, r, w, op, count, root, func, salt):
    def red(vs):
        out = vs[0].clone()
        for v in vs[1:]:
            out = out + v if func == SUM else torch.maximum(out, v)
        return out

    if op == "sendrecv":
        s, d = a.create_buffer(count), a.create_buffer(count)
        s.host[:] = data(count, r, salt)
        nxt, prv = (r + 1) % w, (r - 1) % w
        req = a.send(s, count, nxt, tag=salt & 0xFF, run_async=True)
        a.recv(d, count, prv, tag=salt & 0xFF)
        req.wait()
        assert torch.equal(d.host, data(count, prv, salt))
    elif op == "bcast":
        b = a.create_buffer(count)
        b.host[:] = data(count, r, salt)
        a.bcast(b, count, root)
        assert torch.equal(b.host, data(count, root, salt))
    elif op == "scatter":
        s, d = a.create_buffer(count * w), a.create_buffer(count)
        s.host[:] = data(count * w, r, salt)
        a.scatter(s, d, count, root)
        assert torch.equal(d.host, data(count * w, root, salt)[r * count:(r + 1) * count])
    elif op == "gather":
        s, d = a.create_buffer(count), a.create_buffer(count * w)
        s.host[:] = data(count, r, salt)
        a.gather(s, d, count, root)
        if r == root:
            assert torch.equal(d.host, torch.cat([data(count, q, salt) for q in range(w)]))
    elif op == "allgather":
        s, d = a.create_buffer(count), a.create_buffer(count * w)
        s.host[:] = data(count, r, salt)
        a.allgather(s, d, count)
        assert torch.equal(d.host, torch.cat([data(count, q, salt) for q in range(w)]))
    elif op == "reduce":
        s, d = a.create_buffer(count), a.create_buffer(count)
        s.host[:] = data(count, r, salt)
        a.reduce(s, d, count, root, func)
        if r == root:
            assert torch.equal(d.host, red([data(count, q, salt) for q in range(w)]))
    elif op == "allreduce":
        s, d = a.create_buffer(count), a.create_buffer(count)
        s.host[:] = data(count, r, salt)
        a.allreduce(s, d, count, func)
        assert torch.equal(d.host, red([data(count, q, salt) for q in range(w)]))
    elif op == "reduce_scatter":
        s, d = a.create_buffer(count * w), a.create_buffer(count)
        s.host[:] = data(count * w, r, salt)
        a.reduce_scatter(s, d, count, func)
        assert torch.equal(d.host, red([data(count * w, q, salt) for q in range(w)])[r * count:(r + 1) * count])
    elif op == "alltoall":
        s, d = a.create_buffer(count * w), a.create_buffer(count * w)
        s.host[:] = data(count * w, r, salt)
        a.alltoall(s, d, count)
        exp = torch.cat([data(count * w, q, salt)[r * count:(r + 1) * count] for q in range(w)])
        assert torch.equal(d.host, exp)
    else:
        a.barrier()


step = st.tuples(st.sampled_from(OPS + ["barrier"]), st.integers(1, 3000), st.integers(0, 5), st.sampled_from([SUM, MAX]),
                 st.integers(0, 1000))


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(world=st.integers(2, 6), steps=st.lists(step, min_size=1, max_size=5), cfg=geometry(), one_hop=st.booleans())
def test_collectives_match_reference(world, steps, cfg, one_hop):
    """A random program of 1-5 calls on one world: sequence numbers, rx-pool reuse and parked calls carry over
    from one call to the next.  `one_hop`: the reference-style rings / trees or the B200 backend's one-hop schedules."""
    always_eager = cfg["max_egr_size"] >= (1 << 20)
    if always_eager:
        # messages must fit the rx pool as a whole
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32)

    def fn(a, r, w):
        a.set_timeout(30_000_000)   # eager waits: peers may be seconds late on a loaded machine
        a.set_one_hop_schedules(one_hop)
        for op, count, root, func, salt in steps:
            run_op(a, r, w, op, min(count, 2000) if always_eager else count, root % w, func, salt)
    A.run_ranks(world, fn, cfg, timeout=120.0)


DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.float64, torch.int32, torch.int64]


@settings(max_examples=100, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(op=st.sampled_from(["sendrecv", "bcast", "allgather", "reduce", "allreduce", "reduce_scatter"]),
       world=st.integers(2, 4), count=st.integers(1, 1500), root=st.integers(0, 3), func=st.sampled_from([SUM, MAX]),
       dtype=st.sampled_from(DTYPES), wire=st.sampled_from([None, torch.float16, torch.bfloat16]), cfg=geometry(),
       salt=st.integers(0, 100))
def test_dtypes_and_wire_compression(op, world, count, root, func, dtype, wire, cfg, salt):
    """Every element type, and fp32 over a narrower wire type.  Values are small integers, so every sum is
    exactly representable in every format involved and results must match bit for bit."""
    root %= world
    if dtype != torch.float32:
        wire = None  # compression pairs are registered for fp32 operands
    if cfg["max_egr_size"] >= (1 << 20) or wire is not None:
        # compressed calls are always eager: size the pool for whole messages
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32, max_egr_size=max(cfg["max_egr_size"], 4096),
                   max_rndzv_size=max(cfg["max_rndzv_size"], 1 << 24))
        count = min(count, 1000)

    def vals(n, r):
        g = torch.Generator().manual_seed(131 * salt + 7 * r + 3)
        return torch.randint(-16, 16, (n,), generator=g).to(dtype)

    def red(vs):
        acc = vs[0].to(torch.float64 if dtype.is_floating_point else torch.int64)
        for v in vs[1:]:
            v = v.to(acc.dtype)
            acc = acc + v if func == SUM else torch.maximum(acc, v)
        return acc.to(dtype)

    kw = dict(compress_dtype=wire) if wire is not None else {}

    def fn(a, r, w):
        a.set_timeout(30_000_000)
        if op == "sendrecv":
            s, d = a.create_buffer(count, dtype), a.create_buffer(count, dtype)
            s.host[:] = vals(count, r)
            req = a.send(s, count, (r + 1) % w, tag=1, run_async=True, **kw)
            a.recv(d, count, (r - 1) % w, tag=1, **kw)
            req.wait()
            assert torch.equal(d.host, vals(count, (r - 1) % w))
        elif op == "bcast":
            b = a.create_buffer(count, dtype)
            b.host[:] = vals(count, r)
            a.bcast(b, count, root, **kw)
            assert torch.equal(b.host, vals(count, root))
        elif op == "allgather":
            s, d = a.create_buffer(count, dtype), a.create_buffer(count * w, dtype)
            s.host[:] = vals(count, r)
            a.allgather(s, d, count, **kw)
            assert torch.equal(d.host, torch.cat([vals(count, q) for q in range(w)]))
        elif op == "reduce":
            s, d = a.create_buffer(count, dtype), a.create_buffer(count, dtype)
            s.host[:] = vals(count, r)
            a.reduce(s, d, count, root, func, **kw)
            if r == root:
                assert torch.equal(d.host, red([vals(count, q) for q in range(w)]))
        elif op == "allreduce":
            s, d = a.create_buffer(count, dtype), a.create_buffer(count, dtype)
            s.host[:] = vals(count, r)
            a.allreduce(s, d, count, func, **kw)
            assert torch.equal(d.host, red([vals(count, q) for q in range(w)]))
        else:
            s, d = a.create_buffer(count * w, dtype), a.create_buffer(count, dtype)
            s.host[:] = vals(count * w, r)
            a.reduce_scatter(s, d, count, func, **kw)
            assert torch.equal(d.host, red([vals(count * w, q) for q in range(w)])[r * count:(r + 1) * count])
    A.run_ranks(world, fn, cfg, timeout=120.0)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(world=st.integers(3, 5), data_=st.data(), op=st.sampled_from(["allreduce", "bcast", "allgather", "reduce_scatter"]),
       count=st.integers(1, 1200), func=st.sampled_from([SUM, MAX]), cfg=geometry(), salt=st.integers(0, 100), one_hop=st.booleans())
def test_subcommunicators(world, data_, op, count, func, cfg, salt, one_hop):
    """A random subset of the ranks forms a communicator and runs a collective while the others stay out; then
    everybody meets on the global communicator again (independent sequence spaces, reference test.cpp:701-832)."""
    members = sorted(data_.draw(st.sets(st.integers(0, world - 1), min_size=2, max_size=world)))
    root = data_.draw(st.integers(0, len(members) - 1))
    if cfg["max_egr_size"] >= (1 << 20):
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32)
        count = min(count, 1000)

    def fn(a, r, w):
        a.set_one_hop_schedules(one_hop)
        a.set_timeout(30_000_000)
        if r in members:
            ranks = [a.get_comm_group(0)[m] for m in members]
            me = members.index(r)
            comm = a.create_communicator(ranks, me)
            n = len(members)
            if op == "allreduce":
                s, d = a.create_buffer(count), a.create_buffer(count)
                s.host[:] = data(count, r, salt)
                a.allreduce(s, d, count, func, comm_id=comm)
                vs = [data(count, m, salt) for m in members]
                exp = vs[0].clone()
                for v in vs[1:]:
                    exp = exp + v if func == SUM else torch.maximum(exp, v)
                assert torch.equal(d.host, exp)
            elif op == "bcast":
                b = a.create_buffer(count)
                b.host[:] = data(count, r, salt)
                a.bcast(b, count, root, comm_id=comm)
                assert torch.equal(b.host, data(count, members[root], salt))
            elif op == "allgather":
                s, d = a.create_buffer(count), a.create_buffer(count * n)
                s.host[:] = data(count, r, salt)
                a.allgather(s, d, count, comm_id=comm)
                assert torch.equal(d.host, torch.cat([data(count, m, salt) for m in members]))
            else:
                s, d = a.create_buffer(count * n), a.create_buffer(count)
                s.host[:] = data(count * n, r, salt)
                a.reduce_scatter(s, d, count, func, comm_id=comm)
                vs = [data(count * n, m, salt) for m in members]
                exp = vs[0].clone()
                for v in vs[1:]:
                    exp = exp + v if func == SUM else torch.maximum(exp, v)
                assert torch.equal(d.host, exp[me * count:(me + 1) * count])
        # the global communicator is unaffected
        t = a.create_buffer(8)
        t.host[:] = float(r)
        u = a.create_buffer(8)
        a.allreduce(t, u, 8, SUM)
        assert torch.all(u.host == sum(range(w)))
    A.run_ranks(world, fn, cfg, timeout=120.0)


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(world=st.integers(2, 5), cfg=geometry(), salt=st.integers(0, 1000),
       msgs=st.lists(st.tuples(st.integers(0, 4), st.integers(0, 4), st.integers(1, 2500), st.integers(0, 300), st.booleans()),
                     min_size=1, max_size=12))
def test_point_to_point_programs(world, cfg, salt, msgs):
    """A random list of tagged messages between random pairs, sizes on both sides of the eager threshold.  Every
    rank walks the list in order: asynchronous send if it is the source, blocking receive (exact tag or TAG_ANY) if
    it is the destination — many messages outstanding per pair, eager and rendezvous interleaved."""
    always_eager = cfg["max_egr_size"] >= (1 << 20)
    if always_eager:
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32)
    prog = [(s % world, d % world, min(n, 1000) if always_eager else n, tag, any_) for s, d, n, tag, any_ in msgs
            if s % world != d % world]

    def fn(a, r, w):
        a.set_timeout(30_000_000)
        pending, keep = [], []
        for i, (s, d, n, tag, any_) in enumerate(prog):
            if r == s:
                b = a.create_buffer(n)
                b.host[:] = data(n, s, salt + i)
                keep.append(b)
                pending.append(a.send(b, n, d, tag=tag, run_async=True))
            elif r == d:
                b = a.create_buffer(n)
                a.recv(b, n, s, tag=A.TAG_ANY if any_ else tag)
                assert torch.equal(b.host, data(n, s, salt + i)), (i, s, d, n)
        for q in pending:
            q.wait()
            assert q.retcode() == 0
        a.barrier()
    A.run_ranks(world, fn, cfg, timeout=120.0)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(world=st.integers(2, 4), cfg=geometry(), salt=st.integers(0, 1000),
       calls=st.lists(st.tuples(st.sampled_from(["allreduce", "bcast", "allgather"]), st.integers(1, 1500), st.integers(0, 3)),
                      min_size=2, max_size=6), one_hop=st.booleans())
def test_async_collectives_in_flight(world, cfg, salt, calls, one_hop):
    """Several collectives issued back to back with run_async=True, waited for afterwards: the engine keeps them all
    in flight; collectives on one communicator must still execute in issue order on every rank."""
    always_eager = cfg["max_egr_size"] >= (1 << 20)
    if always_eager:
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32)

    def fn(a, r, w):
        a.set_timeout(30_000_000)
        a.set_one_hop_schedules(one_hop)
        issued = []
        for i, (op, count, root) in enumerate(calls):
            count = min(count, 1000) if always_eager else count
            root %= w
            if op == "allreduce":
                s, d = a.create_buffer(count), a.create_buffer(count)
                s.host[:] = data(count, r, salt + i)
                req = a.allreduce(s, d, count, SUM, run_async=True)
                exp = sum(data(count, q, salt + i) for q in range(w))
            elif op == "bcast":
                s = d = a.create_buffer(count)
                s.host[:] = data(count, r, salt + i)
                req = a.bcast(s, count, root, run_async=True)
                exp = data(count, root, salt + i)
            else:
                s, d = a.create_buffer(count), a.create_buffer(count * w)
                s.host[:] = data(count, r, salt + i)
                req = a.allgather(s, d, count, run_async=True)
                exp = torch.cat([data(count, q, salt + i) for q in range(w)])
            issued.append((req, s, d, exp, i, op))
        for req, s, d, exp, i, op in issued:
            req.wait()
            assert req.retcode() == 0, (i, op)
            d.sync_from_device()
            assert torch.equal(d.host, exp), (i, op)
        a.barrier()
    A.run_ranks(world, fn, cfg, timeout=120.0)


mixed_step = st.one_of(
    st.tuples(st.just("coll"), st.sampled_from(OPS[1:] + ["barrier"]), st.integers(1, 1500), st.integers(0, 5),
              st.sampled_from([SUM, MAX])),
    st.tuples(st.just("msg"), st.integers(0, 4), st.integers(0, 4), st.integers(1, 2500), st.integers(0, 300)))


@settings(max_examples=80, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(world=st.integers(2, 5), cfg=geometry(), salt=st.integers(0, 1000), steps=st.lists(mixed_step, min_size=2, max_size=8),
       one_hop=st.booleans())
def test_mixed_point_to_point_and_collectives(world, cfg, salt, steps, one_hop):
    """Asynchronous (possibly parked) sends interleaved with blocking collectives: a rank sitting in a collective must
    still get its parked rendezvous sends out, or the peer that needs them never joins the collective."""
    always_eager = cfg["max_egr_size"] >= (1 << 20)
    if always_eager:
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32)

    def fn(a, r, w):
        a.set_timeout(30_000_000)
        a.set_one_hop_schedules(one_hop)
        pending, keep = [], []
        for i, st_ in enumerate(steps):
            if st_[0] == "coll":
                _, op, count, root, func = st_
                run_op(a, r, w, op, min(count, 1000) if always_eager else count, root % w, func, salt + i)
            else:
                _, s, d, n, tag = st_
                s, d = s % w, d % w
                n = min(n, 1000) if always_eager else n
                if s == d:
                    continue
                if r == s:
                    b = a.create_buffer(n)
                    b.host[:] = data(n, s, salt + i)
                    keep.append(b)
                    pending.append(a.send(b, n, d, tag=tag, run_async=True))
                elif r == d:
                    b = a.create_buffer(n)
                    a.recv(b, n, s, tag=tag)
                    assert torch.equal(b.host, data(n, s, salt + i)), i
        for q in pending:
            q.wait()
            assert q.retcode() == 0
        a.barrier()
    A.run_ranks(world, fn, cfg, timeout=120.0)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(cfg=geometry(), salt=st.integers(0, 1000),
       ops=st.lists(st.tuples(st.sampled_from(["push", "pop", "through"]), st.integers(1, 2000)), min_size=1, max_size=10))
def test_stream_port_is_a_fifo(cfg, salt, ops):
    """The kernel stream port with loop-back behaves like one byte FIFO, whatever the chunking: copy_to_stream pushes,
    copy_from_stream pops (possibly across push boundaries), copy_from_to_stream moves data from the head to the tail."""
    from accl_b200 import DataType

    def fn(a, r, w):
        model = []          # expected FIFO content (floats)
        k = 0
        for op, n in ops:
            if op == "push":
                b = a.create_buffer(n)
                b.host[:] = data(n, 0, salt + k)
                k += 1
                a.copy_to_stream(b, n)
                model += b.host.tolist()
            elif op == "pop":
                n = min(n, len(model))
                if n == 0:
                    continue
                b = a.create_buffer(n)
                a.copy_from_stream(b, n)
                assert b.host.tolist() == model[:n]
                del model[:n]
            else:
                n = min(n, len(model))
                if n == 0:
                    continue
                a.copy_from_to_stream(DataType.float32, n)
                model += model[:n]
                del model[:n]
        if model:
            b = a.create_buffer(len(model))
            a.copy_from_stream(b, len(model))
            assert b.host.tolist() == model
    A.run_ranks(1, fn, cfg, timeout=60.0)


KINDS = ["device", "host_only", "p2p"]


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(op=st.sampled_from(["sendrecv", "allreduce", "bcast", "allgather", "reduce_scatter"]), world=st.integers(2, 4),
       count=st.integers(1, 1500), ks=st.sampled_from(KINDS), kd=st.sampled_from(KINDS), cfg=geometry(), salt=st.integers(0, 100))
def test_buffer_kinds(op, world, count, ks, kd, cfg, salt):
    """Operands in device memory, host-only memory (the engine reaches across: OP*_HOST / RES_HOST flags) or
    peer-visible memory, in every combination and on both protocols."""
    if cfg["max_egr_size"] >= (1 << 20):
        cfg = dict(cfg, egr_rx_buf_size=4096, n_egr_rx_bufs=32)
        count = min(count, 1000)

    def fn(a, r, w):
        a.set_timeout(30_000_000)

        def mk(n, kind):
            return a.create_buffer(n, kind=getattr(A.BufferKind, kind))
        if op == "sendrecv":
            s, d = mk(count, ks), mk(count, kd)
            s.host[:] = data(count, r, salt)
            req = a.send(s, count, (r + 1) % w, tag=1, run_async=True)
            a.recv(d, count, (r - 1) % w, tag=1)
            req.wait()
            assert torch.equal(d.host, data(count, (r - 1) % w, salt))
        elif op == "allreduce":
            s, d = mk(count, ks), mk(count, kd)
            s.host[:] = data(count, r, salt)
            a.allreduce(s, d, count, SUM)
            assert torch.equal(d.host, sum(data(count, q, salt) for q in range(w)))
        elif op == "bcast":
            b = mk(count, ks)
            b.host[:] = data(count, r, salt)
            a.bcast(b, count, 1 % w)
            assert torch.equal(b.host, data(count, 1 % w, salt))
        elif op == "allgather":
            s, d = mk(count, ks), mk(count * w, kd)
            s.host[:] = data(count, r, salt)
            a.allgather(s, d, count)
            assert torch.equal(d.host, torch.cat([data(count, q, salt) for q in range(w)]))
        else:
            s, d = mk(count * w, ks), mk(count, kd)
            s.host[:] = data(count * w, r, salt)
            a.reduce_scatter(s, d, count, SUM)
            assert torch.equal(d.host, sum(data(count * w, q, salt) for q in range(w))[r * count:(r + 1) * count])
    A.run_ranks(world, fn, cfg, timeout=120.0)
