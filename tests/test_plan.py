"""The CUDA backend's call planner (csrc/include/accl/cuda/plan.hpp) as a pure function: protocol, algorithm
and channel-count decisions must be identical on every rank and follow the documented rules (docs/tuning.md).
Runs without a GPU (the planner is host/device-shared code); skipped on emulator-only builds."""
import pytest

import accl_b200 as A
from accl_b200 import DataType

pytestmark = pytest.mark.skipif(not A._C.with_cuda, reason="built without the CUDA backend")
Op = A._C.Operation
F32 = DataType.float32


def plan(op, nbytes, world=8, dtype=F32, esz=4, **kw):
    return A._C.cuda_plan(op, nbytes // esz, dtype, world, **kw)


ONE_WAY = ("eager", "ll", "staged")  # protocols without a meeting: slot ring, flag-in-data, payload + flag


def test_eager_vs_rendezvous_threshold():
    assert plan(Op.allreduce, 64 << 10)["algo"] in ONE_WAY           # bytes <= max_eager_size
    assert plan(Op.allreduce, (64 << 10) + 4)["algo"] not in ONE_WAY
    assert plan(Op.allreduce, 64 << 20, compressed=True)["algo"] == "eager"   # compressed calls always use slots
    assert plan(Op.send, 1 << 20)["algo"] == "p2p" and plan(Op.recv, 1 << 20)["algo"] == "p2p"
    assert plan(Op.send, 1024)["algo"] == "eager" and plan(Op.send, 1024)["n_ctas"] == 1  # eager point to point: one channel


def test_one_way_protocol_selection():
    # all-reduce: one hop (everybody sends everything) while tiny, reduce-scatter + all-gather above
    p = plan(Op.allreduce, 1024)
    assert p["algo"] == "ll" and p["oneshot"] and p["n_ctas"] == 1
    p = plan(Op.allreduce, 64 << 10)
    assert p["algo"] == "ll" and not p["oneshot"]                     # shards of 8 KiB through the flag-in-data protocol
    p = plan(Op.allreduce, 2 << 20, max_eager_bytes=4 << 20, ll_kb=2048)
    assert p["algo"] == "ll" and not p["oneshot"] and p["n_ctas"] == 32   # shards of 256 KiB: 32 K lines over all channels
    # what does not fit the LL region: the payload + flag protocol when that is enabled, else the slot ring while
    # that is a handful of segments, else the rendezvous algorithms
    assert plan(Op.allreduce, 2 << 20, max_eager_bytes=4 << 20)["algo"] == "nvls"
    assert plan(Op.allreduce, 2 << 20, max_eager_bytes=4 << 20, staged_max_bytes=1 << 20)["algo"] == "staged"
    assert plan(Op.allgather, 128 << 10, max_eager_bytes=4 << 20, ll_kb=128)["algo"] == "eager"   # 2 segments of 64 KiB
    assert plan(Op.allgather, 256 << 10, max_eager_bytes=4 << 20)["algo"] == "p2p"     # 4 segments: user buffer to user buffer
    # sizes that do not split into 16-byte shards go one-shot while small, to the other paths otherwise
    assert plan(Op.allreduce, 20000)["oneshot"]
    assert plan(Op.allreduce, (1 << 20) + 4, max_eager_bytes=4 << 20)["algo"] == "nvls"
    # per-peer message decides for the others
    assert plan(Op.allgather, 8 << 10)["algo"] == "ll"
    assert plan(Op.reduce_scatter, 16 << 10)["algo"] == "ll"
    assert plan(Op.bcast, 32 << 10)["algo"] == "ll"
    assert plan(Op.bcast, 32 << 10, ll_max_bytes=16384, staged_max_bytes=1 << 20)["algo"] == "staged"   # the crossover is a knob
    assert plan(Op.allgather, 256 << 10, max_eager_bytes=1 << 20, staged_max_bytes=1 << 20)["algo"] == "staged"   # 256 KiB > LL capacity of 128 KiB
    assert plan(Op.allgather, 256 << 10, max_eager_bytes=1 << 20, ll_kb=2048)["algo"] == "ll"
    # flag-in-data doubles the wire bytes: budgeted on message x (P - 1) peers (2 MiB), so the crossover moves with P
    big = dict(max_eager_bytes=4 << 20, ll_kb=2048)
    assert plan(Op.allgather, 256 << 10, world=8, **big)["algo"] == "ll"      # 7 x 256 KiB
    assert plan(Op.allgather, 512 << 10, world=8, **big)["algo"] == "p2p"     # 7 x 512 KiB: user buffer to user buffer
    assert plan(Op.allgather, 512 << 10, world=4, **big)["algo"] == "ll"      # 3 x 512 KiB
    assert plan(Op.allreduce, 4 << 20, world=8, **big)["algo"] == "nvls"      # shards of 512 KiB to 7 peers
    assert plan(Op.allreduce, 2 << 20, world=2, **big)["algo"] == "ll"        # one shard of 1 MiB to one peer
    assert plan(Op.bcast, 512 << 10, world=8, **big)["algo"] == "nvls"        # the root would push 7 x 512 KiB x 2
    # no staging configured: everything one-way is the slot ring
    assert plan(Op.allreduce, 1024, stage_kb=0, ll_kb=0)["algo"] == "eager"
    # channel counts: every channel owns 1/32 of a region
    assert plan(Op.allgather, 256 << 10, max_eager_bytes=1 << 20, staged_max_bytes=1 << 20)["n_ctas"] == 8   # 32 KiB per channel
    # inside the engine a message that fits one channel stays on one (executed inline by the control CTA)
    assert plan(Op.allreduce, 8 << 10, ll_kb=2048, engine_mode=True)["n_ctas"] == 1
    assert plan(Op.allreduce, 8 << 10, ll_kb=2048)["n_ctas"] == 2


def test_allreduce_algorithm_by_size_and_world():
    # one-shot while bytes * P <= 2 MiB, then two-shot: through the switch from 3 ranks, peer load/store below
    assert plan(Op.allreduce, 128 << 10, world=8)["algo"] == "p2p_oneshot"
    assert plan(Op.allreduce, 512 << 10, world=8)["algo"] == "nvls"
    assert plan(Op.allreduce, 512 << 10, world=2)["algo"] == "p2p_oneshot"
    assert plan(Op.allreduce, 128 << 10, world=8, max_eager_bytes=1 << 20)["algo"] in ONE_WAY
    assert plan(Op.allreduce, 4 << 20, world=2)["algo"] == "p2p"
    assert plan(Op.allreduce, 4 << 20, world=8, has_mc=False)["algo"] == "p2p"
    assert plan(Op.allreduce, 4 << 20, world=4, nvls_min_ranks=99)["algo"] == "p2p"


def test_nvls_only_for_ops_that_win_through_the_switch():
    for op in (Op.allreduce, Op.bcast, Op.reduce):
        assert plan(op, 16 << 20)["algo"] == "nvls"
    for op in (Op.allgather, Op.reduce_scatter, Op.scatter, Op.gather, Op.alltoall):
        assert plan(op, 1 << 20)["algo"] == "p2p"                     # less traffic per link with peer stores / loads


def test_channel_counts():
    assert plan(Op.allreduce, 256 << 20, max_ctas=128)["n_ctas"] == 32      # measured: few channels win through the switch
    assert plan(Op.allreduce, 64 << 20, world=2, max_ctas=128)["n_ctas"] == 128    # peer loads / stores want them all
    assert plan(Op.allreduce, 256 << 20, world=2, max_ctas=128)["n_ctas"] == 128
    assert plan(Op.reduce_scatter, 32 << 20, max_ctas=128)["n_ctas"] == 128  # 8 x 32 MiB moved: all channels
    assert plan(Op.allgather, 128 << 10, max_ctas=128)["n_ctas"] == 8        # 8 x 128 KiB moved at 128 KiB per channel
    assert plan(Op.allreduce, 1024, max_ctas=128)["n_ctas"] == 1
    assert plan(Op.allreduce, 64 << 10, max_ctas=128, stage_kb=0, ll_kb=0)["n_ctas"] == 4   # slot ring: 16 KiB per channel, <= 16 channels
    assert plan(Op.allreduce, 256 << 20, max_ctas=1000, nvls_min_ranks=99)["n_ctas"] <= 128   # never more channels than sync pads
    assert plan(Op.copy, 1 << 30)["algo"] == "local" and plan(Op.copy, 1 << 30)["n_ctas"] == 296
    assert plan(Op.barrier, 0)["n_ctas"] == 1 and plan(Op.nop, 0)["algo"] == "local"
