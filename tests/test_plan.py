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


def test_eager_vs_rendezvous_threshold():
    assert plan(Op.allreduce, 64 << 10)["algo"] == "eager"            # bytes <= max_eager_size
    assert plan(Op.allreduce, (64 << 10) + 4)["algo"] != "eager"
    assert plan(Op.allreduce, 64 << 20, compressed=True)["algo"] == "eager"   # compressed calls always use slots
    assert plan(Op.send, 1 << 20)["algo"] == "p2p" and plan(Op.recv, 1 << 20)["algo"] == "p2p"
    assert plan(Op.send, 1024)["n_ctas"] == 1                         # eager point to point: one channel


def test_allreduce_algorithm_by_size_and_world():
    # one-shot while bytes * P <= 2 MiB, then two-shot: through the switch from 3 ranks, peer load/store below
    assert plan(Op.allreduce, 128 << 10, world=8)["algo"] == "p2p_oneshot"
    assert plan(Op.allreduce, 512 << 10, world=8)["algo"] == "nvls"
    assert plan(Op.allreduce, 512 << 10, world=2)["algo"] == "p2p_oneshot"
    assert plan(Op.allreduce, 4 << 20, world=2)["algo"] == "p2p"
    assert plan(Op.allreduce, 4 << 20, world=8, has_mc=False)["algo"] == "p2p"
    assert plan(Op.allreduce, 4 << 20, world=4, nvls_min_ranks=99)["algo"] == "p2p"


def test_nvls_only_for_ops_that_win_through_the_switch():
    for op in (Op.allreduce, Op.bcast, Op.reduce):
        assert plan(op, 16 << 20)["algo"] == "nvls"
    for op in (Op.allgather, Op.reduce_scatter, Op.scatter, Op.gather, Op.alltoall):
        assert plan(op, 1 << 20)["algo"] == "p2p"                     # less traffic per link with peer stores / loads


def test_channel_counts():
    assert plan(Op.allreduce, 256 << 20, max_ctas=128)["n_ctas"] == 64      # measured: 64 beats 128 through the switch
    assert plan(Op.allreduce, 64 << 20, world=2, max_ctas=128)["n_ctas"] == 64
    assert plan(Op.allreduce, 256 << 20, world=2, max_ctas=128)["n_ctas"] == 128
    assert plan(Op.reduce_scatter, 32 << 20, max_ctas=128)["n_ctas"] == 128  # 8 x 32 MiB moved: all channels
    assert plan(Op.allgather, 128 << 10, max_ctas=128)["n_ctas"] == 8        # 8 x 128 KiB moved at 128 KiB per channel
    assert plan(Op.allreduce, 1024, max_ctas=128)["n_ctas"] == 1
    assert plan(Op.allreduce, 64 << 10, max_ctas=128)["n_ctas"] == 4         # eager: 16 KiB per channel, <= 16 channels
    assert plan(Op.allreduce, 256 << 20, max_ctas=1000)["n_ctas"] <= 160     # never more channels than sync pads
    assert plan(Op.copy, 1 << 30)["algo"] == "local" and plan(Op.copy, 1 << 30)["n_ctas"] == 296
    assert plan(Op.barrier, 0)["n_ctas"] == 1 and plan(Op.nop, 0)["algo"] == "local"
