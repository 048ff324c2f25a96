"""Random call programs on the CUDA backend (same generator as tests/test_emulator_property.py), for GPU
sessions: `python scripts/gpu_fuzz.py [examples] [max_world]`.  Not part of the pytest suite: a failure here is
a finding to triage, not a regression gate.  Ranks are placed round-robin on the visible GPUs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

import accl_b200 as A  # noqa: E402
import test_emulator_property as t  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
MAXW = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NGPU = max(torch.cuda.device_count(), 1)


@st.composite
def gpu_geometry(draw):
    slot = draw(st.sampled_from([1 << 10, 4 << 10, 16 << 10, 64 << 10]))
    depth = draw(st.sampled_from([2, 4, 8]))
    max_egr = draw(st.sampled_from([slot, 4 * slot]))
    return dict(n_egr_rx_bufs=depth, egr_rx_buf_size=slot, max_egr_size=max_egr, max_rndzv_size=1 << 30)


@settings(max_examples=N, deadline=None, suppress_health_check=list(HealthCheck), database=None, print_blob=True)
@given(world=st.integers(2, MAXW), steps=st.lists(t.step, min_size=1, max_size=5), cfg=gpu_geometry(),
       max_ctas=st.sampled_from([1, 4, 16]))
def program(world, steps, cfg, max_ctas):
    def fn(a, r, w):
        for op, count, root, func, salt in steps:
            t.run_op(a, r, w, op, count * 8, root % w, func, salt)   # x8: reach past the eager thresholds
    A.run_cuda_ranks([r % NGPU for r in range(world)], fn, cfg, heap_mb=128, max_ctas=max_ctas, timeout=120.0)


if __name__ == "__main__":
    t0 = time.time()
    program()
    print(f"{N} programs ok in {time.time() - t0:.0f} s")
