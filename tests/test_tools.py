"""The reporting tools must keep working on the committed measurements (profiles/): sweep report with NCCL ratios
and cost-model column, ncu summary of the collective kernels, cost model sanity.  Counterpart of the reference's
test/host/xrt/parse_bench_results.py being exercised by its CI."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def run(*args):
    r = subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    return r.stdout


@pytest.mark.parametrize("csv", ["sweep_8gpu_direct.csv", "sweep_2gpu_direct.csv", "sweep_2gpu_engine.csv", "sweep_4gpu_wire_bf16.csv"])
def test_sweep_report_renders_every_row(csv):
    path = os.path.join(PROF, csv)
    out = run("bench/report.py", path)
    rows = [ln for ln in out.splitlines() if ln.startswith(("| all", "| reduce_scatter"))]
    n = sum(1 for _ in open(path)) - 1
    assert len(rows) == n
    for ln in rows:
        cells = [c.strip() for c in ln.strip("|").split("|")]
        assert cells[7].endswith("x") and float(cells[7][:-1]) > 0          # speedup vs NCCL
        assert cells[-1].endswith("%") and 0 < float(cells[-1][:-1]) <= 150  # % of the alpha-beta ideal at 900 GB/s/dir


def test_ncu_summary_joins_plan_and_capture():
    out = run("bench/ncu_summary.py", os.path.join(PROF, "ncu_collectives_2gpu.csv"), os.path.join(PROF, "ncu_collectives_2gpu_plan.json"))
    rows = [ln for ln in out.splitlines() if ln.startswith("| all") or ln.startswith("| reduce") or ln.startswith("| bcast")]
    assert len(rows) == 14                                                   # the 14 calls of bench/ncu_target.py, second repetition
    big = [ln for ln in rows if "256 MiB" in ln and ln.startswith("| allreduce |")][0]
    cells = [c.strip() for c in big.strip("|").split("|")]
    assert int(cells[2]) == 128 and 300 < float(cells[3]) < 600              # grid, kernel us
    assert 250 < float(cells[5]) < 300                                       # DRAM read MB: the message once


def test_cost_model_is_anchored_at_the_nominal_link_rate():
    from accl_b200.models.cost_model import ideal_us
    # large messages: the bandwidth term at the NOMINAL 900 GB/s per direction.  1 GiB all-reduce on 8 GPUs through the
    # switch moves M (1 + 1/P) per direction, between peers (2 GPUs) 2 M (P-1)/P
    t = ideal_us("allreduce", 1 << 30, 8)
    assert abs(t - (1.125 * (1 << 30) / 900e9 * 1e6)) / t < 0.01
    t2 = ideal_us("allreduce", 1 << 30, 2)
    assert abs(t2 - ((1 << 30) / 900e9 * 1e6)) / t2 < 0.01
    assert ideal_us("allgather", 1 << 20, 2) < ideal_us("allgather", 1 << 20, 8) * 8
    assert ideal_us("allreduce", 1024, 8) < 10                                # latency floor: one hop + a launch


def test_bench_reference_arm_reports_unavailable_and_exits_zero():
    """Driver contract: `bench.py --impl reference` prints one JSON line {"impl": "reference", "unavailable": ...} and exits 0
    (the reference is an FPGA design and cannot be installed here)."""
    import json
    out = run("bench.py", "--impl", "reference", "--gpus", "8", "--steps", "5", "--warmup", "3")
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    row = json.loads(lines[0])
    assert row["impl"] == "reference" and isinstance(row["unavailable"], str) and len(row["unavailable"]) > 20


def test_graft_entry_exposes_build_and_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    assert callable(g.build) and callable(g.smoke)
