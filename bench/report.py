"""Turn sweep CSVs (bench/sweep.py) into a markdown report with NCCL ratios and the
fraction of the alpha-beta ideal (accl_b200/models/cost_model.py).  Counterpart of the
reference's test/host/xrt/parse_bench_results.py.

  python bench/report.py profiles/sweep_8gpu_nvls.csv > profiles/sweep_8gpu.md
"""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from accl_b200.models.cost_model import ideal_us  # noqa: E402


def human(n):
    for u in ("B", "KiB", "MiB", "GiB"):
        if n < 1024:
            return f"{n:g} {u}"
        n /= 1024
    return f"{n:g} TiB"


def main():
    for path in sys.argv[1:]:
        rows = list(csv.DictReader(open(path)))
        if not rows:
            continue
        world = int(rows[0]["world"])
        print(f"### {os.path.basename(path)} — {world} x B200, {rows[0]['dtype']}\n")
        mode = rows[0].get("mode", "direct")
        print(f"(mode: {mode}; `us` = CUDA-event time per call of a host loop of asynchronous calls, median of batches, max over "
              f"ranks; `device us` = engine-measured duration of the call itself; `graph us` = per call inside a replayed CUDA graph)\n")
        print("| op | size | accl us | device us | accl busbw GB/s | NCCL us | NCCL busbw GB/s | speedup vs NCCL | accl graph us | NCCL graph us | "
              "speedup (graphs) | % of ideal (900 GB/s/dir) |")
        print("|---|---|---|---|---|---|---|---|---|---|---|---|")

        def f(r, k, fmt="{:.1f}"):
            v = r.get(k)
            return fmt.format(float(v)) if v not in (None, "", "None") else "-"
        for r in rows:
            nb = int(r["bytes"])
            a_us = float(r["accl_us"])
            n_us = float(r["nccl_us"]) if r.get("nccl_us") else 0.0
            ideal = ideal_us(r["op"], nb, world)
            sg = f(r, "speedup_graph", "{:.2f}x")
            print(f"| {r['op']} | {human(nb)} | {a_us:.1f} | {f(r, 'accl_device_us')} | {float(r['accl_busbw']):.1f} | "
                  f"{n_us:.1f} | {float(r['nccl_busbw']) if r.get('nccl_busbw') else 0:.1f} | "
                  f"{(n_us / a_us if n_us else 0):.2f}x | {f(r, 'accl_graph_us')} | {f(r, 'nccl_graph_us')} | {sg} | {100 * ideal / a_us:.0f}% |")
        print()


if __name__ == "__main__":
    main()
