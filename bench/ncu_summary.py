"""Join an `ncu --csv` log of collective kernels (bench/ncu_target.py under application replay) with the list of
calls the target wrote, and print a markdown table: per call and rank, kernel time, grid, NVLink bytes sent /
received, DRAM bytes, and the rates they imply.

  python bench/ncu_summary.py gpurun_out/ncu_coll_2gpu.csv gpurun_out/ncu_coll_2gpu_plan.json > profiles/ncu_collectives_2gpu.md
"""
import csv
import json
import sys
from collections import defaultdict


def human(n):
    for u in ("B", "KiB", "MiB", "GiB"):
        if n < 1024:
            return f"{n:g} {u}"
        n /= 1024
    return f"{n:g} TiB"


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    u = unit.lower()
    for name, mult in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if name in u:
            return v * mult
    return v


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.strip(), 1.0)


def main():
    log, plan_path = sys.argv[1], sys.argv[2]
    plan = json.load(open(plan_path))
    lines = [ln for ln in open(log, errors="replace") if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = defaultdict(lambda: defaultdict(dict))  # pid -> kernel id -> metric -> (value, unit)
    for r in rows:
        per[r["Process ID"]][int(r["ID"])][r["Metric Name"]] = (r["Metric Value"], r["Metric Unit"])
    pids = sorted(per)
    calls = plan["calls"]
    print(f"# ncu: collective kernels on {plan['world']} x B200 (application replay, `--clock-control none`)\n")
    print(f"`{plan['describe']}`\n")
    print("One `k_call` launch per call; rows are the second repetition of every call (the first warms up).  NVLink columns are this GPU's "
          "`nvltx__bytes.sum` / `nvlrx__bytes.sum` (all 18 links, headers included; `user` = payload only); rates = bytes / kernel time.\n")
    print("| call | size | rank (pid) | grid | kernel us | NVLink tx MB (user) | NVLink rx MB (user) | tx GB/s | rx GB/s | DRAM rd MB | DRAM wr MB | SM busy % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for pi, pid in enumerate(pids):
        ks = [per[pid][k] for k in sorted(per[pid])]
        # the trailing barrier (and anything else after the planned calls) is ignored
        for i, c in enumerate(calls):
            if i >= len(ks) or c["rep"] != 1:
                continue
            m = ks[i]

            def b(name):
                return to_bytes(*m[name]) if name in m else float("nan")
            t = to_us(*m["gpu__time_duration.sum"])
            tx, rx, txu, rxu = b("nvltx__bytes.sum"), b("nvlrx__bytes.sum"), b("nvltx__bytes_data_user.sum"), b("nvlrx__bytes_data_user.sum")
            print(f"| {c['op']} | {human(c['bytes'])} | {pi} ({pid}) | {m.get('launch__grid_size', ('?', ''))[0]} | {t:.1f} | {tx / 1e6:.2f} ({txu / 1e6:.2f}) | "
                  f"{rx / 1e6:.2f} ({rxu / 1e6:.2f}) | {tx / t * 1e-3:.0f} | {rx / t * 1e-3:.0f} | {b('dram__bytes_read.sum') / 1e6:.1f} | "
                  f"{b('dram__bytes_write.sum') / 1e6:.1f} | {float(m.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', ('nan', ''))[0]):.1f} |")
    print()


if __name__ == "__main__":
    main()
