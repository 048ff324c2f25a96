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
    print(f"# ncu: collective kernels on {plan['world']} x B200 (one rank under ncu, single-pass metric set, `--clock-control none`)\n")
    print(f"`{plan['describe']}`\n")
    have_nvl = any("nvltx__bytes.sum" in m for pid in pids for m in per[pid].values())
    print("One `k_call` launch per call; rows are the second repetition of every call (the first warms up); the profiled rank runs "
          "under ncu with a metric set that fits one pass (nothing is replayed), its peers run unprofiled.  "
          "Rates = bytes / kernel time." + ("" if have_nvl else "  NVLink byte counts for the same calls: `profiles/nvlink_traffic_*.jsonl` "
                                            "(`nvidia-smi nvlink -gt d` deltas; ncu needs a second pass for the nvl* counters, which a "
                                            "multi-rank kernel cannot be replayed for).") + "\n")
    hdr = ["call", "size", "grid", "kernel us", "algorithm bytes / kernel time GB/s"]
    if have_nvl:
        hdr += ["NVLink tx MB", "NVLink rx MB", "tx GB/s", "rx GB/s"]
    hdr += ["DRAM read MB", "DRAM write MB", "DRAM GB/s"]
    print("| " + " | ".join(hdr) + " |")
    print("|" + "---|" * len(hdr))
    P = plan["world"]
    for pi, pid in enumerate(pids):
        ks = [per[pid][k] for k in sorted(per[pid])]
        for i, c in enumerate(calls):
            if i >= len(ks) or c["rep"] != 1:
                continue
            m = ks[i]

            def b(name):
                return to_bytes(*m[name]) if name in m else float("nan")
            t = to_us(*m["gpu__time_duration.sum"])
            # bytes this rank has to send (= receive) for the call, the quantity the bus-bandwidth convention is built on
            nb = c["bytes"]
            wire = {"allreduce": 2.0 * nb * (P - 1) / P, "allreduce bf16 wire": nb * (P - 1) / P, "allgather": nb * (P - 1) / P,
                    "reduce_scatter": nb * (P - 1) / P, "bcast": nb, "reduce": nb}.get(c["op"], nb)
            row = [c["op"], human(nb), m.get("launch__grid_size", ("?", ""))[0], f"{t:.1f}", f"{wire / t * 1e-3:.0f}"]
            if have_nvl:
                tx, rx = b("nvltx__bytes.sum"), b("nvlrx__bytes.sum")
                row += [f"{tx / 1e6:.2f}", f"{rx / 1e6:.2f}", f"{tx / t * 1e-3:.0f}", f"{rx / t * 1e-3:.0f}"]
            rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
            row += [f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}", f"{(rd + wr) / t * 1e-3:.0f}"]
            print("| " + " | ".join(row) + " |")
    print()


if __name__ == "__main__":
    main()
