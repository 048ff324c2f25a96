"""Alpha-beta cost model of the collectives on an NVSwitch domain, used to grade sweep
results (the reference grades its FPGA sweeps against an ideal model the same way:
test/host/xrt/parse_bench_results.py:50-69 — send M/BW; bcast (P-1)M/BW; allreduce
2(P-1)(rtt/2 + (M/P)/BW) ... with BW = 100 Gb/s, rtt = 500 ns).

On NVLink 5 through NVSwitch every peer is one hop, so the ideal times are those of the
one-hop algorithms this library uses, per direction per GPU:

  allreduce, NVLS two-shot   bytes per link direction  M (1 + 1/P)        (ld_reduce up + st down)
  allreduce, peer two-shot                              2 M (P-1)/P
  allgather / reduce_scatter (peer push / pull)         M (P-1)/P          (M = total bytes)
  bcast via multimem.st                                 M   (root sends once, the switch replicates)
  reduce via multimem.ld_reduce at the root             M
  send / recv                                           M
"""
from dataclasses import dataclass


@dataclass
class Fabric:
    bw_GBps: float = 900.0       # NVLink 5, per direction per GPU (nominal)
    bw_measured_GBps: float = 770.0  # peer copy measured on this pool (B200_PROFILING.md)
    alpha_us: float = 1.5        # one-way flag / data latency between GPUs over the switch (LL exchange: ~1.2-1.8 us)
    launch_us: float = 4.0       # kernel launch + completion on the issuing stream


def ideal_us(op: str, nbytes: int, world: int, fabric: Fabric = Fabric(), nvls: bool = True, measured: bool = False) -> float:
    """Ideal device time of one call in microseconds.  Graded against the nominal 900 GB/s per direction per GPU that
    BASELINE.json names (measured=True switches to the 770 GB/s peer-copy rate measured on this pool)."""
    if world == 1:
        return nbytes / 6.5e6  # local copy at HBM rate (read+write), bytes / (6.5 TB/s / 2 ... ) kept simple
    bw = (fabric.bw_measured_GBps if measured else fabric.bw_GBps) * 1e3  # bytes per microsecond
    p = world
    if op == "allreduce":
        link_bytes = nbytes * (1 + 1 / p) if nvls and p >= 3 else 2 * nbytes * (p - 1) / p
        syncs = 2  # two hops (reduce-scatter + all-gather)
    elif op in ("allgather", "reduce_scatter", "alltoall", "scatter", "gather"):
        link_bytes = nbytes * (p - 1) / p
        syncs = 1  # one hop
    elif op in ("bcast", "reduce", "sendrecv"):
        link_bytes = nbytes
        syncs = 1
    elif op == "barrier":
        link_bytes, syncs = 0, 1
    else:
        raise ValueError(op)
    return syncs * fabric.alpha_us + link_bytes / bw


def efficiency(op: str, nbytes: int, world: int, measured_us: float, **kw) -> float:
    """Fraction of the ideal achieved (1.0 = at the model's roofline)."""
    return ideal_us(op, nbytes, world, **kw) / measured_us
