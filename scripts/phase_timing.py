"""Where does a small call's time go?  Run with the ACCL_PHASE_TIMING build:

  python -m accl_b200.utils.build --variant timing --define ACCL_PHASE_TIMING      # on the build host
  ACCL_VARIANT=timing python -m torch.distributed.run --nproc-per-node N ... scripts/phase_timing.py

For every (op, size) it issues K asynchronous calls, device-times them with CUDA events and prints, next to
the per-call time, the engine-measured kernel duration (get_duration) and the control block's phase counters
(time inside flag meetings, number of meetings) averaged per call."""
import os
import re
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accl_b200 as A  # noqa: E402


def counters(acc):
    m = re.search(r"calls=(\d+) kernel_ns=(\d+) sync_ns=(\d+) syncs=(\d+) wait_ns=(\d+) waits=(\d+)", acc.cuda_debug_state())
    return tuple(int(x) for x in m.groups()) if m else None


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    acc = A.cuda_rank(rank, world, local, heap_mb=512, max_ctas=64)
    acc.initialize(n_egr_rx_bufs=4, egr_rx_buf_size=64 << 10, max_egr_size=64 << 10, max_rndzv_size=1 << 30)
    nmax = (4 << 20) // 4
    s, d = acc.create_buffer(nmax * world), acc.create_buffer(nmax * world)
    s.dev.fill_(1.0)
    K = 200
    if counters(acc) is None and rank == 0:
        print("# built without ACCL_PHASE_TIMING: only event and get_duration columns are meaningful")
    for op in ("allreduce", "allgather", "reduce_scatter", "bcast"):
        for nbytes in (1 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20):
            n = nbytes // 4
            per = max(n // world, 1)
            kw = dict(from_fpga=True, to_fpga=True, run_async=True)
            call = {"allreduce": lambda: acc.allreduce(s, d, n, A.SUM, **kw),
                    "allgather": lambda: acc.allgather(s, d, per, **kw),
                    "reduce_scatter": lambda: acc.reduce_scatter(s, d, per, A.SUM, **kw),
                    "bcast": lambda: acc.bcast(s, n, 0, **kw)}[op]
            for _ in range(20):
                call().free()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            c0 = counters(acc)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reqs = [call() for _ in range(K)]
            e1.record()
            torch.cuda.synchronize()
            dur = sum(r.duration_ns() for r in reqs) / K
            for r in reqs:
                r.free()
            c1 = counters(acc)
            line = f"{op:15s} {nbytes:9d} B  events {e0.elapsed_time(e1) / K * 1e3:7.1f} us/call  kernel(get_duration) {dur / 1e3:7.1f} us"
            if c0 and c1 and c1[0] > c0[0]:
                calls = c1[0] - c0[0]
                line += (f"  body {(c1[1] - c0[1]) / calls / 1e3:7.1f} us  in meetings {(c1[2] - c0[2]) / calls / 1e3:7.1f} us"
                         f" ({(c1[3] - c0[3]) / calls:.1f} per call)"
                         f"  flag waits {(c1[4] - c0[4]) / max(c1[5] - c0[5], 1) / 1e3:6.2f} us avg x {(c1[5] - c0[5]) / calls:.1f}")
            if rank == 0:
                print(line, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
