"""Emulator launcher: N rank processes on this machine, one CPU engine each,
talking over loopback TCP (SocketFabric).

  python -m accl_b200.models.emulator -n 4 --selftest          # spawn + run the built-in check
  python -m accl_b200.models.emulator -n 4 -- python my_app.py # spawn an app once per rank

  python -m accl_b200.models.emulator -n 4 --engines           # only the engine processes (cclo_emu); drivers
                                                               # attach with accl_b200.remote_rank()

Every child gets RANK / WORLD_SIZE / ACCL_EMU_PORT; inside, `accl_b200.socket_rank()`
gives the rank's Accl.  Counterpart of the reference's test/model/emulator/run.py
(spawns N cclo_emu processes) + utility.cpp `--startemu`; signal handling tears
the whole group down.
"""
import argparse
import os
import signal
import socket
import subprocess
import sys


def free_port_block(n, also_at=()):
    """A base port with n consecutive free ports (best effort); `also_at`: extra offsets whose n-port blocks must be
    free as well (the engines' control ports live at base + 1000)."""
    for _ in range(100):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        base = s.getsockname()[1]
        s.close()
        if base + max((0,) + tuple(also_at)) + n >= 65000:
            continue
        ok = True
        socks = []
        try:
            for off in (0,) + tuple(also_at):
                for i in range(n):
                    t = socket.socket()
                    t.bind(("127.0.0.1", base + off + i))
                    socks.append(t)
        except OSError:
            ok = False
        for t in socks:
            t.close()
        if ok:
            return base
    raise RuntimeError("no free port block found")


def launch(world, argv, base_port=None, env_extra=None, timeout=None):
    base_port = base_port or free_port_block(world)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), ACCL_EMU_PORT=str(base_port),
                   MASTER_ADDR="127.0.0.1")
        env.update(env_extra or {})
        procs.append(subprocess.Popen(argv, env=env))

    def kill_all(*_):
        for p in procs:
            if p.poll() is None:
                p.terminate()

    old = signal.signal(signal.SIGINT, kill_all)
    try:
        rcs = []
        for p in procs:
            try:
                rcs.append(p.wait(timeout=timeout))
            except subprocess.TimeoutExpired:
                kill_all()
                rcs.append(-9)
        return rcs
    finally:
        signal.signal(signal.SIGINT, old)
        kill_all()


def engine_binary():
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    exe = os.path.join(root, "build", "bin", "cclo_emu")
    if not os.path.exists(exe):
        from ..utils import build as _b
        _b.build_tool("cclo_emu")
    return exe


def spawn_engines(world, base_port=None, mem_mb=64, loopback=True, log_level=None, stderr=None):
    """Start one stand-alone engine process per rank (reference run.py: N x cclo_emu).  Rank r's engine
    listens for the other engines on base_port + r and for its driver on base_port + 1000 + r.
    Returns (procs, base_port); attach with `accl_b200.remote_rank(r, world, ctrl_port=base_port + 1000 + r)`."""
    base_port = base_port or free_port_block(world, also_at=(1000,))
    exe = engine_binary()
    procs = []
    for r in range(world):
        cmd = [exe, "--rank", str(r), "--world", str(world), "--base-port", str(base_port), "--ctrl-port",
               str(base_port + 1000 + r), "--mem-mb", str(mem_mb)]
        if not loopback:
            cmd.append("--no-kernel-loopback")
        if log_level is not None:
            cmd += ["--log-level", str(log_level)]
        procs.append(subprocess.Popen(cmd, stderr=stderr))
    return procs, base_port


def selftest():
    """What each rank runs under --selftest: BASELINE config #1 (send/recv + allreduce fp32)."""
    import torch
    import accl_b200 as A
    a = A.socket_rank()
    a.initialize(n_egr_rx_bufs=16, egr_rx_buf_size=1024, max_egr_size=1024, max_rndzv_size=32768)
    r, w = a.rank, a.world
    for n in (16, 300, 20000):  # eager, segmented eager, rendezvous
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.host[:] = torch.arange(n, dtype=torch.float32) + r
        if w > 1:
            nxt, prv = (r + 1) % w, (r - 1) % w
            req = a.send(s, n, nxt, tag=3, run_async=True)
            a.recv(d, n, prv, tag=3)
            req.wait()
            assert torch.equal(d.host, torch.arange(n, dtype=torch.float32) + prv), "send/recv mismatch"
        a.allreduce(s, d, n, A.SUM)
        ref = sum(torch.arange(n, dtype=torch.float32) + q for q in range(w))
        assert torch.allclose(d.host, ref), "allreduce mismatch"
    a.barrier()
    a.set_one_hop_schedules(True)   # the same all-reduces on the B200 backend's one-hop schedules
    for n in (300, 20000):
        s, d = a.create_buffer(n), a.create_buffer(n)
        s.host[:] = torch.arange(n, dtype=torch.float32) + r
        a.allreduce(s, d, n, A.SUM)
        assert torch.allclose(d.host, sum(torch.arange(n, dtype=torch.float32) + q for q in range(w))), "one-hop allreduce mismatch"
    a.barrier()
    a.deinit()
    print(f"emulator rank {r}/{w}: ok", flush=True)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-n", "--nranks", type=int, default=2)
    ap.add_argument("-p", "--port", type=int, default=0, help="base port (rank r listens on port + r); 0 = pick")
    ap.add_argument("--selftest", action="store_true")
    ap.add_argument("--engines", action="store_true", help="run only the cclo_emu engine processes until interrupted")
    ap.add_argument("--child-selftest", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    if a.child_selftest:
        selftest()
        return 0
    if a.engines:
        procs, base = spawn_engines(a.nranks, a.port or None)
        print(f"engines up: fabric ports {base}..{base + a.nranks - 1}, control ports {base + 1000}..{base + 1000 + a.nranks - 1}",
              flush=True)
        try:
            return max(p.wait() for p in procs)
        except KeyboardInterrupt:
            for p in procs:
                p.terminate()
            return 130
    argv = [sys.executable, "-m", "accl_b200.models.emulator", "--child-selftest"] if a.selftest else [c for c in a.cmd if c != "--"]
    if not argv:
        ap.error("nothing to run: pass --selftest or a command after --")
    rcs = launch(a.nranks, argv, a.port or None, timeout=300)
    print("exit codes:", rcs)
    return 0 if all(rc == 0 for rc in rcs) else 1


if __name__ == "__main__":
    sys.exit(main())
