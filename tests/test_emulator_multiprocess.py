"""One emulator rank per process over loopback TCP (SocketFabric): the
multi-process deployment of the CPU backend, as CI runs the reference's
emulator under mpirun (.github/workflows/build-and-test.yml:52-101)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_selftest_over_sockets(world):
    r = subprocess.run([sys.executable, "-m", "accl_b200.models.emulator", "-n", str(world), "--selftest"], cwd=ROOT,
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(": ok") == world
