"""accl_b200 — ACCL-compatible collectives for NVIDIA B200 (NVLink 5 / NVSwitch).

Backends: `cuda` (symmetric NVLink heap, hand-written sm_100a kernels, a
persistent engine kernel, NVLS multicast) and `emulator` (CPU model of the
engine; runs the whole host API and test matrix without a GPU).
"""
import importlib
import os
import sys


def _load_variant(name):
    """ACCL_VARIANT=<name>: use the extension of build/variants/<name> (python -m accl_b200.utils.build --variant
    <name> --define MACRO ...) instead of the default one — same Python sources, different compile-time switches."""
    import glob
    import importlib.machinery
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hits = glob.glob(os.path.join(root, "build", "variants", name, "accl_b200", "_C*.so"))
    if not hits:
        raise ImportError(f"ACCL_VARIANT={name}: no extension under build/variants/{name}; build it first")
    loader = importlib.machinery.ExtensionFileLoader("accl_b200._C", hits[0])
    spec = importlib.util.spec_from_loader("accl_b200._C", loader, origin=hits[0])
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    mod.__file__ = hits[0]
    return mod


def _load_native():
    if os.environ.get("ACCL_VARIANT"):
        return _load_variant(os.environ["ACCL_VARIANT"])
    try:
        return importlib.import_module("accl_b200._C")
    except ImportError as first:
        # build in-tree on first use (needs g++; nvcc for the CUDA backend)
        from .utils import build as _b
        have_nvcc = os.path.exists(_b.NVCC)
        try:
            _b.build(with_cuda=have_nvcc)
        except Exception as e:  # noqa: BLE001
            raise ImportError(f"accl_b200 native core is not built and building failed: {e}") from first
        return importlib.import_module("accl_b200._C")


_C = _load_native()
sys.modules["accl_b200._C"] = _C

from .core import (Accl, Buffer, BufferKind, DataType, GLOBAL_COMM, MAX, ReduceFunction, SUM, TAG_ANY,  # noqa: E402
                   cuda_rank, cuda_world, emulator_world, remote_rank, run_cuda_ranks, run_ranks, socket_rank)

__all__ = ["Accl", "Buffer", "BufferKind", "DataType", "GLOBAL_COMM", "MAX", "ReduceFunction", "SUM", "TAG_ANY",
           "emulator_world", "run_ranks", "socket_rank", "remote_rank", "cuda_world", "cuda_rank", "run_cuda_ranks", "_C"]
__version__ = "0.1.0"
