"""dtype tables: accl DataType <-> torch / numpy dtypes."""
import numpy as np
import torch

from .. import _C

DataType = _C.DataType

TORCH_TO_ACCL = {
    torch.float16: DataType.float16,
    torch.float32: DataType.float32,
    torch.float64: DataType.float64,
    torch.int32: DataType.int32,
    torch.int64: DataType.int64,
    torch.bfloat16: DataType.bfloat16,
    torch.int8: DataType.int8,
    torch.float8_e4m3fn: DataType.float8_e4m3,
    torch.float8_e5m2: DataType.float8_e5m2,
}
ACCL_TO_TORCH = {v: k for k, v in TORCH_TO_ACCL.items()}

NUMPY_TO_ACCL = {
    np.dtype(np.float16): DataType.float16,
    np.dtype(np.float32): DataType.float32,
    np.dtype(np.float64): DataType.float64,
    np.dtype(np.int32): DataType.int32,
    np.dtype(np.int64): DataType.int64,
    np.dtype(np.int8): DataType.int8,
}


def to_accl(dtype):
    """Accepts an accl DataType, a torch dtype, a numpy dtype or a dtype name."""
    if isinstance(dtype, DataType):
        return dtype
    if isinstance(dtype, torch.dtype):
        return TORCH_TO_ACCL[dtype]
    if isinstance(dtype, str):
        if hasattr(DataType, dtype):
            return getattr(DataType, dtype)
        return to_accl(getattr(torch, dtype))
    return NUMPY_TO_ACCL[np.dtype(dtype)]


def to_torch(dtype):
    return ACCL_TO_TORCH[to_accl(dtype)]


def itemsize(dtype):
    return _C.dtype_bytes(to_accl(dtype))
