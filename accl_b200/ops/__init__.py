from .plugins import gemm_reduce_scatter, stream_loopback, vadd_allreduce  # noqa: F401
