from .plugins import gemm_reduce_scatter, stream_loopback, stream_pull, vadd_allreduce, vadd_put  # noqa: F401
