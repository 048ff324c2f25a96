from .plugins import gemm_reduce_scatter, vadd_allreduce  # noqa: F401
