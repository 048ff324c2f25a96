#!/bin/bash
# compute-sanitizer passes over the single-rank GPU paths (copy / combine / stream FIFO / tcgen05 GEMM plugin at world=1).
# Multi-rank kernels spin on each other and are validated by the numerics tests instead: the sanitizer's
# serialising instrumentation would only trip their watchdogs.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 --target-processes all \
    python -m pytest -q --timeout 200 -x \
      "tests/test_cuda.py::test_copy_combine_nop" \
      "tests/test_cuda_plugins.py::test_gemm_reduce_scatter_matches_fp32_reference[1]" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|hazard" gpurun_out/sanitizer_$tool.log | head -10
done
