#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export ACCL_SEGV_TRACE=1
timeout ${1:-600} python -X faulthandler -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"
grep -nE "passed|failed|Fatal|Error|error:|accl\] fatal|test_cuda.py::" gpurun_out/pytest_gpu.log | head -20
grep -n -A12 "native backtrace" gpurun_out/pytest_gpu.log | c++filt | cut -c1-180 | head -40
