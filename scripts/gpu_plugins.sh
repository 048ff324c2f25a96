#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32 ACCL_SEGV_TRACE=1
timeout ${1:-300} python -X faulthandler -m pytest tests/test_cuda_plugins.py -q --timeout 90 > gpurun_out/pytest_plugins.log 2>&1
echo "pytest rc=$?"
grep -nE "passed|failed|PASS|FAIL|Error|assert|Timeout|accl\] fatal" gpurun_out/pytest_plugins.log | head -30
