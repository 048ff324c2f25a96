#!/bin/bash
# run the GPU test-suite; artefacts -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi -L
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout ${1:-900} python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -40
