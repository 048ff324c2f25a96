#!/bin/bash
# 1-GPU validation of the protocol logic (ranks share cuda:0): one-way protocols, engine, then the old suites
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_cuda_oneway.py -q --timeout 150 -x 2>&1 | tail -25 | tee gpurun_out/t1_oneway.log
timeout 900 python -m pytest tests/test_cuda_engine.py -q --timeout 150 2>&1 | tail -40 | tee gpurun_out/t1_engine.log
timeout 900 python -m pytest tests/test_cuda.py tests/test_cuda_plugins.py tests/test_parallel.py -m gpu -q --timeout 150 2>&1 | tail -25 | tee gpurun_out/t1_old.log
