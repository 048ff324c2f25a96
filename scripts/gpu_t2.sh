#!/bin/bash
# 2-GPU validation (real NVLink + NVLS): new suites, old suites, then the small-message decomposition
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_cuda_oneway.py tests/test_cuda_engine.py -q --timeout 150 2>&1 | tail -30 | tee gpurun_out/t2_new.log
timeout 600 python -m pytest tests/test_cuda.py tests/test_cuda_plugins.py tests/test_parallel.py -m gpu -q --timeout 150 2>&1 | tail -15 | tee gpurun_out/t2_old.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench/diag.py --egr-kb 4096 --graph --out gpurun_out/diag2_oneway.jsonl 2>&1 | grep -v -i warning | tail -60
