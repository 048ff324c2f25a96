#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench/diag.py --egr-kb 4096 --graph --out gpurun_out/diag2_oneway.jsonl 2>&1 | grep -v -i warning | tail -70
