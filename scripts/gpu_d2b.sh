#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests/test_cuda_oneway.py tests/test_cuda_engine.py -q --timeout 150 -x 2>&1 | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench/diag.py --egr-kb 4096 --graph --ops nop,allreduce,allgather --out gpurun_out/diag2_oneway_b.jsonl 2>&1 | grep -v -i warning | tail -40 | cut -c1-330
ACCL_PDL=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 \
  bench/diag.py --egr-kb 4096 --graph --modes direct --ops nop,allreduce --sizes 1024,65536,1048576 --out gpurun_out/diag2_nopdl.jsonl 2>&1 | grep -v -i warning | tail -8 | cut -c1-330
