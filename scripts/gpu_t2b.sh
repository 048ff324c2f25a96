#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests/test_cuda_engine.py -k "torch_distributed" -x -q --timeout 300 > gpurun_out/t2b_pg.log 2>&1; tail -60 gpurun_out/t2b_pg.log | cut -c1-400
timeout 600 python -m pytest tests/test_cuda.py -k "stream_ids or segmentation or in_place or pipelined or chunked" tests/test_cuda_oneway.py -q --timeout 200 2>&1 | tail -15 | tee gpurun_out/t2b_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench/diag.py --egr-kb 4096 --graph --out gpurun_out/diag2_c.jsonl 2>&1 | grep -v -i warning | tail -45 | cut -c1-330
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | cut -c1-300
