#!/bin/bash
# 2-GPU confirmation: multi-GPU tests, smoke, bench N=2, sweep vs NCCL
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --out gpurun_out/sweep_2gpu.csv 2>&1 | grep -v Warning | tail -40
