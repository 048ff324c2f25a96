#!/bin/bash
# 4-GPU run (final numbers): sweep direct, rooted, fused GEMM shapes, vadd, bench.py
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
F='grep -v -i warning'
rm -f gpurun_out/gemm_rs_4gpu.jsonl gpurun_out/vadd_4gpu.jsonl
timeout 300 $T --master-port 29511 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --graph --out gpurun_out/sweep4_direct.csv 2>&1 | $F | tail -34 | cut -c1-330
timeout 200 $T --master-port 29531 bench/sweep.py --dtype bfloat16 --ops bcast,reduce,scatter,gather,alltoall --min-log2 12 --max-log2 28 --step 4 --out gpurun_out/sweep4_rooted_bf16.csv 2>&1 | $F | tail -26 | cut -c1-260
timeout 300 $T --master-port 29561 bench/gemm_rs.py --shapes 8192x8192x2048,8192x8192x4096,16384x8192x2048,8192x4096x4096,4096x8192x2048,16384x16384x1024,8192x8192x2048:f32 \
   --check --out gpurun_out/gemm_rs_4gpu.jsonl 2>&1 | grep '^{' | cut -c1-420
timeout 120 $T --master-port 29571 bench/vadd.py --min-log2 20 --max-log2 28 --step 4 --out gpurun_out/vadd_4gpu.jsonl 2>&1 | $F | tail -4 | cut -c1-300
timeout 150 $T --master-port 29591 bench.py --gpus 4 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench4_direct.json | cut -c1-1500
