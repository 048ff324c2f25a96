#!/bin/bash
# 4-GPU run (final numbers, ordered by priority: the GPU budget may cut the tail): fused GEMM shapes, direct sweep, DDP vs NCCL,
# vadd plugin, compressed-wire all-reduce
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export BENCH_WATCHDOG_S=60
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
F='grep -v -i warning'
rm -f gpurun_out/gemm_rs_4gpu.jsonl gpurun_out/vadd_4gpu.jsonl gpurun_out/ddp_4gpu.jsonl
timeout 200 $T --master-port 29561 bench/gemm_rs.py --shapes 8192x8192x2048,8192x8192x4096,16384x8192x2048,8192x4096x4096,4096x8192x2048,16384x16384x1024,8192x8192x2048:f32 \
   --check --out gpurun_out/gemm_rs_4gpu.jsonl 2>&1 | grep '^{' | cut -c1-420
timeout 200 $T --master-port 29511 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 4 --batches 3 --out gpurun_out/sweep4_direct.csv 2>&1 | $F | grep '^{' | cut -c1-330
timeout 100 $T --master-port 29585 bench/ddp.py --backend nccl --out gpurun_out/ddp_4gpu.jsonl 2>&1 | grep '^{' | cut -c1-300
timeout 120 $T --master-port 29589 bench/ddp.py --backend accl --heap-buckets --out gpurun_out/ddp_4gpu.jsonl > gpurun_out/ddp_accl_heap_4gpu.log 2>&1; grep '^{\|Error' gpurun_out/ddp_accl_heap_4gpu.log | tail -3 | cut -c1-300
timeout 100 $T --master-port 29571 bench/vadd.py --min-log2 20 --max-log2 28 --step 4 --out gpurun_out/vadd_4gpu.jsonl 2>&1 | grep '^{' | cut -c1-300
timeout 100 $T --master-port 29541 bench/sweep.py --compress bfloat16 --ops allreduce --min-log2 22 --max-log2 30 --step 4 --batches 3 --out gpurun_out/sweep4_wire_bf16.csv 2>&1 | $F | grep '^{' | cut -c1-260
timeout 100 $T --master-port 29551 bench/sweep.py --compress float8_e4m3 --ops allreduce --min-log2 22 --max-log2 30 --step 4 --batches 3 --out gpurun_out/sweep4_wire_fp8.csv 2>&1 | $F | grep '^{' | cut -c1-260
