#!/bin/bash
# 8-GPU run C (final numbers, trimmed to the GPU budget): knob grid incl. rooted variants, direct sweep with graph columns, fused GEMM shapes
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
F='grep -v -i warning'
rm -f gpurun_out/tune8c_*.jsonl gpurun_out/gemm_rs_8gpu.jsonl
timeout 150 $T --master-port 29501 bench/tune.py --what allreduce,reduce,bcast --mb 256 --quick --out gpurun_out/tune8c_fp32.jsonl 2>&1 | $F | grep '^{' | cut -c1-230 | tail -30
timeout 240 $T --master-port 29511 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --batches 3 --graph --out gpurun_out/sweep8_direct.csv 2>&1 | $F | grep '^{' | cut -c1-330
timeout 240 $T --master-port 29561 bench/gemm_rs.py --shapes 8192x8192x2048,8192x8192x1024,16384x8192x1024,8192x4096x4096,4096x8192x2048,16384x16384x1024,8192x8192x2048:f32 \
   --check --out gpurun_out/gemm_rs_8gpu.jsonl 2>&1 | grep '^{' | cut -c1-420
