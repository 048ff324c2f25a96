#!/bin/bash
# 8-GPU run B (final numbers): sweeps direct / engine / rooted / compressed, fused GEMM shapes, vadd plugin, NVLink counters, bench.py
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
F='grep -v -i warning'
timeout 420 $T --master-port 29511 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --graph --out gpurun_out/sweep8_direct.csv 2>&1 | $F | tail -36 | cut -c1-460
timeout 240 $T --master-port 29521 bench/sweep.py --engine --engine-workers 64 --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 24 --step 2 --out gpurun_out/sweep8_engine.csv 2>&1 | $F | tail -26 | cut -c1-360
timeout 300 $T --master-port 29531 bench/sweep.py --dtype bfloat16 --ops bcast,reduce,scatter,gather,alltoall --min-log2 12 --max-log2 28 --step 4 --out gpurun_out/sweep8_rooted_bf16.csv 2>&1 | $F | tail -27 | cut -c1-360
timeout 240 $T --master-port 29541 bench/sweep.py --compress bfloat16 --ops allreduce --min-log2 20 --max-log2 30 --step 2 --out gpurun_out/sweep8_wire_bf16.csv 2>&1 | $F | tail -8 | cut -c1-360
timeout 240 $T --master-port 29551 bench/sweep.py --compress float8_e4m3 --ops allreduce --min-log2 20 --max-log2 30 --step 2 --out gpurun_out/sweep8_wire_fp8.csv 2>&1 | $F | tail -8 | cut -c1-360
rm -f gpurun_out/gemm_rs_8gpu.jsonl
for shape in "8192 8192 2048" "8192 8192 1024" "16384 8192 1024" "8192 4096 4096" "4096 8192 2048" "16384 16384 1024"; do
  set -- $shape
  timeout 150 $T --master-port 29561 bench/gemm_rs.py --gm $1 --gn $2 --gk $3 --check --out gpurun_out/gemm_rs_8gpu.jsonl 2>&1 | $F | tail -1 | cut -c1-700
done
timeout 150 $T --master-port 29561 bench/gemm_rs.py --gm 8192 --gn 8192 --gk 2048 --f32 --check --out gpurun_out/gemm_rs_8gpu.jsonl 2>&1 | $F | tail -1 | cut -c1-700
timeout 200 $T --master-port 29571 bench/vadd.py --min-log2 16 --max-log2 28 --step 4 --out gpurun_out/vadd_8gpu.jsonl 2>&1 | $F | tail -6 | cut -c1-300
timeout 200 $T --master-port 29581 bench/nvlink_traffic.py --out gpurun_out/nvlink_traffic_8gpu.jsonl 2>&1 | $F | tail -6 | cut -c1-500
timeout 200 $T --master-port 29591 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench8_direct.json | cut -c1-1800
timeout 200 $T --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 --engine --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench8_engine.json | cut -c1-900
