#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export PG_WATCHDOG_S=60
timeout 400 python -m pytest tests/test_cuda_engine.py -k "torch_distributed" -q --timeout 300 > gpurun_out/pg1.log 2>&1; grep -v "^$" gpurun_out/pg1.log | tail -80 | cut -c1-300
timeout 300 python -m pytest tests/test_cuda.py -k "segmentation or stream_ids" tests/test_cuda_oneway.py -k "graph or wire" -q --timeout 200 2>&1 | tail -8
for v in 2 1; do timeout 200 python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant $v --check --out gpurun_out/gemm_1gpu_b.jsonl 2>&1 | tail -1 | cut -c1-420; done
timeout 200 python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant 2 --f32 --check --out gpurun_out/gemm_1gpu_b.jsonl 2>&1 | tail -1 | cut -c1-420
