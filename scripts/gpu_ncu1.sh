#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export PG_WATCHDOG_S=60
timeout 300 python -m pytest tests/test_cuda_engine.py -k "torch_distributed" -q --timeout 250 > gpurun_out/pg1.log 2>&1; grep -v "^$" gpurun_out/pg1.log | tail -30 | cut -c1-300
timeout 200 python bench/gemm_rs.py --gm 16384 --gn 8192 --gk 2048 --variant 2 --check --out gpurun_out/gemm_1gpu_c.jsonl 2>&1 | tail -1 | cut -c1-420
timeout 200 python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant 2 --out gpurun_out/gemm_1gpu_c.jsonl 2>&1 | tail -1 | cut -c1-420
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_plugin_gemm_rs -s 3 -c 1 -f -o gpurun_out/prof_gemm_pair \
   python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant 2 --iters 2 2>&1 | tail -3
