#!/bin/bash
# 1 GPU: whole GPU suite with ranks sharing cuda:0, then the tcgen05 GEMM variants against cuBLAS
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 900 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -25 | tee gpurun_out/g1_pytest.log
rm -f gpurun_out/gemm_1gpu.jsonl
for v in 1 2; do
  timeout 200 python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant $v --check --out gpurun_out/gemm_1gpu.jsonl 2>&1 | tail -2 | cut -c1-600
done
timeout 200 python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant 2 --f32 --check --out gpurun_out/gemm_1gpu.jsonl 2>&1 | tail -2 | cut -c1-600
timeout 200 python bench/gemm_rs.py --gm 4096 --gn 4096 --gk 4096 --variant 2 --check --out gpurun_out/gemm_1gpu.jsonl 2>&1 | tail -2 | cut -c1-600
timeout 200 python bench/gemm_rs.py --gm 16384 --gn 8192 --gk 2048 --variant 2 --check --out gpurun_out/gemm_1gpu.jsonl 2>&1 | tail -2 | cut -c1-600
