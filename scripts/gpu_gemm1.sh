#!/bin/bash
mkdir -p gpurun_out
python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --check --out gpurun_out/gemm_rs_1gpu.jsonl 2>&1 | tail -2
python bench/gemm_rs.py --gm 4096 --gn 4096 --gk 4096 --check --out gpurun_out/gemm_rs_1gpu.jsonl 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_1gpu.json; cat gpurun_out/bench_1gpu.json | cut -c1-600
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_plugin_gemm_rs -s 3 -c 1 -o gpurun_out/prof_gemm_rs python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --iters 2 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_call -s 6 -c 1 -o gpurun_out/prof_call_copy python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/ncu_call.log 2>&1; tail -2 gpurun_out/ncu_call.log
ls -la gpurun_out/*.ncu-rep
