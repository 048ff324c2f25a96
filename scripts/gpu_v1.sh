#!/bin/bash
# 1 GPU: whole GPU suite (ranks share cuda:0), tcgen05 GEMM variants vs cuBLAS, one ncu capture of the pair kernel,
# smoke(), compute-sanitizer over the new protocols
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 700 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -25 | tee gpurun_out/v1_pytest.log
rm -f gpurun_out/gemm_1gpu_v1.jsonl
timeout 300 python bench/gemm_rs.py --shapes 8192x8192x8192:v2,8192x8192x8192:v1,8192x8192x8192:f32:v2,4096x4096x4096:v2,16384x8192x2048:v2,8192x8192x2048:v2,2048x8192x8192:v2 \
   --check --out gpurun_out/gemm_1gpu_v1.jsonl 2>&1 | grep '^{' | cut -c1-330
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_plugin_gemm_rs -s 3 -c 1 -f -o gpurun_out/prof_gemm_pair2 \
   python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --variant 2 --iters 2 2>&1 | tail -2 | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/v1_smoke.log
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cuda_oneway.py tests/test_cuda_engine.py -q --timeout 160 \
   -k "allreduce_dtypes or send_send_then_recv" > gpurun_out/sanitizer_memcheck_r2.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/sanitizer_memcheck_r2.log | tail -8
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv | tee gpurun_out/v1_smi.txt
