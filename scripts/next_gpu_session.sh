#!/bin/bash
# First GPU session of the next round (docs/roadmap.md): build the variants on the build host first:
#   python -m accl_b200.utils.build --variant timing --define ACCL_PHASE_TIMING
#   python -m accl_b200.utils.build --variant rpush  --define ACCL_EXPERIMENTAL_REDUCE_PUSH
#   python -m accl_b200.utils.build --variant bflags --define ACCL_EXPERIMENTAL_BCAST_FLAGS
#   python -m accl_b200.utils.build --variant gemm2  --define ACCL_EXPERIMENTAL_GEMM_2CTA
#   python -m accl_b200.utils.build --variant hyb3   --define ACCL_EXPERIMENTAL_HYBRID_AR --define ACCL_HYBRID_P2P_16THS=3
# then: gpurun --gpus 4 -- 'bash scripts/next_gpu_session.sh 4'
N=${1:-4}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "--- 1. where the small-message time goes (ACCL_PHASE_TIMING build)"
ACCL_VARIANT=timing timeout 200 $TR --master-port 29571 scripts/phase_timing.py 2>/dev/null | tee gpurun_out/phase_timing_${N}gpu.txt
echo "--- 2. write-only reduce (experimental build): correctness, then the rooted sweep"
ACCL_VARIANT=rpush timeout 200 python -m pytest tests/test_cuda.py -q --timeout 150 -k "large_reduce or rooted" 2>&1 | tail -2
ACCL_VARIANT=rpush timeout 300 $TR --master-port 29573 bench/sweep.py --ops reduce --dtype bfloat16 --min-log2 23 --max-log2 28 --step 1 --out gpurun_out/sweep_${N}gpu_reduce_push.csv 2>/dev/null | grep '^{' | cut -c1-220
echo "--- 2b. flag-driven pipelined bcast (experimental build)"
ACCL_VARIANT=bflags timeout 200 python -m pytest tests/test_cuda.py -q --timeout 150 -k "large_bcast" 2>&1 | tail -2
ACCL_VARIANT=bflags timeout 300 $TR --master-port 29574 bench/sweep.py --ops bcast --dtype bfloat16 --min-log2 27 --max-log2 30 --step 1 --out gpurun_out/sweep_${N}gpu_bcast_flags.csv 2>/dev/null | grep '^{' | cut -c1-220
echo "--- 2c. CTA-pair GEMM (experimental build): numerics first (1 GPU is enough), then throughput"
ACCL_VARIANT=gemm2 timeout 200 python -m pytest tests/test_cuda_plugins.py -q --timeout 120 -k "gemm" 2>&1 | tail -2
ACCL_VARIANT=gemm2 timeout 200 python bench/gemm_rs.py --gm 8192 --gn 8192 --gk 8192 --check 2>&1 | tail -1 | cut -c1-400
echo "--- 2d. hybrid switch + peer all-reduce (experimental build)"
ACCL_VARIANT=hyb3 timeout 200 python -m pytest tests/test_cuda.py -q --timeout 150 -k "allreduce" 2>&1 | tail -2
ACCL_VARIANT=hyb3 timeout 300 $TR --master-port 29576 bench/sweep.py --ops allreduce --min-log2 24 --max-log2 30 --step 2 --no-nccl --out gpurun_out/sweep_${N}gpu_hybrid.csv 2>/dev/null | grep '^{' | cut -c1-200
echo "--- 3. default build: same reduce sizes for comparison, vadd plugin timing, fuzz"
timeout 300 $TR --master-port 29575 bench/sweep.py --ops reduce,bcast --dtype bfloat16 --min-log2 23 --max-log2 30 --step 1 --out gpurun_out/sweep_${N}gpu_rooted_default.csv 2>/dev/null | grep '^{' | cut -c1-220
timeout 300 $TR --master-port 29577 bench/vadd.py --min-log2 12 --max-log2 26 --step 2 --out gpurun_out/vadd_${N}gpu.jsonl 2>/dev/null | cut -c1-200
timeout 300 python scripts/gpu_fuzz.py 150 $N 2>&1 | tail -3
