#!/bin/bash
# 8-GPU run A: correctness on real NVLS, knob grid for the large-message algorithms, first sweep, bench.py
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv,noheader | head -8
timeout 420 python -m pytest tests/test_cuda_oneway.py tests/test_cuda.py tests/test_cuda_engine.py -q --timeout 120 -x \
  -k "allreduce_sizes or rooted_from_every_root or allgather_reduce_scatter or nvls or switch or wire or collective_matrix or two_streams or large_rendezvous" 2>&1 | tail -12 | tee gpurun_out/8a_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29511 bench/tune.py --what allreduce,reduce,bcast --mb 256 --out gpurun_out/tune8_fp32.jsonl 2>&1 | grep -v -i warning | tail -140 | cut -c1-260
timeout 200 $T --master-port 29521 bench/tune.py --what allreduce --mb 256 --dtype bfloat16 --quick --out gpurun_out/tune8_bf16.jsonl 2>&1 | grep -v -i warning | tail -20 | cut -c1-260
timeout 420 $T --master-port 29531 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --graph --out gpurun_out/sweep8_direct.csv 2>&1 | grep -v -i warning | tail -40 | cut -c1-420
timeout 200 $T --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench8_direct.json | cut -c1-1500
