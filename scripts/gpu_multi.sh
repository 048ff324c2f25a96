#!/bin/bash
# multi-GPU validation + sweeps.  usage: gpu_multi.sh N
N=${1:-4}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_cuda.py -m gpu -q --timeout 200 -k "nvls or allreduce or rooted or reduce_scatter" > gpurun_out/pytest_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${N}gpu.log
timeout 300 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_${N}gpu.json; cut -c1-330 gpurun_out/bench_${N}gpu.json; python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu.json'));print('e2e',d.get('e2e',{}).get('value'),'nccl',d.get('nccl_same_run'))"
timeout 500 $TR --master-port 29553 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --out gpurun_out/sweep_${N}gpu_nvls.csv 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-15s %11d  accl %9.1f us %8.1f GB/s | nccl %9.1f us %8.1f GB/s | x%.2f' % (d['op'],d['bytes'],d['accl_us'],d['accl_busbw'],d.get('nccl_us',0),d.get('nccl_busbw',0),d.get('speedup',0)))
    elif l.startswith('#'): print(l.strip())
"
echo "--- P2P only (no NVLS)"
timeout 400 $TR --master-port 29555 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 20 --max-log2 30 --step 2 --nvls-min-ranks 99 --no-nccl --out gpurun_out/sweep_${N}gpu_p2p.csv 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-15s %11d  accl %9.1f us %8.1f GB/s' % (d['op'],d['bytes'],d['accl_us'],d['accl_busbw']))
"
echo "--- rooted, bf16"
timeout 400 $TR --master-port 29557 bench/sweep.py --ops bcast,reduce,scatter,gather,alltoall --dtype bfloat16 --min-log2 12 --max-log2 28 --step 4 --out gpurun_out/sweep_${N}gpu_rooted_bf16.csv 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-15s %11d  accl %9.1f us %8.1f GB/s | nccl %9.1f us %8.1f GB/s | x%.2f' % (d['op'],d['bytes'],d['accl_us'],d['accl_busbw'],d.get('nccl_us',0),d.get('nccl_busbw',0),d.get('speedup',0)))
"
echo "--- gemm -> reduce_scatter"
timeout 300 $TR --master-port 29559 bench/gemm_rs.py --gm 8192 --gn 8192 --gk 2048 --check --out gpurun_out/gemm_rs_${N}gpu.jsonl 2>/dev/null | tail -1 | cut -c1-600
echo "--- vector-add plugin -> all-reduce (device-issued)"
timeout 300 $TR --master-port 29561 bench/vadd.py --min-log2 12 --max-log2 26 --step 2 --out gpurun_out/vadd_${N}gpu.jsonl 2>/dev/null | cut -c1-200
