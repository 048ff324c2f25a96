#!/bin/bash
# rooted collectives (BASELINE config #3) + headline bench.  usage: gpu_rooted8.sh N
N=${1:-8}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
fmt() { python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-15s %11d  accl %9.1f us %8.1f GB/s | nccl %9.1f us %8.1f GB/s | x%.2f' % (d['op'],d['bytes'],d['accl_us'],d['accl_busbw'],d.get('nccl_us',0),d.get('nccl_busbw',0),d.get('speedup',0)))
"; }
timeout 300 $TR --master-port 29557 bench/sweep.py --ops bcast,reduce,scatter,gather,alltoall --dtype bfloat16 --min-log2 12 --max-log2 28 --step 4 --out gpurun_out/sweep_${N}gpu_rooted_bf16.csv 2>/dev/null | fmt
echo "--- fp16"
timeout 300 $TR --master-port 29558 bench/sweep.py --ops bcast,reduce,scatter,gather --dtype float16 --min-log2 16 --max-log2 28 --step 6 --out gpurun_out/sweep_${N}gpu_rooted_fp16.csv 2>/dev/null | fmt
echo "--- scatter, 64 CTAs"
timeout 200 $TR --master-port 29559 bench/sweep.py --ops scatter --dtype bfloat16 --min-log2 24 --max-log2 28 --step 4 --max-ctas 64 --no-nccl 2>/dev/null | fmt
timeout 300 $TR --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_${N}gpu.json; python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu.json'));print('value',d['value'],'e2e',d.get('e2e',{}).get('value'),d.get('e2e',{}).get('ms_per_step'),'nccl',d.get('nccl_same_run'))"
