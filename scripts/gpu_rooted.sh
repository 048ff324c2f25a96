#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_cuda.py -m gpu -q --timeout 200 -k "switch or nvls_paths or distributed" 2>&1 | tail -3
timeout 300 $TR --master-port 29557 bench/sweep.py --ops bcast,reduce --dtype bfloat16 --min-log2 20 --max-log2 30 --step 2 --out gpurun_out/sweep_${N}gpu_rooted2_bf16.csv 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('%-15s %11d  accl %9.1f us %8.1f GB/s | nccl %9.1f us %8.1f GB/s | x%.2f' % (d['op'],d['bytes'],d['accl_us'],d['accl_busbw'],d.get('nccl_us',0),d.get('nccl_busbw',0),d.get('speedup',0)))
"
timeout 300 $TR --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_${N}gpu_b.json; python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu_b.json'));print('value',d['value'],'e2e',d.get('e2e',{}).get('value'),d.get('e2e',{}).get('ms_per_step'),'nccl',d.get('nccl_same_run'))"
