#!/bin/bash
# last 2-GPU check: GEMM test matrix, torch.distributed backend with the communication stream, DDP with heap-resident buckets
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 python -m pytest tests/test_cuda_plugins.py tests/test_cuda_engine.py -m gpu -q --timeout 100 -k "gemm or torch_distributed or heap or vadd" 2>&1 | tail -6 | tee gpurun_out/2j_pytest.log
timeout 90 $T --master-port 29589 bench/ddp.py --backend accl --heap-buckets --out gpurun_out/ddp_2gpu.jsonl > gpurun_out/ddp_accl_heap_2gpu.log 2>&1; grep '^{\|Error' gpurun_out/ddp_accl_heap_2gpu.log | tail -4 | cut -c1-400
