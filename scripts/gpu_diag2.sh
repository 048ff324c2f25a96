#!/bin/bash
# 2-GPU diagnosis: per-call time decomposition (event vs device duration vs NCCL), direct and engine modes,
# then the NVLS (multimem) path forced on at 2 ranks.
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | head -4
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench/diag.py --out gpurun_out/diag_2gpu.jsonl 2>&1 | grep -v -i warning | tail -60
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 \
  bench/diag.py --modes direct --ops allreduce --sizes 4194304,67108864 --nvls-min-ranks 2 --big-mb 256 --out gpurun_out/diag_2gpu_nvls.jsonl 2>&1 | grep -v -i warning | tail -12
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 \
  bench/diag.py --modes direct --ops allreduce --sizes 4194304,67108864 --big-mb 256 --max-ctas 128 --out gpurun_out/diag_2gpu_p2p.jsonl 2>&1 | grep -v -i warning | tail -12
