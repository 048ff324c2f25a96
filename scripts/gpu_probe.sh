mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1; nvidia-smi topo -m >> gpurun_out/smi.txt 2>&1
for r in 0 1; do RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r MASTER_ADDR=127.0.0.1 MASTER_PORT=29400 timeout 150 ./build/accl_probe 512 > gpurun_out/probe_r$r.log 2>&1 & done; wait
cat gpurun_out/probe_r0.log
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench/nccl_baseline.py --min-log2 10 --max-log2 30 --step 2 --out gpurun_out/nccl_2gpu.csv > gpurun_out/nccl_2gpu.log 2>&1
grep -iE "NVLS|Using network|comm 0x.*nranks" gpurun_out/nccl_2gpu.log | head -8
grep '"op"' gpurun_out/nccl_2gpu.log | tail -40
