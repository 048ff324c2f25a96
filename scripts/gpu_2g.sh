#!/bin/bash
# 2-GPU follow-up: vadd plugin (resident engine, adaptive chunks), DDP on the accl backend, ncu on rank 0 of the collective kernels
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='grep -v -i warning'
rm -f gpurun_out/vadd_2gpu.jsonl gpurun_out/ddp_2gpu.jsonl
timeout 150 $T --master-port 29571 bench/vadd.py --min-log2 16 --max-log2 28 --step 4 --out gpurun_out/vadd_2gpu.jsonl 2>&1 | $F | tail -5 | cut -c1-300
timeout 120 $T --master-port 29585 bench/ddp.py --backend nccl --out gpurun_out/ddp_2gpu.jsonl 2>&1 | $F | tail -1 | cut -c1-300
timeout 150 $T --master-port 29586 bench/ddp.py --backend accl --out gpurun_out/ddp_2gpu.jsonl > gpurun_out/ddp_accl_2gpu.log 2>&1; grep -v -i warning gpurun_out/ddp_accl_2gpu.log | tail -12 | cut -c1-300
export NCU_METRICS=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,launch__grid_size
export NCU_LOG=gpurun_out/ncu_coll_2gpu_a.csv
timeout 200 $T --master-port 29587 --no-python scripts/ncu_rank0.sh bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu_a.log 2>&1
echo "ncu A exit $?"; tail -2 gpurun_out/ncu_coll_2gpu_a.log | cut -c1-300; wc -l $NCU_LOG
export NCU_METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__grid_size
export NCU_LOG=gpurun_out/ncu_coll_2gpu_b.csv
timeout 200 $T --master-port 29588 --no-python scripts/ncu_rank0.sh bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu_b.log 2>&1
echo "ncu B exit $?"; tail -2 gpurun_out/ncu_coll_2gpu_b.log | cut -c1-300; wc -l $NCU_LOG
