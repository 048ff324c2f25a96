#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export BENCH_WATCHDOG_S=45
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
rm -f gpurun_out/vadd_2gpu.jsonl
for lg in 16 20 24 28; do
  timeout 80 $T --master-port 295$lg bench/vadd.py --min-log2 $lg --max-log2 $lg --out gpurun_out/vadd_2gpu.jsonl > gpurun_out/vadd_dbg_$lg.log 2>&1
  echo "vadd 2^$lg exit $?"; grep -v -i "warning\|^\*\|OMP_NUM" gpurun_out/vadd_dbg_$lg.log | tail -25 | cut -c1-250
done
T2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export ACCL_PDL=0   # ncu cannot profile launches that carry the programmatic-stream-serialization attribute
export NCU_METRICS=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,launch__grid_size
export NCU_LOG=gpurun_out/ncu_coll_2gpu_a.csv
timeout 150 $T2 --master-port 29587 --no-python scripts/ncu_rank0.sh bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu_a.log 2>&1
echo "ncu A exit $?"; grep "==ERROR\|RuntimeError" gpurun_out/ncu_coll_2gpu_a.log | head -5 | cut -c1-300; wc -l $NCU_LOG
export NCU_METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__grid_size
export NCU_LOG=gpurun_out/ncu_coll_2gpu_b.csv
timeout 150 $T2 --master-port 29588 --no-python scripts/ncu_rank0.sh bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu_b.log 2>&1
echo "ncu B exit $?"; grep "==ERROR\|RuntimeError" gpurun_out/ncu_coll_2gpu_b.log | head -5 | cut -c1-300; wc -l $NCU_LOG
