#!/bin/bash
# 2-GPU follow-up: device-side clients (client_done), vadd plugin, heap pool + DDP buckets in the heap, ncu on rank 0
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
export BENCH_WATCHDOG_S=60
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='grep -v -i warning'
timeout 200 python -m pytest tests/test_cuda_plugins.py tests/test_cuda_engine.py -m gpu -q --timeout 100 2>&1 | tail -8 | tee gpurun_out/2i_pytest.log
rm -f gpurun_out/vadd_2gpu.jsonl gpurun_out/ddp_2gpu.jsonl
timeout 150 $T --master-port 29571 bench/vadd.py --min-log2 16 --max-log2 28 --step 4 --out gpurun_out/vadd_2gpu.jsonl > gpurun_out/vadd_2gpu.log 2>&1; grep '^{\|Timeout\|Error' gpurun_out/vadd_2gpu.log | tail -6 | cut -c1-300
timeout 100 $T --master-port 29585 bench/ddp.py --backend nccl --out gpurun_out/ddp_2gpu.jsonl 2>&1 | $F | tail -1 | cut -c1-300
timeout 120 $T --master-port 29586 bench/ddp.py --backend accl --out gpurun_out/ddp_2gpu.jsonl > gpurun_out/ddp_accl_2gpu.log 2>&1; grep '^{\|Error' gpurun_out/ddp_accl_2gpu.log | tail -3 | cut -c1-300
timeout 120 $T --master-port 29589 bench/ddp.py --backend accl --heap-buckets --out gpurun_out/ddp_2gpu.jsonl > gpurun_out/ddp_accl_heap_2gpu.log 2>&1; grep '^{\|Error' gpurun_out/ddp_accl_heap_2gpu.log | tail -5 | cut -c1-300
export ACCL_PDL=0
export NCU_REPLAY=application
export NCU_METRICS=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,launch__grid_size
export NCU_LOG=gpurun_out/ncu_coll_2gpu_a.csv
timeout 90 $T --master-port 29587 --no-python scripts/ncu_rank0.sh bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu_a.log 2>&1
echo "ncu A exit $?"; grep "==ERROR\|RuntimeError" gpurun_out/ncu_coll_2gpu_a.log | head -5 | cut -c1-300; wc -l $NCU_LOG
export NCU_METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size
export NCU_LOG=gpurun_out/ncu_coll_2gpu_b.csv
timeout 90 $T --master-port 29588 --no-python scripts/ncu_rank0.sh bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu_b.log 2>&1
echo "ncu B exit $?"; grep "==ERROR\|RuntimeError" gpurun_out/ncu_coll_2gpu_b.log | head -5 | cut -c1-300; wc -l $NCU_LOG
