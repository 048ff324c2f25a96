#!/bin/bash
# 2-GPU run (final numbers): multi-GPU tests, sweeps direct / engine, fused GEMM shapes, vadd, NVLink counters, DDP vs NCCL,
# ncu capture of the collective kernels, bench.py direct + engine
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
F='grep -v -i warning'
timeout 420 python -m pytest tests/test_cuda_oneway.py tests/test_cuda_engine.py tests/test_cuda.py -m gpu -q --timeout 120 \
   -k "not dtypes" 2>&1 | tail -12 | tee gpurun_out/2f_pytest.log
timeout 300 $T --master-port 29511 bench/sweep.py --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 30 --step 2 --graph --out gpurun_out/sweep2_direct.csv 2>&1 | $F | tail -36 | cut -c1-330
timeout 200 $T --master-port 29521 bench/sweep.py --engine --engine-workers 64 --ops allreduce,allgather,reduce_scatter --min-log2 10 --max-log2 24 --step 2 --out gpurun_out/sweep2_engine.csv 2>&1 | $F | tail -26 | cut -c1-260
rm -f gpurun_out/gemm_rs_2gpu.jsonl gpurun_out/vadd_2gpu.jsonl gpurun_out/nvlink_traffic_2gpu.jsonl gpurun_out/ddp_2gpu.jsonl
timeout 300 $T --master-port 29561 bench/gemm_rs.py --shapes 8192x8192x4096,8192x8192x2048,16384x8192x2048,8192x4096x4096,4096x8192x4096,16384x16384x2048,8192x8192x4096:f32 \
   --check --out gpurun_out/gemm_rs_2gpu.jsonl 2>&1 | grep '^{' | cut -c1-420
timeout 150 $T --master-port 29571 bench/vadd.py --min-log2 16 --max-log2 28 --step 4 --out gpurun_out/vadd_2gpu.jsonl 2>&1 | $F | tail -6 | cut -c1-300
timeout 150 $T --master-port 29581 bench/nvlink_traffic.py --out gpurun_out/nvlink_traffic_2gpu.jsonl 2>&1 | $F | tail -6 | cut -c1-500
timeout 120 $T --master-port 29585 bench/ddp.py --backend nccl --out gpurun_out/ddp_2gpu.jsonl 2>&1 | $F | tail -1 | cut -c1-300
timeout 120 $T --master-port 29586 bench/ddp.py --backend accl --out gpurun_out/ddp_2gpu.jsonl 2>&1 | $F | tail -1 | cut -c1-300
M=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__block_size
timeout 300 ncu --replay-mode application --target-processes all --clock-control none -k regex:k_call --metrics $M --csv --log-file gpurun_out/ncu_coll_2gpu.csv \
   $T --master-port 29587 bench/ncu_target.py --plan gpurun_out/ncu_coll_2gpu_plan.json > gpurun_out/ncu_coll_2gpu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_coll_2gpu.log | cut -c1-300; wc -l gpurun_out/ncu_coll_2gpu.csv
timeout 150 $T --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench2_direct.json | cut -c1-1500
timeout 150 $T --master-port 29601 bench.py --gpus 2 --steps 20 --warmup 5 --engine --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench2_engine.json | cut -c1-700
