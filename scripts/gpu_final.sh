#!/bin/bash
# last check of a tree on one GPU: test-suite, smoke(), headline bench (what the driver runs at round end)
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_1gpu.json; cut -c1-200 gpurun_out/bench_1gpu.json
