#!/bin/bash
# last check on 1 GPU: the whole GPU suite, then smoke()
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 100 python -m pytest tests -m gpu -q -x --timeout 60 2>&1 | tail -6 | tee gpurun_out/v2_pytest.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/v2_smoke.log
