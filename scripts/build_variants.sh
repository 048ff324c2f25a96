#!/bin/bash
# Build every opt-in variant of the extension next to the default one (build/variants/<name>); run on the build
# host before `gpurun -- 'bash scripts/next_gpu_session.sh N'`.  About one minute each.
set -e
cd "$(dirname "$0")/.."
python -m accl_b200.utils.build
python -m accl_b200.utils.build --variant timing --define ACCL_PHASE_TIMING
python -m accl_b200.utils.build --variant rpush  --define ACCL_EXPERIMENTAL_REDUCE_PUSH
python -m accl_b200.utils.build --variant bflags --define ACCL_EXPERIMENTAL_BCAST_FLAGS
python -m accl_b200.utils.build --variant gemm2  --define ACCL_EXPERIMENTAL_GEMM_2CTA
python -m accl_b200.utils.build --variant hyb3   --define ACCL_EXPERIMENTAL_HYBRID_AR --define ACCL_HYBRID_P2P_16THS=3
ls -la build/variants/*/accl_b200/_C*.so
