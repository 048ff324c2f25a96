#!/bin/bash
# torchrun --no-python ... scripts/ncu_rank0.sh <script.py> [args]: rank 0 runs under ncu (metrics in $NCU_METRICS, CSV to
# $NCU_LOG), the other ranks run plain.  ncu serialises profiled kernels ACROSS processes, so profiling every rank of
# a kernel that waits for its peers deadlocks; with one profiled rank (and a metric set that fits one pass, so nothing is
# replayed) the peers simply wait a little longer.  Use NCU_REPLAY=application: kernel replay with several passes failed with
# UnknownError on kernels that touch the peer-mapped heap.
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --replay-mode ${NCU_REPLAY:-kernel} --clock-control none -k regex:k_call --metrics "$NCU_METRICS" --csv --log-file "$NCU_LOG" python "$@"
else
  exec python "$@"
fi
