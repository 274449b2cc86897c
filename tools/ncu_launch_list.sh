#!/bin/bash
# every launch of one timed train step with its device time (run under gpurun, one GPU); kernels replayed from
# CUDA graphs are profiled per node.  Cold-cache, serialised: compare shares, not absolutes.
set -o pipefail
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
cut -c1-200 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 5300 -c 1700 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
