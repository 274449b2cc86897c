#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU). A plain run of the same command goes first.
set -o pipefail
CMD="python bench.py --steps 1 --warmup 1 --batch-per-gpu 2 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
cat gpurun_out/plain.log
# every launch of one train step with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1800 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# full capture of the FeedForward GEGLU GEMM (largest forward GEMM) and the spatial attention forward
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 4 \
    -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 2 -c 1 \
    -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture exit $?"
ls -la gpurun_out
