#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU). A plain run of the same command goes first.
set -o pipefail
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
cat gpurun_out/plain.log | cut -c1-400
# every launch of ONE timed train step with its device time (cold-cache, serialised: compare shares).
# 3 warm-up steps (~1570 launches each) are skipped.
ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 1600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# full captures: the tcgen05 GEMM family, spatial attention, PEG, LayerNorm backward, patch gather
ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_kernel|attn_fwd_kernel|attn_bwd_dq|attn_bwd_dkv|peg_tile_kernel|layernorm_bwd|patch_norm" \
    -s 60 -c 48 -o gpurun_out/prof_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | head -30
