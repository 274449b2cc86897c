#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU). A plain run of the same command goes first.
set -o pipefail
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
CMD2="python bench.py --steps 1 --warmup 1 --batch-per-gpu 2 --no-cpu-baseline"   # small batch: ncu replays save/restore memory
mkdir -p gpurun_out
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
$CMD2 > gpurun_out/plain2.log 2> gpurun_out/plain2.err || { echo "plain run 2 failed"; tail -5 gpurun_out/plain2.err; exit 1; }
cat gpurun_out/plain.log | cut -c1-400
# every launch of ONE timed train step with its device time (cold-cache, serialised: compare shares).
# 3 warm-up steps (~1570 launches each) are skipped.
ncu --metrics gpu__time_duration.sum --clock-control none -s 4800 -c 1600 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# full captures: the tcgen05 GEMM family, spatial attention, PEG, LayerNorm backward, patch gather
ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_kernel|attn_fwd_kernel|attn_bwd_dq|attn_bwd_dkv|peg_tile_kernel|layernorm_bwd|patch_norm" \
    -s 150 -c 14 -o gpurun_out/prof_full $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | head -30; du -sh gpurun_out
