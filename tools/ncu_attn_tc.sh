#!/bin/bash
# ncu full capture of the attention kernels at bench size (B=8): spatial tcgen05 fwd / dQ / dK,dV, bias-table
# gradient, temporal TMA+MMA fwd / bwd.  Run under gpurun, one GPU.
mkdir -p gpurun_out
python tools/dev_attn_tc.py > gpurun_out/attn_tc_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/attn_tc_plain.log; exit 1; }
python tools/dev_attn_seq24.py > gpurun_out/attn_seq24_plain.log 2>&1 || { echo "plain run 2 failed"; tail -5 gpurun_out/attn_seq24_plain.log; exit 1; }
tail -1 gpurun_out/attn_tc_plain.log; tail -1 gpurun_out/attn_seq24_plain.log
BENCH_ONLY=1 ncu --set full --clock-control none --import-source on  -k regex:"attn_" -s 6 -c 5 -f -o gpurun_out/prof_attn_tc \
    python tools/dev_attn_tc.py > gpurun_out/ncu_attn_tc.log 2>&1
echo "ncu spatial exit $?"
BENCH_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc" -s 1 -c 2 -f -o gpurun_out/prof_attn_tc_fwd \
    python tools/dev_attn_tc.py > gpurun_out/ncu_attn_tc_fwd.log 2>&1
echo "ncu spatial fwd exit $?"
BENCH_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:"attn_seq24" -s 3 -c 2 -f -o gpurun_out/prof_attn_seq24 \
    python tools/dev_attn_seq24.py > gpurun_out/ncu_attn_seq24.log 2>&1
echo "ncu temporal exit $?"
