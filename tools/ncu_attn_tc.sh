#!/bin/bash
# ncu full capture of the tcgen05 attention kernels at bench size (run under gpurun, one GPU)
mkdir -p gpurun_out
python tools/dev_attn_tc.py > gpurun_out/attn_tc_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/attn_tc_plain.log; exit 1; }
tail -2 gpurun_out/attn_tc_plain.log
ncu --set full --clock-control none --import-source on -k regex:"attn_.*_tc_kernel" -s 8 -c 2 -f -o gpurun_out/prof_attn_tc \
    python tools/dev_attn_tc.py > gpurun_out/ncu_attn_tc.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn_tc.log
