#!/bin/bash
# ncu full capture of the HBM-bound encoder kernels at bench size (run under gpurun, one GPU)
mkdir -p gpurun_out
python tools/time_elementwise.py > gpurun_out/elementwise_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/elementwise_plain.log; exit 1; }
cat gpurun_out/elementwise_plain.log
ncu --set full --clock-control none --import-source on -k regex:"peg_|layernorm_" -s 12 -c 12 -f -o gpurun_out/prof_elementwise \
    python tools/time_elementwise.py > gpurun_out/ncu_elementwise.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_elementwise.log
