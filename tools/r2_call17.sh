#!/bin/bash
# 4 GPUs, 4 independent processes: H2D copy durations of the e2e pipeline when several GPUs pull from the host at once
mkdir -p gpurun_out
for mode in stored fp32; do
  for i in 0 1 2 3; do
    CUDA_VISIBLE_DEVICES=$i timeout 600 python tools/e2e_timeline.py --mode $mode --steps 5 > gpurun_out/r2q_${mode}_$i.log 2>&1 &
  done
  wait
  echo "== $mode"; grep -v "Warn\|warn" gpurun_out/r2q_${mode}_0.log | grep "mode \|GB/s\|stream " | head -16
done
