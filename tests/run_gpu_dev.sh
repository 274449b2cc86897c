#!/bin/bash
# dev helper: run GPU test files one process each (a trapped kernel poisons its CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $f exit $?"
  tail -n 25 gpurun_out/$name.log
done
