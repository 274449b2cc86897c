#!/bin/bash
# ncu launch list of the device-timed region of `bench.py --steps 1` (every kernel, CUDA-graph nodes included):
# gpu__time_duration per launch, cold-cache and serialised - compare shares, not absolutes.  Output: a small CSV.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-eager"
CTK_BENCH_PROFILER_RANGE=1 timeout 280 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --graph-profiling node \
    --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list exit $?"; grep -c "gpu__time_duration" gpurun_out/r2_launches.csv; tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-300; ls -la gpurun_out
