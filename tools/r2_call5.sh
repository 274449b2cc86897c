#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
echo "== bench n1 exit $?"; cat gpurun_out/r2e_bench_n1.json; tail -8 gpurun_out/r2e_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
echo "== bench ref exit $?"; cat gpurun_out/r2e_bench_ref.json; tail -4 gpurun_out/r2e_bench_ref.err
