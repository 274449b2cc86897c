#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_tower_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/r2c_text.log 2>&1
echo "== text tower tests exit $?"; tail -n 25 gpurun_out/r2c_text.log
timeout 900 python bench.py --steps 6 --warmup 4 --no-cpu-baseline > gpurun_out/r2c_bench_ctk.json 2> gpurun_out/r2c_bench_ctk.err
echo "== bench ctk exit $?"; cat gpurun_out/r2c_bench_ctk.json; tail -5 gpurun_out/r2c_bench_ctk.err
timeout 900 python bench.py --steps 6 --warmup 4 --no-cpu-baseline --text-dropout 0.1 > gpurun_out/r2c_bench_ctk_drop.json 2> gpurun_out/r2c_bench_ctk_drop.err
echo "== bench ctk dropout exit $?"; cat gpurun_out/r2c_bench_ctk_drop.json; tail -3 gpurun_out/r2c_bench_ctk_drop.err
timeout 600 python tools/e2e_probe.py --text-tower ctk > gpurun_out/r2c_e2e_probe.log 2>&1
echo "== e2e_probe exit $?"; grep '^{' gpurun_out/r2c_e2e_probe.log || tail -20 gpurun_out/r2c_e2e_probe.log
