#!/bin/bash
# final validation of the round on one GPU: smoke, whole GPU suite, both bench arms
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
echo "== smoke exit $?"; grep "smoke:" gpurun_out/r2z_smoke.log
timeout 1700 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2z_suite.log 2>&1
echo "== suite exit $?"; tail -n 3 gpurun_out/r2z_suite.log | cut -c1-200
timeout 1200 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
echo "== bench exit $?"; cut -c1-300 gpurun_out/r2z_bench_n1.json; grep -v "Warn\|warn\|run_backward" gpurun_out/r2z_bench_n1.err | tail -4
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.err
echo "== bench ref exit $?"; cut -c1-300 gpurun_out/r2z_bench_ref.json
TOPN=70 timeout 600 python tools/profile_step.py > gpurun_out/r2z_profile_full.log 2>&1
echo "== profile exit $?"; grep -v "Warn\|warn" gpurun_out/r2z_profile_full.log | head -5
