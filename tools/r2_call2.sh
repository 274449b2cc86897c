#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -rs > gpurun_out/r2b_suite.log 2>&1
echo "== suite exit $?"; tail -n 40 gpurun_out/r2b_suite.log
timeout 600 python tools/e2e_probe.py > gpurun_out/r2b_e2e_probe.log 2>&1
echo "== e2e_probe exit $?"; grep '^{' gpurun_out/r2b_e2e_probe.log || tail -20 gpurun_out/r2b_e2e_probe.log
