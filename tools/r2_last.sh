#!/bin/bash
mkdir -p gpurun_out
timeout 150 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
echo "== bench exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2y_bench.json'))
print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],2), "roofline", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), round(d["roofline"]["gemm_ms_per_step"],2), d["roofline"]["traffic"])
print(d["roofline"]["ms_by_epilogue"])
PY
grep -v "Warn\|warn\|run_backward" gpurun_out/r2y_bench.err | tail -3
