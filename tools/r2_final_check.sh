#!/bin/bash
# last GPU call of round 2: the final tree's bench line (short) and smoke(), inside what is left of the GPU budget
mkdir -p gpurun_out
timeout 30 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-torch-eager > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "== bench exit $?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r2z_bench.json'))
    print(round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["traffic_note"][:120])
except Exception as e:
    print("no bench line:", e)
PY
timeout 14 python __graft_entry__.py --smoke > gpurun_out/r2z_smoke.log 2>&1
echo "== smoke exit $?"; tail -2 gpurun_out/r2z_smoke.log
