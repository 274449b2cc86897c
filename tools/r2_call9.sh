#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_text_tower_gpu.py tests/test_ctvit3d_gpu.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/r2i_mha.log 2>&1
echo "== mha + ctvit3d tests exit $?"; grep -v "Warn\|warn\|run_backward\|^$" gpurun_out/r2i_mha.log | tail -n 15
timeout 300 python tools/bench_mha.py > gpurun_out/r2i_bench_mha.log 2>&1
echo "== bench mha exit $?"; grep '^{' gpurun_out/r2i_bench_mha.log || tail -5 gpurun_out/r2i_bench_mha.log
for mode in encoder_side serial; do
  CTK_TOWER_STREAMS=$mode timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2i_bench_$mode.json 2> gpurun_out/r2i_bench_$mode.err
  echo "== bench $mode exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2i_bench_$mode.json'))
print("$mode", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), "sync-read", round(d["value_sync_loss_read"]["ms_per_step"],2), "roofline", round(d["roofline"]["frac"],3), round(d["roofline"]["gemm_ms_per_step"],2), d["clocks"])
print(d["roofline"]["ms_by_epilogue"])
PY
done
