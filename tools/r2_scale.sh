#!/bin/bash
# the driver's scaling run: N = 1, 2, 4, 8 back to back on one 8-GPU box (weak scaling, 8 volumes / GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader > gpurun_out/r2s_gpus.txt
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 8 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2s_n$n.json 2> gpurun_out/r2s_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 8 --warmup 4 > gpurun_out/r2s_n$n.json 2> gpurun_out/r2s_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2s_n$n.json'))
    print("n=$n", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), "sync-read", round(d["value_sync_loss_read"]["ms_per_step"],2), d["clocks"]["sm_mhz"], d["config"]["ddp"])
except Exception as e:
    print("n=$n failed", e)
PY
  grep -v "Warn\|warn\|run_backward" gpurun_out/r2s_n$n.err | grep -i "error\|Traceback" -A3 | head -8
done
