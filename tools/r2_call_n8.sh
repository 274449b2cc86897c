#!/bin/bash
# 8 GPUs: config 3 (train step, global batch 64), config 4 (zero-shot sharded), config 5 (loss sweep)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader > gpurun_out/r2n8_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 8 --steps 8 --warmup 4 > gpurun_out/r2n8_bench.json 2> gpurun_out/r2n8_bench.err
echo "== bench n8 exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2n8_bench.json'))
print("n8", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), d["clocks"])
PY
grep -v "Warn\|warn\|run_backward" gpurun_out/r2n8_bench.err | tail -3
timeout 600 $TR --master-port 29552 tools/bench_zero_shot.py --volumes 1024 > gpurun_out/r2n8_zero_shot.log 2>&1
echo "== zero-shot n8 exit $?"; grep '^{' gpurun_out/r2n8_zero_shot.log || tail -5 gpurun_out/r2n8_zero_shot.log
timeout 600 $TR --master-port 29553 tools/bench_zero_shot.py --volumes 1024 --batch 4 > gpurun_out/r2n8_zero_shot_b4.log 2>&1
echo "== zero-shot n8 batch 4 exit $?"; grep '^{' gpurun_out/r2n8_zero_shot_b4.log || tail -5 gpurun_out/r2n8_zero_shot_b4.log
timeout 600 $TR --master-port 29554 tools/bench_loss_sweep.py > gpurun_out/r2n8_loss_sweep.log 2>&1
echo "== loss sweep n8 exit $?"; grep '^{' gpurun_out/r2n8_loss_sweep.log || tail -5 gpurun_out/r2n8_loss_sweep.log
