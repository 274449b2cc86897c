#!/bin/bash
# 4 GPUs: config 3 weak (B=8/GPU) and strong (global batch 64 -> B=16/GPU); then 2 of them: B=32/GPU
mkdir -p gpurun_out
run() { tag=$1; n=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $n --steps 6 --warmup 4 "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2l_$tag.json'))
    print("$tag", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), d["clocks"]["sm_mhz"], d["config"]["global_batch"])
except Exception as e:
    print("$tag failed", e)
PY
grep -i "out of memory" gpurun_out/r2l_$tag.err | head -2 | cut -c1-300
}
run n4_b8 4
run n4_b16 4 --batch-per-gpu 16
CUDA_VISIBLE_DEVICES=0,1 run n2_b32 2 --batch-per-gpu 32
