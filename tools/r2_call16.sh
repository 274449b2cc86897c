#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus 2 --steps 8 --warmup 4 > gpurun_out/r2p_$tag.json 2> gpurun_out/r2p_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2p_$tag.json'))
    print("$tag", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), d["clocks"]["sm_mhz"], d["config"]["ddp"])
except Exception as e:
    print("$tag failed", e)
PY
grep -v "Warn\|warn\|run_backward" gpurun_out/r2p_$tag.err | grep -i "error\|Traceback" -A3 | head -12
}
run prio A=1
run noprio CTK_TEXT_STREAM_PRIORITY=0
DDP_MODE=ignore-unused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2p_ddp_timeline.log 2>&1
echo "== ddp timeline exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2p_ddp_timeline.log | grep "GPU activities"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2p_ddp_timeline.log | grep -A48 "1 ms bins" | awk 'NR==1 || ($2+0)>=12' | head -40
for m in A=1 CTK_TEXT_STREAM_PRIORITY=0; do
env $m CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2p_n1.json 2> gpurun_out/r2p_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2p_n1.json'))
print("n1 $m", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"])
PY
done
