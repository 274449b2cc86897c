#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus 2 --steps 8 --warmup 4 "$@" > gpurun_out/r2m_$tag.json 2> gpurun_out/r2m_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2m_$tag.json'))
    print("$tag", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), d["clocks"]["sm_mhz"], d["config"]["ddp"])
except Exception as e:
    print("$tag failed", e)
PY
grep -v "Warn\|warn\|run_backward" gpurun_out/r2m_$tag.err | grep -i "error\|Traceback" -A3 | head -12
}
run ignore --ddp ignore-unused
run static --ddp static-graph
run find --ddp find-unused
DDP_MODE=ignore-unused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2m_ddp_timeline.log 2>&1
echo "== ddp timeline exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2m_ddp_timeline.log | grep "GPU activities\|  stream \|gap \|multi_\|u32" | head -30 | cut -c1-170
grep "AllReduce_Sum_f32" gpurun_out/r2m_ddp_timeline.log | head -22 | cut -c1-50
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2m_n1.json 2> gpurun_out/r2m_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2m_n1.json'))
print("n1", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"])
PY
