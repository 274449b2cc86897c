#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nccl_parity_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/r2j_nccl_parity.log 2>&1
echo "== nccl parity exit $?"; grep -v "Warn\|warn\|run_backward\|^$" gpurun_out/r2j_nccl_parity.log | tail -n 4 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 4 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
echo "== bench n2 exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench_n2.json'))
print("n2", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), d["clocks"])
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
echo "== bench n1 exit $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench_n1.json'))
print("n1", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), d["clocks"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2j_ddp_timeline.log 2>&1
echo "== ddp timeline exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2j_ddp_timeline.log | grep "GPU activities\|  stream \|gap \|Broadcast\|AllGather\|multi_\|patch_norm\|u32" | head -60 | cut -c1-170
