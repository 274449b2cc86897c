#!/bin/bash
# 2 GPUs: DDP contention experiments (NCCL CTA cap, SMs reserved from the persistent GEMM) + new e2e feed
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus 2 --steps 8 --warmup 4 > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2k_$tag.json'))
    print("$tag", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), "fp32-host", round(d["e2e_fp32_host"]["ms_per_step"],2), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag failed", e)
PY
}
run base A=1
run ncclcta8 NCCL_MAX_CTAS=8
run ncclcta8_res8 NCCL_MAX_CTAS=8 CTK_GEMM_RESERVE_SMS=8
run ncclcta4_res4 NCCL_MAX_CTAS=4 CTK_GEMM_RESERVE_SMS=4
run res16 CTK_GEMM_RESERVE_SMS=16
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2k_n1.json 2> gpurun_out/r2k_n1.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2k_n1.json'))
print("n1", round(d["value"],1), "vol/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["ms_per_step"],2), d["clocks"]["sm_mhz"])
PY
