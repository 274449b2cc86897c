#!/bin/bash
mkdir -p gpurun_out
for m in stored fp32; do
E2E=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2r_ddp_e2e_$m.log 2>&1
echo "== ddp e2e timeline $m exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2r_ddp_e2e_$m.log | grep "GPU activities\|GB/s\|  stream \|gap \|patch_norm\|multi_adam" | head -40 | cut -c1-170
done
