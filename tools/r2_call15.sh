#!/bin/bash
mkdir -p gpurun_out
DDP_MODE=ignore-unused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2o_ddp_timeline.log 2>&1
echo "== ddp timeline exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2o_ddp_timeline.log | grep -A48 "1 ms bins" | head -52
