#!/bin/bash
# dev tool: 2-GPU DDP variants of the bench (run under gpurun --gpus 2)
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        bench.py --gpus 2 --no-cpu-baseline $EXTRA > gpurun_out/ddp_$tag.json 2> gpurun_out/ddp_$tag.err; python -c "
import json
for l in open('gpurun_out/ddp_$tag.json'):
    if l.startswith('{'):
        d=json.loads(l); print('$tag', round(d['value'],1), 'vol/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['ms_per_step'],2))"; }
EXTRA="" run neworder_graphs A=1
EXTRA="" run neworder_nographs CTK_CUDA_GRAPHS=0
EXTRA="" run reforder_graphs CTK_REFERENCE_PARAM_ORDER=1
EXTRA="--bucket-mb 100" run neworder_graphs_b100 A=1
