#!/bin/bash
mkdir -p gpurun_out
DDP_MODE=ignore-unused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2n_ddp_timeline.log 2>&1
echo "== ddp timeline exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2n_ddp_timeline.log | grep -A60 "1 ms bins" | head -64
(CUDA_VISIBLE_DEVICES=1 timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_text_tower_gpu.py tests/test_head_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/r2n_tests.log 2>&1; echo "== tests exit $?"; tail -n 3 gpurun_out/r2n_tests.log | cut -c1-200)
