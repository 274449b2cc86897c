#!/bin/bash
mkdir -p gpurun_out
(CUDA_VISIBLE_DEVICES=0 timeout 1700 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/r2h_suite.log 2>&1; echo "== suite exit $?"; tail -n 6 gpurun_out/r2h_suite.log)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 4 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
echo "== bench n2 exit $?"; cut -c1-600 gpurun_out/r2h_bench_n2.json; grep -v "Warn\|warn\|run_backward" gpurun_out/r2h_bench_n2.err | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/ddp_timeline.py > gpurun_out/r2h_ddp_timeline.log 2>&1
echo "== ddp timeline exit $?"; grep -v "Warning\|warn\|run_backward" gpurun_out/r2h_ddp_timeline.log | grep -v "AllReduce_Sum_f32" | tail -n 40
timeout 900 python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-torch-eager > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err
echo "== bench n1 exit $?"; cut -c1-400 gpurun_out/r2h_bench_n1.json
