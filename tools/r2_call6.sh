#!/bin/bash
# 2 GPUs: NCCL parity test, weak-scaling bench (B=8/GPU), config-3 strong-scaling point (global batch 64 -> B=32/GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2f_gpus.txt
timeout 900 python -m pytest tests/test_nccl_parity_gpu.py -m gpu -q --no-header -p no:cacheprovider -s > gpurun_out/r2f_nccl_parity.log 2>&1
echo "== nccl parity exit $?"; tail -n 12 gpurun_out/r2f_nccl_parity.log
run_bench() { # name, extra args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 4 "$@" > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
  echo "== $name exit $?"; cat gpurun_out/r2f_$name.json | cut -c1-1500; tail -3 gpurun_out/r2f_$name.err
}
run_bench bench_n2
run_bench bench_n2_b32 --batch-per-gpu 32
