#!/bin/bash
# First GPU call of round 2: validate what was written after round 1's GPU budget ran out, in isolation
# (one process per file, each under its own timeout so that a hung tcgen05 epilogue cannot hang the box),
# then measure it.  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Everything lands in gpurun_out/r2_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu.txt 2>&1

run() {   # name, timeout seconds, command...
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > gpurun_out/r2_$name.log 2>&1
  echo "== $name exit $?"; tail -n 12 gpurun_out/r2_$name.log
}

# 1. the new GEMM epilogues and the text tower (GELU / GELU_BWD), then the tensor-core loss path
export CTK_TEST_UNVERIFIED=1
run gelu_epilogues 300 python -m pytest tests/test_text_tower_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "gelu"
run text_tower 600 python -m pytest tests/test_text_tower_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "not gelu"
run clip_loss_tc 900 python -m pytest tests/test_clip_loss_tc_gpu.py -m gpu -q --no-header -p no:cacheprovider
run volume_prep 300 python -m pytest tests/test_volume_prep_gpu.py -m gpu -q --no-header -p no:cacheprovider
run ctvit3d 600 python -m pytest tests/test_ctvit3d_gpu.py -m gpu -q --no-header -p no:cacheprovider
run zero_shot 300 python -m pytest tests/test_zero_shot_gpu.py -m gpu -q --no-header -p no:cacheprovider
unset CTK_TEST_UNVERIFIED

# 2. the validated suite must still be green with the rebuilt library
run suite 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider

# 3. measurements: stock vs libctk text tower; loss sweep with and without the tensor-core path
run bench_hf 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --stored-host --half-host
grep -h '^{' gpurun_out/r2_bench_hf.log > gpurun_out/r2_bench_hf.json
run bench_ctk 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --text-tower ctk
grep -h '^{' gpurun_out/r2_bench_ctk.log > gpurun_out/r2_bench_ctk.json
CTK_CLIP_LOSS_TC=0 run configs_simt 600 python tools/bench_configs.py
CTK_CLIP_LOSS_TC=1 run configs_tc 600 python tools/bench_configs.py
run zero_shot_bench 600 python tools/bench_zero_shot.py --volumes 32
run zero_shot_bench_b8 600 python tools/bench_zero_shot.py --volumes 64 --batch 8
run config1 300 python tools/bench_config1.py --gpu
run h2d_probe 300 python tools/h2d_probe.py
CTK_EVAL_GRAPHS=1 run configs_evalgraph 600 python tools/bench_configs.py
CTK_EVAL_GRAPHS=1 run zero_shot_bench_graph 600 python tools/bench_zero_shot.py --volumes 32
