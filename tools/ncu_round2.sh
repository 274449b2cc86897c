#!/bin/bash
# ncu evidence for round 2 (run under gpurun, one GPU).  A plain run of the same command goes first.
#   1. launch list of the device-timed region of `bench.py --steps 1` (every kernel, graph nodes included)
#   2. `--set full` of the kernels that changed this round + the GEMM family at the bench batch
set -o pipefail
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-eager"
$CMD > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_ncu_plain.err; exit 1; }
cut -c1-300 gpurun_out/r2_ncu_plain.json
CTK_BENCH_PROFILER_RANGE=1 timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --graph-profiling node \
    --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list exit $?"; grep -c "gpu__time_duration" gpurun_out/r2_launches.csv
# NOTE: .ncu-rep files with --import-source are ~80 MB each and gpurun refuses to copy back a gpurun_out/ above 64 MiB
# (round 2 lost a 35-minute capture to that): reports stay in /tmp on the box, only the raw-page CSVs come back.
full() {  # name, ncu selection args...
  local name=$1; shift
  CTK_BENCH_PROFILER_RANGE=1 timeout 1200 ncu --profile-from-start off --set full --clock-control none \
      --graph-profiling node "$@" -o /tmp/r2_prof_$name -f $CMD > gpurun_out/r2_ncu_$name.log 2>&1
  echo "full capture $name exit $?"
  ncu -i /tmp/r2_prof_$name.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$name.raw.csv 2>/dev/null
  ls -la gpurun_out/r2_ncu_$name.raw.csv
}
full gemm_fwd -k regex:gemm_kernel -c 44
full gemm_bwd -k regex:gemm_kernel -s 100 -c 60
full new_kernels -k regex:"mha_|vq_select|latent_|bert_embed" -c 44
du -sh gpurun_out
