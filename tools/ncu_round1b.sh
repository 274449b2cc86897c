#!/bin/bash
# round-1 evidence, second pass (run under gpurun, one GPU): launch list of one train step and full
# captures of the tcgen05 GEMM family at the bench batch size (B=8), forward and backward ranges.
set -o pipefail
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
$CMD1 > gpurun_out/plain1.log 2> gpurun_out/plain1.err || { echo "plain run 1 failed"; tail -5 gpurun_out/plain1.err; exit 1; }
cut -c1-300 gpurun_out/plain.log
# one timed step: 3 warm-up steps (~1615 launches each) are skipped
ncu --metrics gpu__time_duration.sum --clock-control none -s 4900 -c 1650 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
# GEMM family, second step of the run: forward launches (patch, q, kv, out, FF1, FF2 ...) then backward (dgrad, wgrad)
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 124 -c 8 -f -o gpurun_out/prof_gemm_fwd \
    $CMD1 > gpurun_out/ncu_gemm_fwd.log 2>&1
echo "gemm fwd capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel" -s 190 -c 8 -f -o gpurun_out/prof_gemm_bwd \
    $CMD1 > gpurun_out/ncu_gemm_bwd.log 2>&1
echo "gemm bwd capture exit $?"
python - <<'PY'
import torch, time
x = torch.empty(1 << 28, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(4): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print(f"H2D pinned 1 GiB x4: {4 * x.numel() * 4 / dt / 1e9:.1f} GB/s")
PY
ls -la gpurun_out | grep -E "prof_gemm|launches"; du -sh gpurun_out
