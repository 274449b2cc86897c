#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_properties_gpu.py tests/test_model_gpu.py -m gpu -q --no-header -p no:cacheprovider -k "vq or tiny_ctvit_forward" > gpurun_out/r2d_vq.log 2>&1
echo "== vq tests exit $?"; tail -n 5 gpurun_out/r2d_vq.log
timeout 600 python tools/e2e_timeline.py --mode fp32 > gpurun_out/r2d_timeline_fp32.log 2>&1
echo "== timeline fp32 exit $?"; tail -n 60 gpurun_out/r2d_timeline_fp32.log
timeout 600 python tools/e2e_timeline.py --mode resident > gpurun_out/r2d_timeline_resident.log 2>&1
echo "== timeline resident exit $?"; tail -n 30 gpurun_out/r2d_timeline_resident.log
TOPN=60 timeout 600 python tools/profile_step.py > gpurun_out/r2d_profile_full.log 2>&1
echo "== profile full exit $?"; head -n 70 gpurun_out/r2d_profile_full.log
TEXT_STUB=1 TOPN=10 timeout 600 python tools/profile_step.py > gpurun_out/r2d_profile_stub.log 2>&1
echo "== profile stub exit $?"; head -n 14 gpurun_out/r2d_profile_stub.log; tail -n 3 gpurun_out/r2d_profile_stub.log
