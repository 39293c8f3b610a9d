#!/bin/bash
# round 2, call 22: faster statistics reduce + sanitizer passes over the kernels changed since the full sanitizer run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -3
echo "== UNet call"; timeout 300 python scripts/graph_unet.py 2>&1 | tail -3
timeout 300 python scripts/profile_unet.py 2>&1 | grep -E "UNet call|gn_colstats_reduce|gn_apply|emote_gemm_bf16 "
export EMOTE_PARITY_LOG=
CS="compute-sanitizer --error-exitcode 1 --print-limit 20"
K="gemm or conv3x3 or group_norm or video_grid or vae_postprocess or geglu"
timeout 900 $CS --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$K" > gpurun_out/racecheck_r02b.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_r02b.log | tail -5
timeout 900 $CS --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$K" > gpurun_out/memcheck_r02b.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02b.log | tail -4
timeout 600 $CS --tool synccheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm_plain or conv3x3_implicit or gemm_fused" > gpurun_out/synccheck_r02b.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/synccheck_r02b.log | tail -3
