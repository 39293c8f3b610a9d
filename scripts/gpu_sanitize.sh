#!/bin/bash
# compute-sanitizer passes over the GPU test suite (profiles/r02_sanitizer_*.log are the tracked summaries)
mkdir -p gpurun_out
export EMOTE_PARITY_LOG=
CS="compute-sanitizer --error-exitcode 1 --print-limit 20"
# memcheck: every kernel-level test + the tiny-network / pipeline tests (full-width tests are deselected: too slow under the tool)
timeout 1500 $CS --tool memcheck python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_pipeline_gpu.py tests/test_audio_gpu.py \
  -m "gpu and not slow" -q -x -k "not full_width and not full_size and not config0 and not smoke" > gpurun_out/memcheck_r02.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/memcheck_r02.log | tail -5
# racecheck: shared-memory hazards of the warp-specialised kernels (mbarrier pipelines): GEMM variants, conv, both attention paths
timeout 1500 $CS --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x \
  -k "gemm_plain or gemm_weight_stationary or conv3x3_implicit or flash_attention_tcgen05 or temporal_attention or group_norm or layer_norm" \
  > gpurun_out/racecheck_r02.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_r02.log | tail -8
# synccheck: barrier misuse
timeout 900 $CS --tool synccheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x \
  -k "gemm_plain or conv3x3_implicit or flash_attention_tcgen05" > gpurun_out/synccheck_r02.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/synccheck_r02.log | tail -4
