#!/bin/bash
# round 2, call 24 (last GPU minutes): bench with the 3-pass roofline median, sanitizer passes over the kernels changed in this
# session (statistics fold, gn_apply prologue, knob-selected geometries), ncu duration + DRAM traffic capture of one UNet call
mkdir -p gpurun_out
timeout 150 python bench.py --no-variants --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err
echo "bench rc=$?"; tail -2 gpurun_out/bench24.err; cut -c1-300 gpurun_out/bench24.json
export EMOTE_PARITY_LOG=
CS="compute-sanitizer --error-exitcode 1 --print-limit 20"
K="gn_colstats or group_norm or tuning_knobs or layer_norm or temporal_attention or gemm_fused_gn"
timeout 120 $CS --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$K" > gpurun_out/memcheck_r02c.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r02c.log | tail -3
timeout 120 $CS --tool racecheck python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "$K" > gpurun_out/racecheck_r02c.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_r02c.log | tail -3
timeout 200 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/unet_traffic_r02c.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python scripts/one_unet_call.py > gpurun_out/ncu_traffic24.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/unet_traffic_r02c.csv
python scripts/summarize_traffic.py gpurun_out/unet_traffic_r02c.csv gpurun_out/r02c_unet_call_dram_traffic.json && head -c 600 gpurun_out/r02c_unet_call_dram_traffic.json
