#!/bin/bash
# round 2, call 23 (the last 13 GPU-minutes): knob sweeps of the HBM-bound kernels, the full GPU suite, a profile and a
# short bench; everything is also written under gpurun_out/ as it goes
mkdir -p gpurun_out
timeout 240 python scripts/bench_norms.py "" "gn_reduce=0" "ln_warps=8" "temporal_warps=8" "gn_apply_blocks=592" \
  "ln_warps=8,temporal_warps=8,gn_apply_blocks=592" 2>&1 | tee gpurun_out/bench_norms_r02.txt | tail -45
echo "== full GPU suite"
EMOTE_PARITY_LOG=gpurun_out/parity23_fp16.log timeout 420 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest23.txt 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/pytest23.txt
echo "== profile"
timeout 120 python scripts/profile_unet.py > gpurun_out/profile_unet_r02c.txt 2>&1; head -16 gpurun_out/profile_unet_r02c.txt
echo "== bench"
timeout 200 python bench.py --no-variants --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err
echo "bench rc=$?"; tail -2 gpurun_out/bench23.err; cut -c1-700 gpurun_out/bench23.json
