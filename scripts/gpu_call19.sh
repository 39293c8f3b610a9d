#!/bin/bash
# round 2, call 19: weight-stationary kernel with store warp + column halves (linear outputs); full fp16 suite with parity log
mkdir -p gpurun_out
echo "== gemm";      timeout 300 python scripts/bench_gemm.py 4 17 0 19 2>&1 | grep "TF/s"
echo "== PAIR=3 (no weight-stationary kernel)"; PAIR=3 timeout 300 python scripts/bench_gemm.py 4 17 2>&1 | grep "TF/s"
echo "== UNet call"; timeout 300 python scripts/graph_unet.py 2>&1 | tail -4
rm -f gpurun_out/parity19_fp16.log
EMOTE_PARITY_LOG=gpurun_out/parity19_fp16.log timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest19_fp16.txt 2>&1
echo "fp16 pytest rc=$?"; tail -12 gpurun_out/pytest19_fp16.txt
