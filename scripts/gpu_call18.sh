#!/bin/bash
# round 2, call 18: GEGLU double staging (mode 1) + cleanup check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
EMOTE_OPERAND=bf16 timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv or geglu" 2>&1 | tail -3
echo "== gemm";      timeout 300 python scripts/bench_gemm.py 0 5 6 1 2 7 2>&1 | grep "TF/s"
echo "== UNet call"; timeout 300 python scripts/graph_unet.py 2>&1 | tail -4
