#!/bin/bash
# round 2, call 17: mode-2 store warp + column halves in the CTA-pair kernel too; mode 3 off by default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
echo "== default";      timeout 300 python scripts/bench_gemm.py 1 2 3 7 8 18 10 11 2>&1 | grep "TF/s"
echo "== UNet call"; timeout 300 python scripts/graph_unet.py 2>&1 | tail -4
timeout 300 python scripts/bench_conv.py 2>&1 | tail -10
