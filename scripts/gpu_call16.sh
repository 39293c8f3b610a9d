#!/bin/bash
# round 2, call 16: mode-2 epilogue with a store warp and column halves (single-CTA kernel); per-box bulk groups in the pair kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv" 2>&1 | tail -4
echo "== default";      timeout 300 python scripts/bench_gemm.py 1 2 8 3 7 14 16 2>&1 | grep "TF/s"
echo "== TMA_STORE=3 (never the two-buffer variant)";  TMA_STORE=3 timeout 300 python scripts/bench_gemm.py 1 16 2>&1 | grep "TF/s"
echo "== PAIR=2 (single-CTA kernel for the K >= 1024 shapes)"; PAIR=2 timeout 300 python scripts/bench_gemm.py 3 7 18 2>&1 | grep "TF/s"
echo "== UNet call"; timeout 300 python scripts/graph_unet.py 2>&1 | tail -4
