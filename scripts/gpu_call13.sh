#!/bin/bash
# round 2, call 13: which epilogue flavour is best for the fp32-residual GEMMs now (registers / one staging buffer / two)
mkdir -p gpurun_out
S="1 2 3 7 8 18"
echo "== default";            python scripts/bench_gemm.py $S 2>&1 | grep "TF/s"
echo "== TMA_STORE=2 (register residual, direct stores)"; TMA_STORE=2 python scripts/bench_gemm.py $S 2>&1 | grep "TF/s"
echo "== TMA_STORE=4 (two staging buffers, BN=128)";      TMA_STORE=4 python scripts/bench_gemm.py $S 2>&1 | grep "TF/s"
echo "== TMA_STORE=3";        TMA_STORE=3 python scripts/bench_gemm.py $S 2>&1 | grep "TF/s"
echo "== PAIR=1";             PAIR=1 python scripts/bench_gemm.py 2 3 8 18 2>&1 | grep "TF/s"
echo "== PAIR=2";             PAIR=2 python scripts/bench_gemm.py 2 3 7 8 18 2>&1 | grep "TF/s"
echo "== BLOCK_N=128";        BLOCK_N=128 python scripts/bench_gemm.py 2 3 7 8 18 2>&1 | grep "TF/s"
