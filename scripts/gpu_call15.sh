#!/bin/bash
# round 2, call 15: residual L2-prefetch distance (1 = old behaviour, 2, 3) + st.shared P-tile stores in the attention kernel
mkdir -p gpurun_out
S="1 2 3 7 8 18 10 11"
for d in 1 2 3; do echo "== EMOTE_RES_PF=$d"; EMOTE_RES_PF=$d python scripts/bench_gemm.py $S 2>&1 | grep "TF/s"; done
echo "== attention"; python scripts/bench_attn.py 2>&1 | grep -v "^operand"
for d in 1 2 3; do echo "== UNet call, EMOTE_RES_PF=$d"; EMOTE_RES_PF=$d timeout 300 python scripts/graph_unet.py 2>&1 | tail -3; done
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
