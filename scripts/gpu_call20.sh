#!/bin/bash
# round 2, call 20: weight-stationary kernel vs the streaming kernel on the small-K 16-bit-output shapes (policy knob)
mkdir -p gpurun_out
for b in 2 1 0; do
  echo "== EMOTE_BRES=$b"; EMOTE_BRES=$b timeout 300 python scripts/bench_gemm.py 0 4 17 2>&1 | grep "TF/s"
  EMOTE_BRES=$b timeout 300 python scripts/graph_unet.py 2>&1 | tail -3
done
