#!/bin/bash
# round 2, call 26: closing verification of the final tree — full GPU suite (fp16 build)
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest26.txt
