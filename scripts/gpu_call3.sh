#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity3_fp16.log
EMOTE_PARITY_LOG=gpurun_out/parity3_fp16.log timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest3_fp16.txt 2>&1
echo "fp16 pytest rc=$?"
tail -60 gpurun_out/pytest3_fp16.txt
grep -E "audio|smoke|info" gpurun_out/parity3_fp16.log
