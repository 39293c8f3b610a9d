#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity7_fp16.log gpurun_out/parity7_bf16.log
EMOTE_PARITY_LOG=gpurun_out/parity7_fp16.log timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest7_fp16.txt 2>&1
echo "fp16 pytest rc=$?"; tail -25 gpurun_out/pytest7_fp16.txt
EMOTE_OPERAND=bf16 EMOTE_PARITY_LOG=gpurun_out/parity7_bf16.log timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest7_bf16.txt 2>&1
echo "bf16 pytest rc=$?"; tail -12 gpurun_out/pytest7_bf16.txt
timeout 300 python scripts/profile_unet.py > gpurun_out/profile_unet_r02.txt 2>&1; grep -E "UNet call|VAE decode|emote_attention_wide" gpurun_out/profile_unet_r02.txt
bash scripts/gpu_sanitize.sh
