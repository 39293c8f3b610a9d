#!/bin/bash
# GPU call: full GPU suite in both operand modes with parity logging, then a short bench in both modes
mkdir -p gpurun_out
rm -f gpurun_out/parity_fp16.log gpurun_out/parity_bf16.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
EMOTE_PARITY_LOG=gpurun_out/parity_fp16.log timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_fp16.txt 2>&1
echo "fp16 pytest rc=$?"
tail -30 gpurun_out/pytest_fp16.txt
EMOTE_OPERAND=bf16 EMOTE_PARITY_LOG=gpurun_out/parity_bf16.log timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_bf16.txt 2>&1
echo "bf16 pytest rc=$?"
tail -15 gpurun_out/pytest_bf16.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-variants > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
echo "bench fp16 rc=$?"; cat gpurun_out/bench_fp16.json | cut -c1-600
EMOTE_OPERAND=bf16 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-variants > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
echo "bench bf16 rc=$?"; cat gpurun_out/bench_bf16.json | cut -c1-600
