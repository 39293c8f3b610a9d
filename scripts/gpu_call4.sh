#!/bin/bash
mkdir -p gpurun_out
timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2.txt 2>&1; echo "attn v2 rc=$?"; cat gpurun_out/attn_v2.txt | tail -20
EMOTE_ATTN_TC=1 timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v1.txt 2>&1; echo "attn v1 rc=$?"; cat gpurun_out/attn_v1.txt | tail -12
EMOTE_ATTN_EMU=0 timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2_emu0.txt 2>&1; grep "d=40 n=4096 n1=0 tc\|d=80" gpurun_out/attn_v2_emu0.txt
EMOTE_ATTN_EMU=2 timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2_emu2.txt 2>&1; grep "d=40 n=4096 n1=0 tc\|d=80" gpurun_out/attn_v2_emu2.txt
rm -f gpurun_out/parity4_fp16.log
EMOTE_PARITY_LOG=gpurun_out/parity4_fp16.log timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest4_fp16.txt 2>&1
echo "fp16 pytest rc=$?"
tail -25 gpurun_out/pytest4_fp16.txt
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-variants --no-parity > gpurun_out/bench4_fp16.json 2> gpurun_out/bench4_fp16.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench4_fp16.json'))
print(d['value'], d['e2e']['value'], d['clocks'])
print({k:(v['launches'],v['ms']) for k,v in d['kernel_breakdown_one_unet_call'].items() if isinstance(v,dict)})
print(d['roofline_classes'])
PY
