#!/bin/bash
mkdir -p gpurun_out
timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2b.txt 2>&1; echo "attn v2b rc=$?"; grep "tc:\|diff" gpurun_out/attn_v2b.txt
EMOTE_ATTN_EMU=0 timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2b_emu0.txt 2>&1; grep "tc:" gpurun_out/attn_v2b_emu0.txt
EMOTE_ATTN_EMU=2 timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2b_emu2.txt 2>&1; grep "tc:" gpurun_out/attn_v2b_emu2.txt
rm -f gpurun_out/parity5_fp16.log
EMOTE_PARITY_LOG=gpurun_out/parity5_fp16.log timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest5_fp16.txt 2>&1
echo "fp16 pytest rc=$?"
tail -30 gpurun_out/pytest5_fp16.txt
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench5_fp16.json 2> gpurun_out/bench5_fp16.err
echo "bench rc=$?"; tail -3 gpurun_out/bench5_fp16.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5_fp16.json'))
print(d['value'], d['e2e']['value'], d['clocks'], d['parity'])
print({k:(v['launches'],v['ms']) for k,v in d['kernel_breakdown_one_unet_call'].items() if isinstance(v,dict)})
print(d['roofline_classes'][0])
print({k:(v.get('ms_per_ddim_step'), v.get('value')) for k,v in d['variants'].items()})
PY
