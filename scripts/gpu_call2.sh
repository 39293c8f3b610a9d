#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity2_fp16.log
EMOTE_PARITY_LOG=gpurun_out/parity2_fp16.log timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest2_fp16.txt 2>&1
echo "fp16 pytest rc=$?"
tail -40 gpurun_out/pytest2_fp16.txt
timeout 1200 python bench.py --steps 3 --warmup 2 > gpurun_out/bench2_fp16.json 2> gpurun_out/bench2_fp16.err
echo "bench fp16 rc=$?"; tail -5 gpurun_out/bench2_fp16.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2_fp16.json'))
for k in ('value','e2e','parity','cpu_baseline','variants','roofline','roofline_classes','clocks'):
    print(k, json.dumps(d.get(k))[:1500])
PY
