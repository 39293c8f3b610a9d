#!/bin/bash
# round 2, final measurement call (1 GPU): DRAM traffic capture, fp16 suite + bench, bf16 suite + bench, ncu launch list
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/unet_traffic_r02.csv \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python scripts/one_unet_call.py > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"
python scripts/summarize_traffic.py gpurun_out/unet_traffic_r02.csv profiles/r02_unet_call_dram_traffic.json && cp profiles/r02_unet_call_dram_traffic.json gpurun_out/
rm -f gpurun_out/parity21_fp16.log gpurun_out/parity21_bf16.log
EMOTE_PARITY_LOG=gpurun_out/parity21_fp16.log timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest21_fp16.txt 2>&1
echo "fp16 pytest rc=$?"; tail -6 gpurun_out/pytest21_fp16.txt
timeout 1500 python bench.py > gpurun_out/bench21_fp16.json 2> gpurun_out/bench21_fp16.err
echo "bench fp16 rc=$?"; tail -3 gpurun_out/bench21_fp16.err
EMOTE_OPERAND=bf16 EMOTE_PARITY_LOG=gpurun_out/parity21_bf16.log timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest21_bf16.txt 2>&1
echo "bf16 pytest rc=$?"; tail -6 gpurun_out/pytest21_bf16.txt
EMOTE_OPERAND=bf16 timeout 900 python bench.py --no-variants --no-cpu-baseline > gpurun_out/bench21_bf16.json 2> gpurun_out/bench21_bf16.err
echo "bench bf16 rc=$?"
timeout 300 python scripts/profile_unet.py > gpurun_out/profile_unet_r02.txt 2>&1; head -20 gpurun_out/profile_unet_r02.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 800 --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 1 --warmup 1 --no-variants --no-parity --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench21_fp16.json", "gpurun_out/bench21_bf16.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"], d["parity"], d["roofline"]["achieved"], d["roofline"]["frac"])
        print({k: (v.get("value"), v.get("ms_per_ddim_step")) for k, v in (d.get("variants") or {}).items()})
        print(d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
