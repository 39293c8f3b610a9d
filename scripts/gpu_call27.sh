#!/bin/bash
# round 2, call 27 (last GPU seconds): ncu --set full of the two small attention kernels
mkdir -p gpurun_out
timeout 70 ncu --set full --clock-control none -k regex:"temporal_attn|short_kv" -c 4 --csv --page raw \
  --log-file gpurun_out/small_attn_ncu.csv python scripts/ncu_small_attn.py > gpurun_out/ncu27.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/small_attn_ncu.csv; tail -2 gpurun_out/ncu27.log
