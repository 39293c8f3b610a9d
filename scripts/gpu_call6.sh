#!/bin/bash
mkdir -p gpurun_out
timeout 180 python scripts/bench_attn.py > gpurun_out/attn_v2c.txt 2>&1; grep "tc:\|diff" gpurun_out/attn_v2c.txt
EMOTE_ATTN_TC=1 timeout 180 python scripts/bench_attn.py 2>&1 | grep "tc:"
timeout 300 python scripts/bench_conv.py > gpurun_out/conv_r02.txt 2>&1; cat gpurun_out/conv_r02.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -c 1 -f -o gpurun_out/r02_attn_tc2 python scripts/bench_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/unet_traffic_r02.csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python scripts/one_unet_call.py > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 800 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 --no-variants --no-parity --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*_r02.csv 2>/dev/null | tail -5
