#!/bin/bash
# 2-GPU box: NCCL multi-GPU test + statistics / VAE tests after the slot change + bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_kernels_gpu.py tests/test_pipeline_gpu.py tests/test_unet_gpu.py -m gpu -q -x --durations=5 > gpurun_out/pytest8_n2.txt 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest8_n2.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench8_n2.json 2> gpurun_out/bench8_n2.err
echo "bench N=2 rc=$?"; tail -5 gpurun_out/bench8_n2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench8_n2.json'))
print(d['value'], d['e2e'], d['n_gpus'])
print(json.dumps(d['variants'])[:3000])
PY
