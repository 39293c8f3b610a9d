#!/bin/bash
# round 2, call 12 (N GPUs): bench.py at N = $1 with the variants
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench12_n$N.json 2> gpurun_out/bench12_n$N.err
echo "bench N=$N rc=$?"
grep -v "^\s*$" gpurun_out/bench12_n$N.err | tail -8
python - <<PY
import json
d = json.loads(open("gpurun_out/bench12_n$N.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["ms_per_step"], d.get("clocks"))
print(json.dumps(d.get("variants"), indent=1))
PY
