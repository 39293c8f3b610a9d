#!/bin/bash
# round 2, call 9 (2 GPUs): bench.py at N=2 with the variants (unit-sharded long form vs single GPU)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench9_n2.json 2> gpurun_out/bench9_n2.err
echo "bench N=2 rc=$?"
tail -5 gpurun_out/bench9_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench9_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["ms_per_step"])
print(json.dumps(d.get("variants"), indent=1))
PY
