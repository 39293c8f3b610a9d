#!/bin/bash
# round 2, call 25 (last ~45 s of GPU time): the bf16 build over the kernels changed in this session + smoke()
mkdir -p gpurun_out
EMOTE_OPERAND=bf16 timeout 40 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x \
  -k "gn_colstats or group_norm or tuning_knobs or layer_norm or temporal_attention or gemm_fused_gn" 2>&1 | tail -2 | tee gpurun_out/pytest25_bf16.txt
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke25.txt
