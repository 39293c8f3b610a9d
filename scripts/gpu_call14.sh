#!/bin/bash
# round 2, call 14: register-residual + staged-store epilogue (tma_store = 5) vs the default, then the fp16 suite
mkdir -p gpurun_out
S="1 2 8 16 14"
echo "== default";      python scripts/bench_gemm.py 1 2 8 2>&1 | grep "TF/s"
echo "== TMA_STORE=5";  TMA_STORE=5 python scripts/bench_gemm.py 1 2 8 2>&1 | grep "TF/s"
echo "== TMA_STORE=5 PAIR=2"; TMA_STORE=5 PAIR=2 python scripts/bench_gemm.py 3 7 2>&1 | grep "TF/s"
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from emote_hack_b200 import ops
torch.manual_seed(0)
for (M, N, K) in [(32768, 640, 640), (131072, 320, 320), (1000, 320, 320), (2048, 1280, 1280)]:
    a = torch.randn(M, K, device="cuda").to(ops.OP16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(ops.OP16)
    bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda")
    o0 = ops.gemm(a, w, bias=bias, residual=res, pair_mode=2)
    o5 = ops.gemm(a, w, bias=bias, residual=res, pair_mode=2, tma_store=5)
    ref = a.float() @ w.float().t() + bias + res
    print(M, N, K, "default vs ref", ((o0 - ref).norm() / ref.norm()).item(), "mode5 vs default max abs", (o5 - o0).abs().max().item())
PY
EMOTE_PARITY_LOG=gpurun_out/parity14_fp16.log timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest14_fp16.txt 2>&1
echo "fp16 pytest rc=$?"; tail -12 gpurun_out/pytest14_fp16.txt
