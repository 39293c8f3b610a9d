"""Dev tool for ncu: one launch each of the two small attention kernels at the 64x64 level of config #2 —
temporal_attn_kernel<48,16,4> (2 samples x 4096 pixels x 8 heads over 16 frames) and short_kv_attn_kernel<48,5>
(32 images x 8 heads, 4096 queries against 77 context keys).  Usage:
  ncu --set full --clock-control none -k regex:"temporal_attn|short_kv" --csv --page raw \
      --log-file gpurun_out/small_attn_ncu.csv python scripts/ncu_small_attn.py"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import ops  # noqa: E402

dev = torch.device("cuda")
OP16 = ops.OP16
B, F_, HW, heads, d = 2, 16, 4096, 8, 40
C = heads * d
qkv = torch.randn(B * F_ * HW, 3 * C, device=dev).to(OP16)
q = torch.randn(B * F_ * HW, C, device=dev).to(OP16)
kv = torch.randn(2 * 77, 2 * C, device=dev).to(OP16)
out = torch.empty(B * F_ * HW, C, dtype=OP16, device=dev)
for _ in range(2):
    ops.temporal_attention(qkv, B, F_, HW, heads, d)
    ops.attention(q, kv[:, :C], kv[:, C:], out, batch=B * F_, heads=heads, head_dim=d, nq=HW, n0=77,
                  q_strides=(HW * C, C), kv0_strides=(77 * 2 * C, 2 * C), o_strides=(HW * C, C), scale=d ** -0.5,
                  kv0_batch_div=F_)
torch.cuda.synchronize()
