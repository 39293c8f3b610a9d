"""Dev tool: time the tcgen05 vs mma.sync attention kernels on the UNet's self-attention shapes."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import ops
BF16 = torch.bfloat16
for (batch, heads, d, n) in [(32, 8, 40, 4096), (32, 8, 80, 1024)]:
    C = heads * d
    qkv = torch.randn(batch, n, 3 * C, device="cuda").to(BF16)
    out = torch.empty(batch, n, C, device="cuda", dtype=BF16)
    res = {}
    for impl in ("mma", "tc"):
        fn = lambda: ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], out, batch=batch, heads=heads,
                                   head_dim=d, nq=n, n0=n, q_strides=(n * 3 * C, 3 * C), kv0_strides=(n * 3 * C, 3 * C),
                                   o_strides=(n * C, C), scale=d ** -0.5, impl=impl)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[impl] = out.clone()
        print(f"batch={batch} heads={heads} d={d} n={n} {impl}: {ms:.3f} ms  {4.0*batch*heads*d*n*n/ms/1e9:.1f} TF/s")
    print("  tc vs mma rel diff", ((res['tc'].float() - res['mma'].float()).norm() / res['mma'].float().norm()).item())
