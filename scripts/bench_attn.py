"""Dev tool: time the tcgen05 attention kernel (EMOTE_ATTN_EMU = exp split) against the mma.sync kernel on the UNet's
self-attention shapes."""
import os
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import _lib, ops
OP16 = ops.OP16
print(f"operand={_lib.OPERAND} EMOTE_ATTN_EMU={os.environ.get('EMOTE_ATTN_EMU', 'default')}")
shapes = [(32, 8, 40, 4096, 0), (32, 8, 80, 1024, 0), (32, 8, 160, 256, 0), (32, 8, 40, 4096, 4096), (1, 12, 64, 499, 0)]
for (batch, heads, d, n, n1) in shapes:
    C = heads * d
    qkv = torch.randn(batch, n, 3 * C, device="cuda").to(OP16)
    bank = torch.randn(2, max(n1, 1), 2 * C, device="cuda").to(OP16)
    out = torch.empty(batch, n, C, device="cuda", dtype=OP16)
    res = {}
    impls = ("mma", "tc") if ops._lib.load().emote_attention_tc_supported(d) else ("mma",)
    for impl in impls:
        kw = {}
        if n1:
            kw = dict(k1=bank[..., :C], v1=bank[..., C:], n1=n1, kv1_strides=(n1 * 2 * C, 2 * C), kv1_batch_div=batch // 2,
                      kv1_first_batch=batch // 2)
        fn = lambda: ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], out, batch=batch, heads=heads,
                                   head_dim=d, nq=n, n0=n, q_strides=(n * 3 * C, 3 * C), kv0_strides=(n * 3 * C, 3 * C),
                                   o_strides=(n * C, C), scale=d ** -0.5, impl=impl, **kw)
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
        flops = 4.0 * batch * heads * d * n * (n + 0.5 * n1)
        print(f"batch={batch} heads={heads} d={d} n={n} n1={n1} {impl}: {ms:.3f} ms  {flops/ms/1e9:.1f} TF/s")
    if "tc" in res:
        print("  tc vs mma rel diff", ((res['tc'].float() - res['mma'].float()).norm() / res['mma'].float().norm()).item())
