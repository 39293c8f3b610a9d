"""Dev tool: bitwise run-to-run determinism of individual kernels at the tiny-UNet shapes."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import ops
BF16 = torch.bfloat16
g = torch.Generator(device="cuda"); g.manual_seed(0)

def check(name, fn, reps=5):
    outs = [fn().clone() for _ in range(reps)]
    torch.cuda.synchronize()
    bad = [i for i in range(1, reps) if not torch.equal(outs[0], outs[i])]
    d = max(((outs[0].float() - o.float()).abs().max().item() for o in outs[1:]), default=0)
    print(f"{name:40s} {'DETERMINISTIC' if not bad else 'NONDETERMINISTIC'} maxdiff={d:.3e}")

M, C = 2048, 64
x = torch.randn(M, C, device="cuda", generator=g)
gamma = torch.randn(C, device="cuda", generator=g); beta = torch.randn(C, device="cuda", generator=g)
check("layer_norm", lambda: ops.layer_norm(x, gamma, beta))
a = torch.randn(M, C, device="cuda", generator=g).to(BF16)
w = torch.randn(3 * C, C, device="cuda", generator=g).to(BF16)
check("gemm qkv bf16 out", lambda: ops.gemm(a, w, out_dtype=BF16))
res = torch.randn(M, C, device="cuda", generator=g)
wo = torch.randn(C, C, device="cuda", generator=g).to(BF16); bo = torch.randn(C, device="cuda", generator=g)
check("gemm out-proj +residual", lambda: ops.gemm(a, wo, bias=bo, residual=res))
wg = torch.randn(8 * C, C, device="cuda", generator=g); bg = torch.randn(8 * C, device="cuda", generator=g)
wp, bp = ops.pack_geglu(wg, bg)
check("gemm geglu", lambda: ops.gemm(a, wp, bias=bp, geglu=True, out_dtype=BF16))
qkv = torch.randn(8, 256, 3 * C, device="cuda", generator=g).to(BF16)
def self_attn():
    out = torch.empty(8, 256, C, device="cuda", dtype=BF16)
    ops.attention(qkv[..., :C], qkv[..., C:2*C], qkv[..., 2*C:], out, batch=8, heads=4, head_dim=16, nq=256, n0=256,
                  q_strides=(256*3*C, 3*C), kv0_strides=(256*3*C, 3*C), o_strides=(256*C, C), scale=0.25)
    return out
check("flash self n=256 d=16", self_attn)
q = torch.randn(8, 256, C, device="cuda", generator=g).to(BF16)
kv = torch.randn(2, 7, 2 * C, device="cuda", generator=g).to(BF16)
def cross_attn():
    out = torch.empty(8, 256, C, device="cuda", dtype=BF16)
    ops.attention(q, kv[..., :C], kv[..., C:], out, batch=8, heads=4, head_dim=16, nq=256, n0=7,
                  q_strides=(256*C, C), kv0_strides=(7*2*C, 2*C), o_strides=(256*C, C), scale=0.25, kv0_batch_div=4)
    return out
check("flash cross n0=7", cross_attn)
tq = torch.randn(2 * 4 * 256, 3 * C, device="cuda", generator=g).to(BF16)
check("temporal attn", lambda: ops.temporal_attention(tq, 2, 4, 256, 4, 16))
xs = [torch.randn(2048, 64, device="cuda", generator=g)]
check("group_norm (atomics)", lambda: ops.group_norm(xs, 32, 256, 8, gamma, beta, 1e-6, False)[0])
xi = torch.randn(8, 16, 16, 64, device="cuda", generator=g).to(BF16)
wc = ops.pack_conv3x3(torch.randn(64, 64, 3, 3, device="cuda", generator=g))
check("conv3x3 implicit", lambda: ops.conv3x3(xi, wc, 8, 16, 16, 64))
