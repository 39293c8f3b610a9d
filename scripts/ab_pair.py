"""Dev tool: CTA-pair (cta_group::2) GEMM kernel vs torch and vs the single-CTA kernel."""
import sys, math
from pathlib import Path
import torch
import torch.nn.functional as F
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import ops
BF16 = torch.bfloat16
g = torch.Generator(device="cuda"); g.manual_seed(0)
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
def tm(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("correctness (pair_mode=1)")
for (M, N, K) in [(512, 160, 64), (256, 320, 1280), (300, 136, 72), (4096, 640, 2560), (131072, 320, 320)]:
    a = torch.randn(M, K, device="cuda", generator=g).to(BF16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(BF16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias + res
    out = ops.gemm(a, w, bias=bias, residual=res, pair_mode=1)
    torch.cuda.synchronize()
    print(f"  gemm+res M={M} N={N} K={K}: rel={rel(out, ref):.2e}")
    outb = ops.gemm(a, w, bias=bias, out_dtype=BF16, pair_mode=1)
    print(f"  gemm bf16 M={M} N={N} K={K}: rel={rel(outb, ref - res):.2e}")
a = torch.randn(1024, 1280, device="cuda", generator=g).to(BF16)
w = torch.randn(2560, 1280, device="cuda", generator=g) / math.sqrt(1280); b = torch.randn(2560, device="cuda", generator=g)
h = a.float() @ w.to(BF16).float().t() + b
wp, bp = ops.pack_geglu(w, b)
print(f"  geglu: rel={rel(ops.gemm(a, wp, bias=bp, geglu=True, out_dtype=BF16, pair_mode=1), h[:, :1280] * F.gelu(h[:, 1280:])):.2e}")
for (n_img, hw, C, N) in [(2, 8, 64, 128), (4, 16, 128, 160), (32, 64, 320, 320), (3, 32, 64, 320)]:
    x = torch.randn(n_img, hw, hw, C, device="cuda", generator=g).to(BF16)
    wc = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(BF16)
    bias = torch.randn(N, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wc.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    out = ops.gemm(x, ops.pack_conv3x3(wc.float()), bias=bias, conv=(n_img, hw, hw, C), pair_mode=1)
    print(f"  conv n={n_img} hw={hw} C={C} N={N}: rel={rel(out, ref):.2e}")
print("timing pair (1) vs single (2)")
for (M, N, K, conv) in [(131072, 320, 2880, True), (8192, 1280, 11520, True), (8192, 8192, 8192, False), (32768, 640, 2560, False),
                        (131072, 320, 1280, False), (8192, 10240, 1280, False)]:
    if conv:
        C = K // 9; hw = {320: 64, 1280: 16}[C]; n_img = M // (hw * hw)
        x = torch.randn(n_img, hw, hw, C, device="cuda").to(BF16)
        wc = ops.pack_conv3x3(torch.randn(N, C, 3, 3, device="cuda") / K ** 0.5)
        res = torch.randn(M, N, device="cuda"); out = torch.empty(M, N, device="cuda")
        fns = {m: (lambda m=m: ops.gemm(x, wc, residual=res, out=out, conv=(n_img, hw, hw, C), pair_mode=m)) for m in (1, 2)}
    else:
        a = torch.randn(M, K, device="cuda").to(BF16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF16)
        out = torch.empty(M, N, device="cuda", dtype=BF16)
        fns = {m: (lambda m=m: ops.gemm(a, w, out_dtype=BF16, out=out, pair_mode=m)) for m in (1, 2)}
    t1, t2 = tm(fns[1]), tm(fns[2])
    print(f"  M={M} N={N} K={K} conv={conv}: pair {t1:8.1f} us ({2.0*M*N*K/t1/1e6:7.1f} TF/s)   single {t2:8.1f} us ({2.0*M*N*K/t2/1e6:7.1f} TF/s)")
