"""Dev tool: time individual GEMM shapes (CUDA events over many back-to-back launches)."""
import os
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import ops
BF16 = ops.OP16   # operand type of the loaded build
TS = int(os.environ.get('TMA_STORE', '0'))
PM = int(os.environ.get('PAIR', '0'))
ops.FORCE_BLOCK_N = int(os.environ.get('BLOCK_N', '0'))
shapes = [  # (M, N, K, mode)
    (131072, 2560, 320, "geglu"), (131072, 320, 320, "res"), (32768, 640, 640, "res"), (8192, 1280, 1280, "res"),
    (131072, 960, 320, "bf16"), (32768, 5120, 640, "geglu"), (8192, 10240, 1280, "geglu"), (131072, 320, 1280, "res"),
    (2048, 1280, 1280, "res"), (8192, 8192, 8192, "bf16"),
    (131072, 320, 2880, "conv320"), (8192, 1280, 11520, "conv1280"),
    (8192, 1280, 1280, "f32"), (8192, 1280, 1280, "bf16"), (32768, 640, 640, "f32"), (32768, 640, 640, "bf16"),  # 12-15
    (131072, 320, 320, "f32"), (131072, 320, 320, "bf16"), (8192, 1280, 5120, "res"), (2048, 1280, 1280, "bf16"),  # 16-19
]
only = sys.argv[1:] and [int(a) for a in sys.argv[1:]]
_w = torch.randn(8192, 8192, device="cuda").to(BF16)
for _ in range(60):  # bring the GPU to its steady clocks before timing anything
    _w @ _w
torch.cuda.synchronize()
del _w
for i, (M, N, K, mode) in enumerate(shapes):
    if only and i not in only:
        continue
    a = torch.randn(M, K, device="cuda").to(BF16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF16)
    bias = torch.randn(N, device="cuda")
    if mode.startswith("conv"):
        Cc = K // 9
        hw = {320: 64, 1280: 16}[Cc]
        n_img = M // (hw * hw)
        x = torch.randn(n_img, hw, hw, Cc, device="cuda").to(BF16)
        wc = ops.pack_conv3x3(torch.randn(N, Cc, 3, 3, device="cuda") / K ** 0.5)
        res = torch.randn(M, N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        fn = lambda: ops.conv3x3(x, wc, n_img, hw, hw, Cc, bias=bias, residual=res, out=out, tma_store=TS, pair_mode=PM)
    elif mode == "geglu":
        wp, bp = ops.pack_geglu(w.float(), bias)
        fn = lambda: ops.gemm(a, wp, bias=bp, geglu=True, out_dtype=BF16, tma_store=TS, pair_mode=PM)
    elif mode == "res":
        res = torch.randn(M, N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        fn = lambda: ops.gemm(a, w, bias=bias, residual=res, out=out, pair_mode=PM, tma_store=TS)
    elif mode == "f32":
        out = torch.empty(M, N, device="cuda")
        fn = lambda: ops.gemm(a, w, bias=bias, out=out, pair_mode=PM, tma_store=TS)
    else:
        out = torch.empty(M, N, device="cuda", dtype=BF16)
        fn = lambda: ops.gemm(a, w, bias=bias, out_dtype=BF16, out=out, tma_store=TS, pair_mode=PM)
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    n, rounds = 30, 7
    ts = []
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    ts.sort()
    us, med = ts[0], ts[len(ts) // 2]
    print(f"{i}: M={M} N={N} K={K} {mode:6s} min {us:8.1f} us  med {med:8.1f} us  {2.0*M*N*K/us/1e6:8.1f} TF/s")
