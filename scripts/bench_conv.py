"""Dev tool: 3x3 implicit-GEMM conv timing at the BASELINE config #4 map sizes (96 / 48 / 24 / 12: 2-D patch tiles) next to
the config #2 sizes (64 / 32 / 16 / 8: consecutive-pixel tiles) and to the explicit im2col + GEMM path."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from emote_hack_b200 import _lib, ops
from emote_hack_b200._lib import check
OP16 = ops.OP16


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (n_img, hw, c, n) in [(32, 96, 320, 320), (32, 64, 320, 320), (32, 48, 640, 640), (32, 32, 640, 640),
                          (32, 24, 1280, 1280), (32, 16, 1280, 1280), (32, 12, 1280, 1280), (32, 8, 1280, 1280),
                          (32, 96, 960, 320)]:
    x = torch.randn(n_img, hw, hw, c, device="cuda").to(OP16)
    w = ops.pack_conv3x3(torch.randn(n, c, 3, 3, device="cuda") / (9 * c) ** 0.5)
    res = torch.randn(n_img * hw * hw, n, device="cuda")
    out = torch.empty_like(res)
    t = timeit(lambda: ops.conv3x3(x, w, n_img, hw, hw, c, residual=res, out=out))
    fl = 2.0 * n_img * hw * hw * n * 9 * c

    def explicit():
        cols = torch.empty((n_img * hw * hw, 9 * c), dtype=OP16, device="cuda")
        check(_lib.load().emote_im2col3x3_bf16(x.data_ptr(), n_img, hw, hw, c, 1, cols.data_ptr(), ops._stream()), "im2col")
        ops.gemm(cols, w, residual=res, out=out)
    t2 = timeit(explicit, 3)
    print(f"conv3x3 n_img={n_img} {hw}x{hw} C={c} N={n}: implicit {t*1e3:.0f} us = {fl/t/1e9:.0f} TF/s | im2col+GEMM {t2*1e3:.0f} us = {fl/t2/1e9:.0f} TF/s")
