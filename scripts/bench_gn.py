"""Dev tool: A/B of the GroupNorm statistics fold (emote_set_tuning "gn_reduce": 1 = flat fold, 0 = per-slot walk) on the
UNet's shapes, each timed as 40 launches replayed from one CUDA graph, then one full-size UNet call under both settings."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from emote_hack_b200 import _lib, ops  # noqa: E402
from emote_hack_b200._lib import check  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")


def graph_time(fn, n=40, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)   # us per launch


# (C_src, c_offset, C_total, slots_per_batch, n_batches): 5-D GroupNorm of the resnets (2 samples x 16 frames), the per-frame
# GroupNorm of Transformer3DModel (32 frames), second source of an up-block concat
shapes = [(320, 0, 320, 2048, 2), (640, 0, 640, 512, 2), (1280, 0, 1280, 128, 2), (1280, 0, 1280, 32, 2),
          (320, 0, 320, 128, 32), (640, 0, 640, 32, 32), (1280, 0, 1280, 8, 32), (320, 320, 640, 2048, 2),
          (640, 1280, 1920, 128, 2)]
print(f"operand={_lib.OPERAND}")
for (cs, off, ct, slots, nb) in shapes:
    sl = torch.randn(nb, slots, cs, 2, device=dev)
    res = {}
    for mode in (0, 1):
        ops.set_tuning("gn_reduce", mode)
        sums = torch.zeros(nb, 32, 2, dtype=torch.float64, device=dev)
        fn = lambda: check(lib.emote_gn_colstats_reduce(sl.data_ptr(), cs, off, ct, 32, slots, nb, sums.data_ptr(), 1,
                                                        torch.cuda.current_stream().cuda_stream), "reduce")
        us = graph_time(fn)
        res[mode] = (us, sums.clone())
    d = (res[0][1] - res[1][1]).abs().max().item()
    print(f"reduce C_src={cs} off={off} C_total={ct} slots={slots} batches={nb}: per-slot {res[0][0]:.2f} us, flat {res[1][0]:.2f} us, "
          f"max |diff| {d:.2e}")

# gn_apply on the same levels (rows_per_batch, n_batches, C)
for (rows, nb, c) in [(65536, 2, 320), (16384, 2, 640), (4096, 2, 1280), (1024, 2, 1280), (4096, 32, 320), (1024, 32, 640),
                      (256, 32, 1280)]:
    x = torch.randn(rows * nb, c, device=dev)
    gamma, beta = torch.randn(c, device=dev), torch.randn(c, device=dev)
    sums = torch.empty(nb, 32, 2, dtype=torch.float64, device=dev)
    xs = x.view(nb, rows, 32, c // 32).double()
    sums[..., 0], sums[..., 1] = xs.sum((1, 3)), (xs * xs).sum((1, 3))
    out = torch.empty(rows * nb, c, dtype=ops.OP16, device=dev)
    fn = lambda: check(lib.emote_gn_apply(x.data_ptr(), c, 0, c, 32, rows, nb, sums.data_ptr(), gamma.data_ptr(),
                                          beta.data_ptr(), 1e-5, 1, out.data_ptr(), None,
                                          torch.cuda.current_stream().cuda_stream), "apply")
    us = graph_time(fn, n=10)
    gb = rows * nb * c * 6 / 1e9
    print(f"gn_apply rows/batch={rows} batches={nb} C={c}: {us:.2f} us  {gb / us * 1e6:.0f} GB/s")

if "--no-unet" not in sys.argv:
    from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
    from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402
    torch.manual_seed(0)
    with torch.device(dev):
        unet = UNet3DConditionModel(**FULL_CFG).eval()
    rerandomise_zero_inits(unet)
    x = torch.randn(2, 4, 16, 64, 64, device=dev); ctx = torch.randn(2, 77, 768, device=dev)
    outs = {}
    for mode in (0, 1, 0, 1):
        ops.set_tuning("gn_reduce", mode)
        for _ in range(2):
            unet(x, 981, ctx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            o = unet(x, 981, ctx).sample
        e1.record(); torch.cuda.synchronize()
        outs[mode] = o.clone()
        print(f"UNet call gn_reduce={mode}: {e0.elapsed_time(e1) / 5:.2f} ms")
    print("flat vs per-slot rel diff", ((outs[0] - outs[1]).norm() / outs[0].norm()).item())
    ops.set_tuning("gn_reduce", 1)
