"""Dev tool: same-process A/B of library switches on one full-size UNet call replayed as a CUDA graph.
usage: python scripts/ab_unet.py [pdl] [fused]   (which switches to toggle; default both)"""
import itertools
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
from emote_hack_b200 import ops  # noqa: E402
from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402

which = sys.argv[1:] or ["pdl", "fused"]
dev = torch.device("cuda")
torch.manual_seed(0)
with torch.device(dev):
    unet = UNet3DConditionModel(**FULL_CFG).eval()
rerandomise_zero_inits(unet)
x = torch.randn(2, 4, 16, 64, 64, device=dev); ctx = torch.randn(2, 77, 768, device=dev)
t = torch.tensor([981.0], device=dev)


def configure(cfg):
    ops.set_pdl(cfg.get("pdl", True))
    ops.FUSED_GN_STATS = cfg.get("fused", True)


graphs = {}
for vals in itertools.product([True, False], repeat=len(which)):
    cfg = dict(zip(which, vals))
    configure(cfg)
    for _ in range(2):
        unet(x, t, ctx)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = unet(x, t, ctx).sample
    g.replay(); torch.cuda.synchronize()
    graphs[tuple(cfg.items())] = (g, out)
ref = None
for k, (g, out) in graphs.items():
    g.replay(); torch.cuda.synchronize()
    o = out.clone()
    if ref is None:
        ref = o
    print(k, "rel diff vs first config:", ((o - ref).norm() / ref.norm()).item(), "finite:", bool(torch.isfinite(o).all()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rnd in range(3):
    for k, (g, out) in graphs.items():
        g.replay(); torch.cuda.synchronize()
        e0.record()
        for _ in range(8):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"round {rnd} {dict(k)}: {e0.elapsed_time(e1)/8:.3f} ms/call")
