"""Dev tool: CPU issue time vs device time of one full UNet call, and the same call replayed as a CUDA graph."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
with torch.device(dev):
    unet = UNet3DConditionModel(**FULL_CFG).eval()
rerandomise_zero_inits(unet)
x = torch.randn(2, 4, 16, 64, 64, device=dev); ctx = torch.randn(2, 77, 768, device=dev)
t = torch.tensor([981.0], device=dev)
for _ in range(3):
    ref = unet(x, t, ctx).sample
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0 = time.perf_counter(); e0.record()
for _ in range(5):
    unet(x, t, ctx)
c1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); c2 = time.perf_counter()
print(f"eager: cpu issue {(c1-c0)/5*1e3:.2f} ms/call, device {e0.elapsed_time(e1)/5:.2f} ms/call, wall {(c2-c0)/5*1e3:.2f}")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    unet(x, t, ctx)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    out = unet(x, t, ctx).sample
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
print("graph vs eager rel diff", ((out - ref).norm() / ref.norm()).item())
for rep in range(3):
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"graph replay: device {e0.elapsed_time(e1)/20:.3f} ms/call")
