"""Dev tool: per-kernel / per-shape time of one full-size UNet call and one 16-frame VAE decode (CUDA events)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
from emote_hack_b200 import ops  # noqa: E402
from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402
from emote_hack_b200.vae import AutoencoderKL  # noqa: E402

import os
ops.FUSED_GN_STATS = os.environ.get('FUSED', '1') != '0'
dev = torch.device("cuda")
torch.manual_seed(0)
with torch.device(dev):
    unet = UNet3DConditionModel(**FULL_CFG).eval()
    vae = AutoencoderKL().eval()
rerandomise_zero_inits(unet)
x = torch.randn(2, 4, 16, 64, 64, device=dev); ctx = torch.randn(2, 77, 768, device=dev)
for _ in range(2):
    unet(x, 981, ctx)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    unet(x, 981, ctx)
e1.record(); torch.cuda.synchronize()
print(f"UNet call: {e0.elapsed_time(e1)/3:.2f} ms (device), host wall {(time.perf_counter()-t0)/3*1e3:.2f} ms")
with ops.KernelProfiler() as prof:
    unet(x, 981, ctx)
tot = sum(ms for _, ms, _ in prof.times)
print(f"sum of kernels {tot:.2f} ms")
for k, (c, v) in sorted(prof.summary().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:36s} n={c:4d} {v:8.3f} ms {100*v/tot:5.1f}%")
print("GEMM by shape (M,N,K,taps): count, ms, TFLOP/s")
for m, (c, v) in sorted(prof.by_shape().items(), key=lambda kv: -kv[1][1]):
    print(f"  {str(m):34s} n={c:3d} {v:8.3f} ms  {2.0*m[0]*m[1]*m[2]*c/v/1e9:8.1f} TF/s")
print("attention by shape (batch,heads,d,nq,n0,n1)")
for m, (c, v) in sorted(prof.by_shape("emote_attention_bf16").items(), key=lambda kv: -kv[1][1]):
    fl = 4.0 * m[0] * m[1] * m[2] * m[3] * (m[4] + m[5]) * c
    print(f"  {str(m):34s} n={c:3d} {v:8.3f} ms  {fl/v/1e9:8.1f} TF/s")
for nm in ("emote_layernorm", "emote_gn_stats", "emote_gn_apply", "emote_temporal_attention_bf16"):
    print(nm)
    for m, (c, v) in sorted(prof.by_shape(nm).items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"  {str(m):40s} n={c:3d} {v:8.3f} ms")
lat = torch.randn(1, 4, 16, 64, 64, device=dev) * 0.18215
vae.decode_video(lat); torch.cuda.synchronize()
e0.record(); vae.decode_video(lat, want_u8=True); e1.record(); torch.cuda.synchronize()
print(f"VAE decode 16 frames: {e0.elapsed_time(e1):.2f} ms")
with ops.KernelProfiler() as prof:
    vae.decode_video(lat)
tot = sum(ms for _, ms, _ in prof.times)
for k, (c, v) in sorted(prof.summary().items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:36s} n={c:4d} {v:8.3f} ms {100*v/tot:5.1f}%")
for m, (c, v) in sorted(prof.by_shape().items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {str(m):34s} n={c:3d} {v:8.3f} ms  {2.0*m[0]*m[1]*m[2]*c/v/1e9:8.1f} TF/s")
