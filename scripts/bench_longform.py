"""BASELINE config #5 (long-form 240 frames, sliding 16-frame windows with overlap 4 = 20 windows per DDIM step) on N GPUs:
the windows of one step are sharded across ranks (`windows[rank::world]`, EMOAnimationPipeline.py:757) and the accumulated
prediction is summed with ONE NCCL all-reduce per step; every rank applies the fused CFG + DDIM update redundantly.
Strong scaling: total work fixed.  Launch:  python scripts/bench_longform.py            (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_longform.py
Prints one JSON line (rank 0): ms per DDIM step (max over ranks, CUDA events) and frames/s extrapolated to 50 steps."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline, uniform  # noqa: E402
from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=240)
ap.add_argument("--steps", type=int, default=6, help="timed DDIM steps (after 2 warm-up steps)")
args = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    dist.all_reduce(torch.zeros(1, device=dev))
torch.manual_seed(0)
with torch.device(dev):
    unet = UNet3DConditionModel(**FULL_CFG).eval()
rerandomise_zero_inits(unet)
pipe = EMOAnimationPipeline(None, unet, DDIMScheduler(), rank=rank, world_size=world)
g = torch.Generator().manual_seed(1234)
lat = torch.randn(1, 4, args.frames, 64, 64, generator=g).to(dev)
ctx = torch.randn(2, 77, 768, generator=g).to(dev)
windows = list(uniform(0, 50, args.frames, 16, 1, 4))
marks = []


def cb(i, t, latents):   # called after every DDIM step
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append(e)


warm = 2
pipe.denoise(lat.clone(), ctx, num_inference_steps=warm + args.steps, context_frames=16, context_overlap=4, callback=cb)
torch.cuda.synchronize()
ms = marks[warm - 1].elapsed_time(marks[-1]) / args.steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
if rank == 0:
    print(json.dumps({"workload": f"{args.frames} frames, {len(windows)} windows of 16 (overlap 4) per DDIM step, CFG pair, "
                      "SD-1.5-width UNet3D with motion modules (BASELINE.json configs[4]); VAE decode not included",
                      "n_gpus": world, "windows_per_rank": [len(windows[r::world]) for r in range(world)],
                      "ms_per_ddim_step": round(ms, 2), "timed_steps": args.steps,
                      "frames_per_s_at_50_steps": round(args.frames / (50 * ms / 1000.0), 3), "scaling": "strong",
                      "collective": "one all-reduce of the fp32 prediction [2,4,F,64,64] per step" if world > 1 else "none"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
