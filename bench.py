#!/usr/bin/env python
"""bench.py — denoised frames/sec of the Emote-hack hot path on B200 (BASELINE.json metric).

A "step" is one pass of the whole path over one synthetic clip: 50 DDIM timesteps of
[UNet3DConditionModel forward on the CFG pair [2,4,16,64,64] + fused CFG/DDIM update] followed by the VAE decode of the
16 frames to 512x512 (config #2 of BASELINE.json: 512x512 ref, 16-frame window, 50 DDIM steps, bf16 tensor-core
operands, random-init SD-1.5-width weights with motion modules on, synthetic latents / text context).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU PyTorch path (oracle port)

Under `python -m torch.distributed.run --nproc-per-node N` every rank generates its own clip (weak scaling: the
noise-sample axis of the per-step latent batch is sharded, no data-path collective) and the decoded uint8 frames are
all-gathered once per step over NCCL.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "denoised frames/sec at 512x512x16f, 50-step DDIM"
WORKLOAD = ("512x512 ref, 16-frame window, 50 DDIM steps, CFG pair [2,4,16,64,64], SD-1.5-width UNet3D (1276.7 M params, "
            "motion modules on), text ctx [2,77,768], VAE decode to 16x3x512x512 (BASELINE.json configs[1]); one clip per GPU")
UNIT = "frames/s"
FRAMES, LAT, DDIM_STEPS, GUIDANCE = 16, 64, 50, 7.5
# algorithmic FLOPs of one UNet call on [2,4,16,64,64] (FlopCounterMode on the reference)
UNET_TFLOP_PER_CALL = 35.35
GEMM_TFLOP_PER_CALL = 31.27   # conv 14.21 + addmm 12.27 + mm 4.79 : the part the tcgen05 GEMM kernel executes
VAE_TFLOP_PER_FRAME = 2.5


def full_unet_cfg():
    from util_models import FULL_CFG
    return dict(FULL_CFG)


def host_cores() -> int:
    """Threads the CPU arm may really use: affinity mask capped by the cgroup CPU quota (a 128-thread pool on a
    box whose container is limited to a few cores only thrashes)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        v2 = Path("/sys/fs/cgroup/cpu.max")
        if v2.exists():
            quota, period = v2.read_text().split()
            if quota != "max":
                n = max(1, min(n, int(float(quota) / float(period))))
        else:
            quota = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            period = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if quota > 0:
                n = max(1, min(n, quota // period))
    except Exception:
        pass
    return n


def gemm_dram_traffic_per_launch():
    """DRAM bytes per launch of the dominant (GEMM / implicit-GEMM conv) kernels, from the committed ncu capture of one
    eager UNet call (scripts/one_unet_call.py + scripts/summarize_traffic.py); None when the capture is absent."""
    try:
        d = json.loads((ROOT / "profiles" / "r01_unet_call_dram_traffic.json").read_text())
        return round(float(d["all_gemm_kernels"]["dram_bytes_per_launch"]))
    except Exception:
        return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"bf16_tflops": d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "hbm_gbs": d.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in Path(self.path).read_text().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# =============================================================================================== CPU reference arm
def cpu_reference_sample(threads: int, unet_sd=None, vae_sd=None, seed: int = 0):
    """Times the reference's CPU PyTorch path (oracle port of UNet3D + restated VAE decoder) on a bounded sample and
    extrapolates the clip: t_clip = 50 steps x 32 frame-evaluations x t(frame-eval) + 16 x t(VAE frame).
    A frame-evaluation = one UNet forward on [1,4,1,64,64] (1.105 TFLOP); 5-D GroupNorm / temporal attention cost
    scales linearly in frames, so this is the per-frame cost of the [2,4,16,64,64] call up to cache effects."""
    import torch
    from oracle.unet3d_port import UNet3DOracle
    from oracle.vae_decoder import VAEDecoderOracle, random_vae_decoder_state_dict
    torch.set_num_threads(threads)
    cfg = full_unet_cfg()
    if unet_sd is None:
        from emote_hack_b200.unet3d import UNet3DConditionModel
        torch.manual_seed(seed)
        unet_sd = UNet3DConditionModel(**cfg).state_dict()
    if vae_sd is None:
        vae_sd = random_vae_decoder_state_dict(seed=seed)
    unet = UNet3DOracle(unet_sd, cfg)
    vae = VAEDecoderOracle(vae_sd)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 4, 1, LAT, LAT, generator=g)
    ctx = torch.randn(1, 77, 768, generator=g)
    z = torch.randn(1, 4, LAT, LAT, generator=g)
    t0 = time.perf_counter(); unet(x, 981, ctx); t_first = time.perf_counter() - t0
    t0 = time.perf_counter(); unet(x, 961, ctx); t_unet = time.perf_counter() - t0
    t0 = time.perf_counter(); vae.decode(z); t_vae = time.perf_counter() - t0
    t_clip = DDIM_STEPS * 2 * FRAMES * t_unet + FRAMES * t_vae
    return {"fps": FRAMES / t_clip, "t_frame_eval_s": t_unet, "t_frame_eval_first_s": t_first, "t_vae_frame_s": t_vae,
            "t_clip_extrapolated_s": t_clip}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_cores()
    vals = []
    sample = None
    # each "step" = one bounded sample (1 UNet frame-evaluation + 1 VAE frame), extrapolated to the clip
    import torch
    torch.set_num_threads(threads)
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from oracle.vae_decoder import random_vae_decoder_state_dict
    torch.manual_seed(0)
    unet_sd = UNet3DConditionModel(**full_unet_cfg()).state_dict()
    vae_sd = random_vae_decoder_state_dict(seed=0)
    for i in range(args.warmup + args.steps):
        s = cpu_reference_sample(threads, unet_sd, vae_sd)
        if i >= args.warmup:
            vals.append(s["fps"])
            sample = s
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * FRAMES / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "oracle port (fp32 restatement of the reference, validated against it): 1 UNet "
                                   f"frame-evaluation [1,4,1,64,64] ({sample['t_frame_eval_s']:.2f} s) + 1 VAE frame "
                                   f"({sample['t_vae_frame_s']:.2f} s), extrapolated x(50 steps x 32 frame-evals) + 16 frames"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# =============================================================================================== CUDA arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))  # communicator warm-up (reference dist_tools.py:55)

    from emote_hack_b200 import _lib, ops
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from emote_hack_b200.vae import AutoencoderKL
    from util_models import rerandomise_zero_inits

    cfg = full_unet_cfg()
    torch.manual_seed(0)
    with torch.device(dev):
        unet = UNet3DConditionModel(**cfg).eval()
        vae = AutoencoderKL().eval()
    rerandomise_zero_inits(unet)
    pipe = EMOAnimationPipeline(vae, unet, DDIMScheduler(), rank=rank, world_size=1)  # one sample per rank

    g = torch.Generator().manual_seed(1234 + rank)
    host_lat = torch.randn(1, 4, FRAMES, LAT, LAT, generator=g).pin_memory()
    host_ctx = torch.randn(2, 77, 768, generator=g).pin_memory()
    host_out = torch.empty((1, 3, FRAMES, 8 * LAT, 8 * LAT), dtype=torch.uint8).pin_memory()
    lat_dev, ctx_dev = host_lat.to(dev), host_ctx.to(dev)

    def one_clip(lat, ctx, gather: bool):
        lat = pipe.denoise(lat, ctx, num_inference_steps=DDIM_STEPS, guidance_scale=GUIDANCE, context_frames=FRAMES)
        _, u8 = vae.decode_video(lat, want_u8=True)
        u8 = u8.contiguous()
        if gather and world > 1:
            outs = [torch.empty_like(u8) for _ in range(world)]
            dist.all_gather(outs, u8)  # the single collective of the path: decoded frames over NVLink
        return u8

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_clip(lat_dev.clone(), ctx_dev, True)
    sync_all()

    # ---- timed region 1: inputs resident in HBM ("value")
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    from emote_hack_b200.pipeline import GraphedUNet
    launches0 = _lib.launch_count() + GraphedUNet.replayed_kernels
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lats = [lat_dev.clone() for _ in range(args.steps)]
    sync_all()
    e0.record()
    for i in range(args.steps):
        one_clip(lats[i], ctx_dev, True)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() + GraphedUNet.replayed_kernels - launches0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * FRAMES * args.steps / (ms_max / 1000.0)

    # ---- timed region 2: end to end through the public API with HOST buffers ("e2e")
    sync_all()
    e0.record()
    for i in range(args.steps):
        lat = host_lat.to(dev, non_blocking=True)
        ctx = host_ctx.to(dev, non_blocking=True)
        u8 = one_clip(lat, ctx, True)
        host_out.copy_(u8, non_blocking=True)
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * FRAMES * args.steps / (float(t.item()) / 1000.0)
    h2d = host_lat.numel() * 4 + host_ctx.numel() * 4
    d2h = host_out.numel()

    # ---- secondary measurement (N=1, rank 0): the same clip with the ReferenceNet on — AppearanceEncoderModel writer once
    # per timestep + reference attention in the 10 mid/up reader blocks; not part of `value`
    variants = None
    if rank == 0 and world == 1 and not args.no_variants:
        from emote_hack_b200.appearance_encoder import AppearanceEncoderModel
        torch.manual_seed(1)
        with torch.device(dev):
            enc = AppearanceEncoderModel(sample_size=64, cross_attention_dim=768).eval()
        ref_lat = torch.randn(1, 4, LAT, LAT, generator=torch.Generator().manual_seed(7)).to(dev)

        def ref_clip(lat):
            lat = pipe.denoise(lat, ctx_dev, num_inference_steps=DDIM_STEPS, guidance_scale=GUIDANCE,
                               context_frames=FRAMES, appearance_encoder=enc, ref_image_latents=ref_lat)
            return vae.decode_video(lat, want_u8=True)[1]

        ref_clip(lat_dev.clone())
        torch.cuda.synchronize()
        n_var = 2
        e0.record()
        for _ in range(n_var):
            ref_clip(lat_dev.clone())
        e1.record()
        torch.cuda.synchronize()
        variants = {"reference_net_on": {"value": round(FRAMES * n_var / (e0.elapsed_time(e1) / 1000.0), 4), "unit": UNIT,
                                         "clips": n_var, "what": "config #2 + AppearanceEncoderModel (2-D SD UNet, 859.5 M "
                                         "params) run once per DDIM step on the reference-image latents [2,4,64,64]; its 10 "
                                         "LayerNorm1 banks extend the keys of the mid/up spatial self-attention (cond half)"}}
        del enc
        torch.cuda.empty_cache()

    # ---- roofline pass (untimed): every launch of one UNet call bracketed by CUDA events on the launching stream
    roof = None
    breakdown = None
    if rank == 0:
        with ops.KernelProfiler() as prof:
            unet(lat_dev.expand(2, -1, -1, -1, -1).contiguous(), 981, ctx_dev)
        peaks = measured_peaks()
        tf = prof.gemm_flops / 1e12
        achieved = tf / (prof.gemm_ms / 1e3)
        n_gemm = sum(1 for n, _, _ in prof.times if n == "emote_gemm_bf16")
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": round(achieved, 1),
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": round(achieved / peaks["bf16_tflops"], 4),
                "traffic": gemm_dram_traffic_per_launch(), "traffic_unit": "bytes/launch (dram__bytes_read.sum + "
                "dram__bytes_write.sum, mean over the GEMM launches of one UNet call; profiles/r01_unet_call_dram_traffic.json)",
                "peak_source": peaks["source"],
                "note": f"{tf:.2f} algorithmic TFLOP (2*M*N*K of the reference's conv/linear ops) in {n_gemm} launches "
                        f"of one UNet call, {prof.gemm_ms:.2f} ms total"}
        total = sum(ms_ for _, ms_, _ in prof.times)
        breakdown = {k: {"launches": c, "ms": round(v, 3), "share": round(v / total, 4)}
                     for k, (c, v) in sorted(prof.summary().items(), key=lambda kv: -kv[1][1])}
        breakdown["_total_ms_one_unet_call"] = round(total, 3)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_cores()
        sd = {k: v.detach().cpu() for k, v in unet.state_dict().items()}
        vsd = {k: v.detach().cpu() for k, v in vae.state_dict().items()}
        s = cpu_reference_sample(threads, sd, vsd)
        cpu = {"value": s["fps"], "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle port, same weights: 1 UNet frame-evaluation [1,4,1,64,64] ({s['t_frame_eval_s']:.2f} s) "
                         f"+ 1 VAE frame 512x512 ({s['t_vae_frame_s']:.2f} s), extrapolated to 50 steps x 32 "
                         "frame-evaluations + 16 frames"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if _lib.OPERAND == "fp16" else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "precision": f"{_lib.OPERAND} tensor-core operands (EMOTE_OPERAND), fp32 accumulate / residual stream / statistics",
                       "l2": "per-step working set (2.6 GB packed weights + >5 GB activations) >> 126 MB L2; no flush needed",
                       "unet_tflop_per_call": UNET_TFLOP_PER_CALL},
            "clocks": clk,
            "e2e": {"value": round(e2e_value, 4), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_breakdown_one_unet_call": breakdown,
            "variants": variants,
            "model_tflops_per_s": round((DDIM_STEPS * UNET_TFLOP_PER_CALL + FRAMES * VAE_TFLOP_PER_FRAME) * args.steps
                                        / (ms_max / 1000.0), 1),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """the ONE JSON line of the contract goes to the real stdout; everything else (NCCL banners, library prints)
    was redirected to stderr at start-up"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the rest of the process (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the secondary ReferenceNet-on measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: self-launch one rank per GPU the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", str(Path(__file__).resolve()),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd, stdout=_REAL_STDOUT)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
