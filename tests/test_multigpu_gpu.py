"""Multi-GPU path on real NCCL (skipped with fewer than 2 GPUs): the (window x CFG-branch) units of one timestep dealt to
the ranks (`pipeline.plan_units`; the reference splits whole windows, EMOAnimationPipeline.py:757) with ONE
all-reduce of the accumulated prediction per step instead of the reference's gather + broadcast + barriers (:796-821), and
the frame-sharded VAE decode with its single uint8 all-gather.  Every scenario must reproduce the single-GPU result:
4 windows (whole pairs per rank), 3 windows (one window's two branches on different ranks), 1 window (one branch per rank —
the case where ranks > 0 used to diverge), and the reference's own window split.  The CPU/gloo twin of the reduction lives
in test_host_logic.py."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    import torch.distributed as dist
    from util_models import TINY_CFG, make_banks, rerandomise_zero_inits
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from emote_hack_b200.vae import AutoencoderKL
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    unet = rerandomise_zero_inits(UNet3DConditionModel(**TINY_CFG).eval()).to(dev)
    torch.manual_seed(1)
    vae = AutoencoderKL(block_out_channels=(64, 64, 128, 128)).eval().to(dev)
    banks = {k: [t.to(dev) for t in v] for k, v in make_banks(unet, 8).items()}
    single = EMOAnimationPipeline(vae, unet, DDIMScheduler(), rank=0, world_size=1)
    sharded = EMOAnimationPipeline(vae, unet, DDIMScheduler(), rank=rank, world_size=world)
    results = []
    for frames, shard, use_banks in ((24, "units", True), (18, "units", True), (8, "units", False), (24, "windows", True)):
        g = torch.Generator().manual_seed(31 + frames)
        lat = torch.randn(1, 4, frames, 8, 8, generator=g).to(dev)
        ctx = torch.randn(2, 7, 64, generator=g).to(dev)
        kw = dict(num_inference_steps=2, guidance_scale=7.5, context_frames=8, context_overlap=2,
                  reference_banks=banks if use_banks else None)
        want = single.denoise(lat.clone(), ctx, **kw)
        _, want_u8 = single.decode_latents_device(want, want_u8=True)
        got = sharded.denoise(lat.clone(), ctx, shard=shard, **kw)
        _, got_u8 = sharded.decode_latents_device(got, want_u8=True, shard=True)
        err = ((got - want).norm() / want.norm()).item()
        px = (got_u8.int() - want_u8.int()).abs().max().item()
        # whole windows per rank run the very same batch-2 calls as the single GPU: fp32 round-off.  When a window's two CFG
        # branches sit on different ranks they run as batch-1 calls, whose GEMM tile schedule differs from the batch-2
        # call's; this random-weight network amplifies that round-off to the operand-rounding noise floor (see
        # tests/test_unet_gpu.py::test_single_branch_calls_equal_the_cfg_pair), still 10x below any real divergence
        from emote_hack_b200 import _lib
        split = shard == "units" and frames != 24
        lim, pxlim = ((4e-3 if _lib.OPERAND == "fp16" else 3e-2), 4) if split else (1e-5, 1)
        ok = torch.tensor([1.0 if (err < lim and px <= pxlim) else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # every rank must hold the right latents, not only rank 0
        results.append((frames, shard, bool(ok.item() == 1.0), err, px))
    if rank == 0:
        ret.put(results)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_window_sharding_and_frame_sharded_decode():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for frames, shard, ok, err, px in ret.get(timeout=10):
        print(f"2-GPU {shard} sharding, {frames} frames: rel diff vs single GPU {err:.2e}, max pixel diff {px}")
        assert ok, (frames, shard, err, px)
