"""TEST INFRASTRUCTURE ONLY — generates tests/golden/* from the UNTOUCHED reference modules (run in the build
container, where /root/reference exists):   python -m oracle.make_golden

The reference ships no tests / golden vectors, so known answers are produced here by importing the
reference's own `UNet3DConditionModel`, `ReferenceAttentionControl` and `context.uniform` through
oracle/ref_shim.py, loading deterministic name-keyed weights (tests/util_models.seeded_unet_state_dict) and recording
outputs for seeded inputs.  Only inputs' seeds, outputs and key/shape lists are stored (small); weights are
re-derived from the seeds by the tests.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from oracle import ref_shim  # noqa: E402
from util_models import (FULL_CFG, TINY_CFG, controlnet_residuals, make_banks, make_inputs, reader_block_names,  # noqa: E402
                         seeded_unet_state_dict, writer_cfg, writer_inputs)

GOLD = ROOT / "tests" / "golden"
# tag -> (block_out_channels, frames, latent size, weight seed); tests/test_oracle.py re-derives weights and inputs
VAE_CASES = {"tiny": ((32, 32, 64, 64), 2, 8, 1), "mid": ((64, 128, 128, 128), 1, 16, 2)}
# encoder: tag -> (block_out_channels, images, image size, weight seed)
VAE_ENC_CASES = {"tiny": ((32, 32, 64, 64), 2, 32, 1), "mid": ((64, 128, 128, 128), 1, 64, 2)}


DDIM_CASES = [(50, 981), (50, 481), (50, 21), (50, 1), (20, 951), (20, 1)]   # (num_inference_steps, timestep)


def make_ddim_golden():
    """Known answers of the reference's own DDIM algebra (EMOAnimationPipeline.next_step :379-400 and
    magicanimate/utils/util.py:64-74), executed from the reference sources (oracle/ref_ddim.py): the inversion step as
    written, and the same code run with the two alphas exchanged = the forward sampler step of `scheduler.step`."""
    from oracle import ref_ddim
    from oracle.ddim import DDIMOracle
    method, util_fn = ref_ddim.reference_next_step_method(), ref_ddim.reference_next_step_util()
    abar = DDIMOracle().alphas_cumprod       # betas of configs/inference.yaml:23-26
    out = {"cases": DDIM_CASES}
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 4, 2, 8, 8, generator=g)
    eps = torch.randn(1, 4, 2, 8, 8, generator=g)
    out["x"], out["eps"] = x, eps
    import types
    for n, t in DDIM_CASES:
        stub = ref_ddim.scheduler_stub(abar, n)
        x_next, pred_x0 = method(types.SimpleNamespace(scheduler=stub), eps, t, x)
        assert torch.equal(x_next, util_fn(eps, t, x, stub))        # the two reference copies agree bit for bit
        out[f"invert_{n}_{t}"], out[f"x0_{n}_{t}"] = x_next, pred_x0
        fwd = ref_ddim.swapped_alpha_stub(abar, t, n)
        out[f"step_{n}_{t}"] = method(types.SimpleNamespace(scheduler=fwd), eps, t, x)[0]
    torch.save(out, GOLD / "ddim_reference_steps.pt")


def make_cold_branch_golden():
    """Cold branches of UNet3DConditionModel.forward recorded from the reference: a latent size that is not a multiple of
    2**num_upsamplers (forced `upsample_size`, unet_controlnet.py:355-364,458-460) and the three class-embedding types
    (:121-128, 400-408)."""
    U = ref_shim.load_reference_unet_class()
    out = {}
    with torch.no_grad():
        for tag, extra in (("odd_size", {}), ("class_table", {"num_class_embeds": 4}),
                           ("class_timestep", {"class_embed_type": "timestep"}), ("class_identity", {"class_embed_type": "identity"})):
            m = U(**dict(TINY_CFG, **extra)).eval()
            shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
            out[f"{tag}_keys"] = shapes
            m.load_state_dict(seeded_unet_state_dict(shapes, seed=4), strict=True)
            hw = 10 if tag == "odd_size" else 8
            x, ctx = make_inputs(2, 2, hw, seed=77)
            kw = {}
            if tag == "class_table":
                kw["class_labels"] = torch.tensor([1, 3])
            elif tag == "class_timestep":
                kw["class_labels"] = torch.tensor([7.0, 250.0])
            elif tag == "class_identity":
                kw["class_labels"] = torch.randn(2, 256, generator=torch.Generator().manual_seed(78))
            out[tag] = m(x, torch.tensor(301), ctx, **kw).sample.contiguous()
            out[f"{tag}_labels"] = kw.get("class_labels")
    torch.save(out, GOLD / "unet3d_tiny_cold_branches.pt")


def make_speed_encoder_golden():
    """SpeedEncoder (Net.py:198-258) executed from the reference source: inputs, its own randomly initialised MLP weights
    and outputs."""
    from oracle import ref_audio
    SE = ref_audio.reference_speed_encoder_class()
    torch.manual_seed(5)
    enc = SE(9, 64)
    speeds = torch.tensor([-1.3, -0.55, -0.12, 0.0, 0.07, 0.21, 0.5, 0.93, 2.0], dtype=torch.float32)
    with torch.no_grad():
        out = enc(speeds)
    torch.save({"speeds": speeds, "out": out, "state_dict": {k: v.clone() for k, v in enc.state_dict().items()},
                "centers": list(enc.bucket_centers), "radii": list(enc.bucket_radii)}, GOLD / "speed_encoder.pt")


VIDEO_GRID_CASES = [  # (b, c, t, h, w, n_rows, rescale)
    (3, 3, 2, 5, 7, 2, False), (1, 3, 3, 6, 4, 6, False), (7, 3, 1, 4, 4, 6, True), (2, 1, 2, 3, 5, 6, False),
    (4, 3, 2, 8, 8, 4, True)]


def make_video_grid_golden():
    """Frames `save_videos_grid` (magicanimate/utils/util.py:21-33) hands to its writer, and `linear` / `slerp`
    (:125-141), from the untouched reference file.  `imageio` is absent here: a stand-in module records the arguments of
    `mimsave` (it performs no arithmetic)."""
    import importlib.util
    import tempfile
    import types
    rec = {}
    fake = types.ModuleType("imageio")
    fake.mimsave = lambda path, frames, fps=None, **kw: rec.update(frames=frames, fps=fps)
    had = sys.modules.get("imageio")
    sys.modules["imageio"] = fake
    try:
        spec = importlib.util.spec_from_file_location("_ref_util", "/root/reference/magicanimate/utils/util.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        out = {"cases": []}
        for i, (b, c, t, h, w, n_rows, rescale) in enumerate(VIDEO_GRID_CASES):
            v = torch.rand(b, c, t, h, w, generator=torch.Generator().manual_seed(900 + i))
            if rescale:
                v = v * 2 - 1
            with tempfile.TemporaryDirectory() as d:
                ref.save_videos_grid(v.clone(), f"{d}/x/y.mp4", rescale=rescale, n_rows=n_rows, fps=25)
            frames = torch.from_numpy(__import__("numpy").stack(rec["frames"]))
            if frames.dim() == 3:
                frames = frames[..., None]
            out["cases"].append({"shape": (b, c, t, h, w), "n_rows": n_rows, "rescale": rescale, "seed": 900 + i,
                                 "frames": frames, "fps": rec["fps"]})
        g = torch.Generator().manual_seed(950)
        v0, v1 = torch.randn(2, 4, 6, 6, generator=g), torch.randn(2, 4, 6, 6, generator=g)
        out["interp"] = {"v0": v0, "v1": v1, "t": [0.25, 0.5, 0.8],
                         "linear": [ref.linear(v0, v1, t) for t in (0.25, 0.5, 0.8)],
                         "slerp": [ref.slerp(v0, v1, t) for t in (0.25, 0.5, 0.8)],
                         "slerp_parallel": ref.slerp(v0, v0 * 1.5 + 1e-4 * v1, 0.3)}
        torch.save(out, GOLD / "video_grid.pt")
        print("video grid golden:", [tuple(c["frames"].shape) for c in out["cases"]])
    finally:
        if had is None:
            sys.modules.pop("imageio", None)
        else:
            sys.modules["imageio"] = had


def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    make_ddim_golden()
    make_speed_encoder_golden()
    make_video_grid_golden()
    make_cold_branch_golden()
    U = ref_shim.load_reference_unet_class()
    RC = ref_shim.load_reference_control_class()
    uniform = ref_shim.load_reference_context_uniform()

    # ---- key / shape lists (state_dict compatibility contract)
    m = U(**TINY_CFG).eval()
    tiny_shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    (GOLD / "unet3d_tiny_keys.json").write_text(json.dumps(tiny_shapes, indent=0, sort_keys=True))
    with torch.device("meta"):
        big = U(**FULL_CFG)
    full_shapes = {k: list(v.shape) for k, v in big.state_dict().items()}
    (GOLD / "unet3d_full_keys.json").write_text(json.dumps(full_shapes, indent=0, sort_keys=True))

    # ---- tiny network known answers
    sd = seeded_unet_state_dict(tiny_shapes, seed=0)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    out = {}
    with torch.no_grad():
        for tag, (b, f, hw) in {"a": (2, 4, 8), "b": (2, 8, 16)}.items():
            x, ctx = make_inputs(b, f, hw)
            out[f"plain_{tag}"] = m(x, torch.tensor(481), ctx).sample
        # per-frame (audio-token) context: batch == b*f, passes through un-repeated (attention.py:118-119)
        x, ctx = make_inputs(2, 4, 8, ctx_tokens=5, per_frame_ctx=True)
        out["per_frame_ctx"] = m(x, torch.tensor(21), ctx).sample
        # reference attention (reader hooks on mid+up blocks, CFG on)
        reader = RC(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup", batch_size=1)
        x, ctx = make_inputs(2, 4, 16)
        banks = make_banks(m, 16)
        mods = dict(m.named_modules())
        for name, tensors in banks.items():
            mods[name].bank = [t.clone() for t in tensors]
        out["with_banks"] = m(x, torch.tensor(301), ctx).sample
        out["without_banks"] = m(x, torch.tensor(301), ctx).sample  # banks were consumed
        # ControlNet residual inputs (unet_controlnet.py:414-448) and center_input_sample (:373-374)
        x, ctx = make_inputs(2, 2, 8)
        down, mid = controlnet_residuals()
        out["controlnet"] = m(x, torch.tensor(10), ctx, down_block_additional_residuals=tuple(down),
                              mid_block_additional_residual=mid).sample
        mc = U(**dict(TINY_CFG, center_input_sample=True)).eval()
        mc.load_state_dict(sd)
        out["center_input"] = mc(x, torch.tensor(10), ctx).sample
        # single blocks
        g = torch.Generator().manual_seed(3)
        h = torch.randn(2, 64, 4, 8, 8, generator=g)
        emb = torch.randn(2, 256, generator=g)
        ctx7 = torch.randn(2, 7, 64, generator=g)
        out["blk_h"], out["blk_emb"], out["blk_ctx"] = h, emb, ctx7
        out["blk_resnet"] = mods["down_blocks.1.resnets.0"](h, emb)
        out["blk_motion"] = mods["down_blocks.0.motion_modules.0"](h, None, None)
    # un-hook transformer blocks for the plain transformer golden (fresh model, same weights)
    m2 = U(**TINY_CFG).eval()
    m2.load_state_dict(sd)
    with torch.no_grad():
        out["blk_transformer"] = dict(m2.named_modules())["down_blocks.0.attentions.0"](h, encoder_hidden_states=ctx7).sample
    torch.save({k: v.contiguous() for k, v in out.items()}, GOLD / "unet3d_tiny_outputs.pt")

    # ---- context windows of the reference scheduler
    wins = {}
    for nf, cs, stride, ov in [(16, 16, 1, 4), (32, 16, 1, 4), (240, 16, 1, 4), (24, 8, 1, 2), (48, 16, 2, 4), (20, 16, 1, 0)]:
        wins[f"{nf},{cs},{stride},{ov}"] = [list(map(int, w)) for w in uniform(0, 50, nf, cs, stride, ov)]
    (GOLD / "context_windows.json").write_text(json.dumps(wins))
    # ---- ReferenceNet writer: the reference's own ReferenceAttentionControl(mode="write") on the reference UNet3D with one
    # frame and no motion modules (= the 2-D SD UNet the AppearanceEncoderModel is, unet_controlnet.py:485-525)
    wcfg = writer_cfg()
    mw = U(**wcfg).eval()
    wshapes = {k: list(v.shape) for k, v in mw.state_dict().items()}
    (GOLD / "writer_tiny_keys.json").write_text(json.dumps(wshapes, indent=0, sort_keys=True))
    mw.load_state_dict(seeded_unet_state_dict(wshapes, seed=3))
    RC(mw, do_classifier_free_guidance=True, mode="write", fusion_blocks="midup", batch_size=1)
    xw, ctxw = writer_inputs()
    with torch.no_grad():
        mw(xw[:, :, None], torch.tensor(441), ctxw)
    wmods = dict(mw.named_modules())
    torch.save({n: wmods[n].bank[0].contiguous() for n in reader_block_names(mw)}, GOLD / "writer_tiny_banks.pt")

    # ---- VAE decoder: the reference's own leaf modules wired in the published SD-VAE decoder order
    from oracle.vae_decoder import random_vae_decoder_state_dict
    vae_out = {}
    for tag, (widths, n, hw, seed) in VAE_CASES.items():
        vsd = random_vae_decoder_state_dict(block_out_channels=widths, seed=seed)
        dec = ref_shim.build_reference_vae_decoder(vsd, block_out_channels=widths)
        z = torch.randn(n, 4, hw, hw, generator=torch.Generator().manual_seed(100 + seed))
        vae_out[tag] = dec(z[:, :, None])[:, :, 0].contiguous()
    torch.save(vae_out, GOLD / "vae_decoder_outputs.pt")
    enc_out = {}
    for tag, (widths, n, hw, seed) in VAE_ENC_CASES.items():
        vsd = random_vae_decoder_state_dict(block_out_channels=widths, seed=seed)
        enc = ref_shim.build_reference_vae_encoder(vsd, block_out_channels=widths)
        img = torch.rand(n, 3, hw, hw, generator=torch.Generator().manual_seed(200 + seed)) * 2 - 1
        enc_out[tag] = enc(img[:, :, None])[:, :, 0].contiguous()
    torch.save(enc_out, GOLD / "vae_encoder_outputs.pt")
    print("golden written:", sorted(p.name for p in GOLD.iterdir()))


if __name__ == "__main__":
    if sys.argv[1:] == ["ddim"]:
        make_ddim_golden()
    elif sys.argv[1:] == ["speed"]:
        make_speed_encoder_golden()
    elif sys.argv[1:] == ["cold"]:
        make_cold_branch_golden()
    elif sys.argv[1:] == ["util"]:
        make_video_grid_golden()
    else:
        main()
