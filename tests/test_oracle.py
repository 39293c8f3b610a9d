"""Pins the oracle (oracle/*.py) — the checker every GPU parity test relies on.

(a) against golden vectors generated from the UNTOUCHED reference modules (oracle/make_golden.py, committed under
    tests/golden/), using weights re-derived from seeds;
(b) directly against the reference when /root/reference is present (build container only);
(c) DDIM restatement against the reference's own in-repo algebra (inversion round trip), window scheduler against the
    reference's `context.uniform` output.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from util_models import TINY_CFG, make_banks, make_inputs, rel_l2, seeded_unet_state_dict

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD / "unet3d_tiny_outputs.pt")


@pytest.fixture(scope="module")
def port():
    from oracle.unet3d_port import UNet3DOracle
    shapes = json.loads((GOLD / "unet3d_tiny_keys.json").read_text())
    return UNet3DOracle(seeded_unet_state_dict(shapes, 0), TINY_CFG)


def test_port_matches_reference_golden_outputs(port, gold):
    for tag, (b, f, hw) in {"a": (2, 4, 8), "b": (2, 8, 16)}.items():
        x, ctx = make_inputs(b, f, hw)
        assert rel_l2(port(x, 481, ctx), gold[f"plain_{tag}"]) < 1e-5
    x, ctx = make_inputs(2, 4, 8, ctx_tokens=5, per_frame_ctx=True)
    assert rel_l2(port(x, 21, ctx), gold["per_frame_ctx"]) < 1e-5


def test_port_reference_attention_golden(port, gold):
    from emote_hack_b200.unet3d import UNet3DConditionModel
    with torch.device("meta"):
        skeleton = UNet3DConditionModel(**TINY_CFG)
    banks = make_banks(skeleton, 16)
    x, ctx = make_inputs(2, 4, 16)
    assert rel_l2(port(x, 301, ctx, banks=banks), gold["with_banks"]) < 1e-5
    assert rel_l2(port(x, 301, ctx), gold["without_banks"]) < 1e-5
    assert rel_l2(gold["with_banks"], gold["without_banks"]) > 1e-2


def test_port_controlnet_residuals_and_center_input_golden(port, gold):
    """cold-but-signature branches of UNet3DConditionModel.forward the GPU tests rely on the oracle for"""
    from oracle.unet3d_port import UNet3DOracle
    from util_models import controlnet_residuals
    x, ctx = make_inputs(2, 2, 8)
    down, mid = controlnet_residuals()
    got = port(x, 10, ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
    assert rel_l2(got, gold["controlnet"]) < 1e-5
    assert rel_l2(port(x, 10, ctx), gold["controlnet"]) > 1e-3   # the residuals matter
    centred = UNet3DOracle(port.sd, dict(TINY_CFG, center_input_sample=True))(x, 10, ctx)
    assert rel_l2(centred, gold["center_input"]) < 1e-5


def test_port_blocks_golden(port, gold):
    h, emb, ctx = gold["blk_h"], gold["blk_emb"], gold["blk_ctx"]
    assert rel_l2(port._resnet("down_blocks.1.resnets.0", h, emb), gold["blk_resnet"]) < 1e-5
    assert rel_l2(port._motion("down_blocks.0.motion_modules.0", h), gold["blk_motion"]) < 1e-5
    assert rel_l2(port._transformer3d("down_blocks.0.attentions.0", h, ctx, 4, None, True), gold["blk_transformer"]) < 1e-5


def test_state_dict_keys_match_reference():
    """state_dict names/shapes are the checkpoint compatibility contract."""
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from util_models import FULL_CFG
    for cfg, fn in ((TINY_CFG, "unet3d_tiny_keys.json"), (FULL_CFG, "unet3d_full_keys.json")):
        want = json.loads((GOLD / fn).read_text())
        with torch.device("meta"):
            ours = UNet3DConditionModel(**cfg)
        got = {k: list(v.shape) for k, v in ours.state_dict().items()}
        assert got == want


def test_port_against_live_reference():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("/root/reference only exists in the build container")
    from oracle.unet3d_port import UNet3DOracle
    U = ref_shim.load_reference_unet_class()
    torch.manual_seed(5)
    m = U(**TINY_CFG).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "temporal_transformer.proj_out" in n:
                p.normal_(0, 0.05)
    x, ctx = make_inputs(2, 3, 8, seed=77)
    with torch.no_grad():
        ref = m(x, torch.tensor(700), ctx).sample
    assert rel_l2(UNet3DOracle(m.state_dict(), dict(m.config))(x, 700, ctx), ref) < 1e-5


def test_context_windows_match_reference():
    from emote_hack_b200.pipeline import uniform
    from oracle.ddim import uniform_windows
    wins = json.loads((GOLD / "context_windows.json").read_text())
    for key, want in wins.items():
        nf, cs, stride, ov = map(int, key.split(","))
        assert uniform_windows(0, 50, nf, cs, stride, ov) == want
        assert [list(map(int, w)) for w in uniform(0, 50, nf, cs, stride, ov)] == want


def test_ddim_restatement():
    from emote_hack_b200.pipeline import DDIMScheduler
    from oracle.ddim import DDIMOracle
    o = DDIMOracle()
    ts = o.set_timesteps(50)
    assert ts[0] == 981 and ts[-1] == 1 and len(ts) == 50
    s = DDIMScheduler()
    s.set_timesteps(50)
    assert s.timesteps.tolist() == ts.tolist()
    for t in (981, 501, 1):
        a, b = o.alphas(t), s.alphas_for(t)
        assert abs(a[0] - b[0]) < 1e-6 and abs(a[1] - b[1]) < 1e-6
    assert o.alphas(1)[1] == 1.0  # final alpha_cumprod = 1
    # round trip against the reference's own inversion algebra (EMOAnimationPipeline.py:379-400): with the same eps,
    # next_step (x_{t-20} -> x_t) followed by step (x_t -> x_{t-20}) is the identity
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 2, 8, 8, generator=g)
    eps = torch.randn(1, 4, 2, 8, 8, generator=g)
    for t in (981, 481, 21):
        up = o.ddim_inversion_step(eps, t, x)
        back = o.step(eps, t, up)
        assert rel_l2(back, x) < 1e-5


def test_ddim_oracle_pinned_on_executed_reference_code():
    """oracle/ddim.py against vectors produced by EXECUTING the reference's own DDIM algebra
    (EMOAnimationPipeline.next_step :379-400 == magicanimate/utils/util.py:64-74, cut out with ast by oracle/ref_ddim.py):
    the inversion step as written, and the forward `scheduler.step` = the same code with the two alphas exchanged."""
    from oracle.ddim import DDIMOracle
    gold = torch.load(GOLD / "ddim_reference_steps.pt")
    x, eps = gold["x"], gold["eps"]
    for n, t in gold["cases"]:
        o = DDIMOracle()
        o.set_timesteps(n)
        assert rel_l2(o.ddim_inversion_step(eps, t, x), gold[f"invert_{n}_{t}"]) < 1e-6
        assert rel_l2(o.step(eps, t, x), gold[f"step_{n}_{t}"]) < 1e-6
        a_from = o.alphas_cumprod[t - 1000 // n] if t - 1000 // n >= 0 else 1.0
        x0 = (x - (1 - a_from) ** 0.5 * eps) / a_from ** 0.5
        assert rel_l2(x0, gold[f"x0_{n}_{t}"]) < 1e-6


def test_ddim_oracle_against_live_reference_code():
    """same check against the reference sources as they lie in /root/reference (no committed vectors involved)"""
    if not Path("/root/reference/EMOAnimationPipeline.py").exists():
        pytest.skip("/root/reference only exists in the build container")
    import types
    from oracle import ref_ddim
    from oracle.ddim import DDIMOracle
    method, util_fn = ref_ddim.reference_next_step_method(), ref_ddim.reference_next_step_util()
    g = torch.Generator().manual_seed(5)
    x, eps = torch.randn(2, 4, 3, 8, 8, generator=g), torch.randn(2, 4, 3, 8, 8, generator=g)
    for n in (50, 25, 20):
        o = DDIMOracle()
        for t in o.set_timesteps(n).tolist():
            stub = ref_ddim.scheduler_stub(o.alphas_cumprod, n)
            want, _ = method(types.SimpleNamespace(scheduler=stub), eps, t, x)
            assert torch.equal(want, util_fn(eps, t, x, stub))
            assert rel_l2(o.ddim_inversion_step(eps, t, x), want) < 1e-6
            fwd = ref_ddim.swapped_alpha_stub(o.alphas_cumprod, t, n)
            assert rel_l2(o.step(eps, t, x), method(types.SimpleNamespace(scheduler=fwd), eps, t, x)[0]) < 1e-6


def test_port_cold_branches_match_reference_golden():
    """forced upsample size (latent not a multiple of 8) and the three class-embedding types, against outputs recorded
    from the reference UNet3DConditionModel (oracle/make_golden.py cold)"""
    from oracle.unet3d_port import UNet3DOracle
    gold = torch.load(GOLD / "unet3d_tiny_cold_branches.pt")
    for tag, extra in (("odd_size", {}), ("class_table", {"num_class_embeds": 4}),
                       ("class_timestep", {"class_embed_type": "timestep"}), ("class_identity", {"class_embed_type": "identity"})):
        o = UNet3DOracle(seeded_unet_state_dict(gold[f"{tag}_keys"], 4), dict(TINY_CFG, **extra))
        x, ctx = make_inputs(2, 2, 10 if tag == "odd_size" else 8, seed=77)
        got = o(x, 301, ctx, class_labels=gold[f"{tag}_labels"])
        assert got.shape == gold[tag].shape and rel_l2(got, gold[tag]) < 1e-5, tag


def test_speed_encoder_restatement_pinned_on_reference_class():
    """oracle/ref_audio.speed_encoder_restated against vectors recorded by executing the reference's SpeedEncoder
    (Net.py:198-258, cut out with ast) — and against the live class when /root/reference is present"""
    from oracle import ref_audio
    gold = torch.load(GOLD / "speed_encoder.pt")
    sd = gold["state_dict"]
    got = ref_audio.speed_encoder_restated(gold["speeds"], gold["centers"], gold["radii"], sd["mlp.0.weight"], sd["mlp.0.bias"],
                                           sd["mlp.2.weight"], sd["mlp.2.bias"])
    assert rel_l2(got, gold["out"]) < 1e-6
    if ref_audio.REFERENCE_NET.exists():
        SE = ref_audio.reference_speed_encoder_class()
        enc = SE(9, 64)
        enc.load_state_dict(sd)
        with torch.no_grad():
            assert torch.equal(enc(gold["speeds"]), gold["out"])
        with pytest.raises(AssertionError):
            SE(10, 64)          # 9 hard-coded bucket centres (Net.py:225-229): what the pipeline's SpeedEncoder(10, 64) hits


def test_wav2vec2_module_tree_matches_transformers_state_dict():
    """boundary of the audio front-end: same parameter names / shapes as transformers.Wav2Vec2Model (wav2vec2-base)"""
    from emote_hack_b200.audio import Wav2Vec2Model
    from oracle import ref_audio
    hf = ref_audio.hf_wav2vec2()
    with torch.device("meta"):
        ours = Wav2Vec2Model()
    a = {k: tuple(v.shape) for k, v in hf.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert a == b and len(a) == 211
    assert ours.frames_for(16000) == 49 and ours.frames_for(160000) == 499


def test_oracle_writer_banks_match_reference_golden():
    """ReferenceNet writer: oracle run with collect_banks against banks recorded from the reference's own
    ReferenceAttentionControl(mode="write") on its UNet3D (one frame, no motion modules)."""
    from oracle.unet3d_port import UNet3DOracle
    from util_models import APPEARANCE_TRIMMED, reader_block_names, writer_cfg, writer_inputs
    shapes = json.loads((GOLD / "writer_tiny_keys.json").read_text())
    gold_banks = torch.load(GOLD / "writer_tiny_banks.pt")
    sd = seeded_unet_state_dict(shapes, 3)
    x, ctx = writer_inputs()
    for trimmed in (False, True):   # full 2-D UNet weights, and the AppearanceEncoderModel subset of them
        use = {k: v for k, v in sd.items() if not (trimmed and k.startswith(APPEARANCE_TRIMMED))}
        banks = {}
        UNet3DOracle(use, writer_cfg())(x[:, :, None], 441, ctx, collect_banks=banks)
        assert len(gold_banks) == 10
        for name, want in gold_banks.items():
            assert len(banks[name]) == 1 and rel_l2(banks[name][0], want) < 1e-5, name


def test_appearance_encoder_state_dict_contract():
    """AppearanceEncoderModel keys == the 2-D SD UNet keys (reference UNet3D without motion modules) minus the tail the
    reference strips (appearance_encoder.py:613-621 and the missing conv_norm_out / conv_out)."""
    from emote_hack_b200.appearance_encoder import AppearanceEncoderModel
    from util_models import APPEARANCE_TRIMMED, appearance_cfg
    shapes = json.loads((GOLD / "writer_tiny_keys.json").read_text())
    want = {k: v for k, v in shapes.items() if not k.startswith(APPEARANCE_TRIMMED)}
    with torch.device("meta"):
        ours = AppearanceEncoderModel(**appearance_cfg())
    got = {k: list(v.shape) for k, v in ours.state_dict().items()}
    assert got == want and len(shapes) - len(want) == 24
    with pytest.raises(NotImplementedError):
        AppearanceEncoderModel(addition_embed_type="text", **appearance_cfg())
    with pytest.raises(TypeError):
        AppearanceEncoderModel(not_an_option=1, **appearance_cfg())


VAE_CASES = {"tiny": ((32, 32, 64, 64), 2, 8, 1), "mid": ((64, 128, 128, 128), 1, 16, 2)}  # = oracle/make_golden.py


def test_vae_oracle_matches_reference_leaf_golden():
    """oracle/vae_decoder.py against the decoder wired out of the reference's own ResnetBlock3D / AttentionBlock /
    Upsample3D / InflatedConv3d (oracle/ref_shim.build_reference_vae_decoder), recorded in tests/golden/."""
    from oracle.vae_decoder import VAEDecoderOracle, random_vae_decoder_state_dict
    gold = torch.load(GOLD / "vae_decoder_outputs.pt")
    for tag, (widths, n, hw, seed) in VAE_CASES.items():
        sd = random_vae_decoder_state_dict(block_out_channels=widths, seed=seed)
        z = torch.randn(n, 4, hw, hw, generator=torch.Generator().manual_seed(100 + seed))
        got = VAEDecoderOracle(sd).decode(z)
        assert got.shape == gold[tag].shape == (n, 3, 8 * hw, 8 * hw)
        assert rel_l2(got, gold[tag]) < 1e-5


VAE_ENC_CASES = {"tiny": ((32, 32, 64, 64), 2, 32, 1), "mid": ((64, 128, 128, 128), 1, 64, 2)}  # = oracle/make_golden.py


def test_vae_encoder_oracle_matches_reference_leaf_golden():
    """oracle encode() (images2latents, EMOAnimationPipeline.py:402-414) against the encoder wired out of the reference's
    leaf modules (oracle/ref_shim.build_reference_vae_encoder)."""
    from oracle.vae_decoder import VAEDecoderOracle, random_vae_decoder_state_dict
    gold = torch.load(GOLD / "vae_encoder_outputs.pt")
    for tag, (widths, n, hw, seed) in VAE_ENC_CASES.items():
        sd = random_vae_decoder_state_dict(block_out_channels=widths, seed=seed)
        img = torch.rand(n, 3, hw, hw, generator=torch.Generator().manual_seed(200 + seed)) * 2 - 1
        got = VAEDecoderOracle(sd).encode(img)
        assert got.shape == gold[tag].shape == (n, 8, hw // 8, hw // 8)
        assert rel_l2(got, gold[tag]) < 1e-5
    o = VAEDecoderOracle(random_vae_decoder_state_dict(block_out_channels=(32, 32, 64, 64), seed=1))
    u8 = torch.randint(0, 256, (2, 16, 16, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    lat = o.images2latents(u8)
    assert lat.shape == (2, 4, 2, 2) and torch.isfinite(lat).all()


def test_vae_oracle_against_live_reference_leaves():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("/root/reference only exists in the build container")
    from oracle.vae_decoder import VAEDecoderOracle, random_vae_decoder_state_dict
    widths = (32, 64, 64, 96)
    sd = random_vae_decoder_state_dict(block_out_channels=widths, seed=9)
    z = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(9))
    ref = ref_shim.build_reference_vae_decoder(sd, block_out_channels=widths)(z[:, :, None])[:, :, 0]
    assert rel_l2(VAEDecoderOracle(sd).decode(z), ref) < 1e-5


def test_vae_oracle_shapes():
    from oracle.vae_decoder import VAEDecoderOracle, random_vae_decoder_state_dict
    sd = random_vae_decoder_state_dict(block_out_channels=(32, 32, 64, 64), seed=1)
    o = VAEDecoderOracle(sd)
    lat = torch.randn(1, 4, 2, 4, 4)
    v = o.decode_latents(lat)
    assert v.shape == (1, 3, 2, 32, 32) and float(v.min()) >= 0 and float(v.max()) <= 1
    # legacy and new attention key names are interchangeable
    sd2 = {k.replace(".query.", ".to_q.").replace(".key.", ".to_k.").replace(".value.", ".to_v.").replace(".proj_attn.", ".to_out.0."): v
           for k, v in sd.items()}
    assert torch.equal(VAEDecoderOracle(sd2).decode_latents(lat), v)


def _video_case_input(case):
    v = torch.rand(*case["shape"], generator=torch.Generator().manual_seed(case["seed"]))
    return v * 2 - 1 if case["rescale"] else v


def test_video_grid_restatement_pinned_on_reference_frames():
    """oracle/video_grid.py against the frames the reference's save_videos_grid (utils/util.py:21-33) handed to its writer,
    and against torchvision.make_grid itself"""
    import torchvision
    from oracle import video_grid
    gold = torch.load(GOLD / "video_grid.pt")
    for case in gold["cases"]:
        v = _video_case_input(case)
        got = video_grid.video_frames_u8(v.numpy(), case["rescale"], case["n_rows"])
        assert got.shape == tuple(case["frames"].shape) and np.array_equal(got, case["frames"].numpy()), case["shape"]
        assert case["fps"] == 25
        tv = torchvision.utils.make_grid(v[:, :, 0], nrow=case["n_rows"]).numpy()
        assert np.array_equal(video_grid.make_grid(v[:, :, 0].numpy(), case["n_rows"]), tv)
    it = gold["interp"]
    for i, t in enumerate(it["t"]):
        assert np.allclose(video_grid.linear(it["v0"].numpy(), it["v1"].numpy(), t), it["linear"][i].numpy(), atol=1e-6)
        assert np.allclose(video_grid.slerp(it["v0"].numpy(), it["v1"].numpy(), t), it["slerp"][i].numpy(), atol=1e-5)


def test_latent_interpolation_helpers_and_frame_files(tmp_path):
    """host side of the frames -> file step: linear / slerp (utils/util.py:125-141) against the reference's outputs,
    `interpolate_latents` (EMOAnimationPipeline.py:479-512) against a direct loop, and the writer / reader pair"""
    from emote_hack_b200.magicanimate.utils import util
    from emote_hack_b200.pipeline import EMOAnimationPipeline
    it = torch.load(GOLD / "video_grid.pt")["interp"]
    for i, t in enumerate(it["t"]):
        assert torch.allclose(util.linear(it["v0"], it["v1"], t), it["linear"][i], atol=1e-6)
        assert torch.allclose(util.slerp(it["v0"], it["v1"], t), it["slerp"][i], atol=1e-5)
    assert torch.allclose(util.slerp(it["v0"], it["v0"] * 1.5 + 1e-4 * it["v1"], 0.3), it["slerp_parallel"], atol=1e-6)

    pipe = EMOAnimationPipeline.__new__(EMOAnimationPipeline)   # interpolate_latents uses no state
    lat = torch.randn(1, 4, 3, 2, 2, generator=torch.Generator().manual_seed(3))
    assert pipe.interpolate_latents(lat, 1) is lat              # the reference's fixed factor (:825): no-op
    for is_slerp in (False, True):
        util.set_tensor_interpolation_method(is_slerp)
        out = pipe.interpolate_latents(lat, 3)
        assert out.shape == (1, 4, 7, 2, 2)
        fn = util.slerp if is_slerp else util.linear
        assert torch.equal(out[:, :, 0], lat[:, :, 0]) and torch.equal(out[:, :, 3], lat[:, :, 1]) and torch.equal(out[:, :, 6], lat[:, :, 2])
        assert torch.allclose(out[:, :, 4], fn(lat[:, :, 1], lat[:, :, 2], 1 / 3), atol=1e-6)
    util.set_tensor_interpolation_method(False)

    frames = (np.random.default_rng(0).random((5, 16, 24, 3)) * 255).astype(np.uint8)
    for name in ("a/clip.npy", "clip.gif", "clip.avi"):
        util.images2video(list(frames), str(tmp_path / name), fps=8)
        assert (tmp_path / name).stat().st_size > 0
    back = util.video2images(str(tmp_path / "a/clip.npy"), step=2, length=2, start=1)
    assert len(back) == 2 and np.array_equal(back[0], frames[1]) and np.array_equal(back[1], frames[3])
    avi = util.video2images(str(tmp_path / "clip.avi"), step=1, length=16)
    assert len(avi) == 5 and avi[0].shape == (16, 24, 3)      # MJPG is lossy: only the geometry is checked
