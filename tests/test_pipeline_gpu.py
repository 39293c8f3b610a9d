"""VAE decoder, scheduler step and the denoise loop on the GPU against the oracle restatements."""
import json
from pathlib import Path

import pytest
import torch

from util_models import (TINY_CFG, check_parity, make_banks, make_inputs, rel_l2, rerandomise_zero_inits,
                         seeded_unet_state_dict)

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def test_cuda_unet_against_reference_golden_vectors():
    """CUDA path vs outputs recorded from the untouched reference (same seeded weights), bf16-operand tolerance."""
    from emote_hack_b200.unet3d import ReferenceAttentionControl, UNet3DConditionModel
    gold = torch.load(GOLD / "unet3d_tiny_outputs.pt")
    shapes = json.loads((GOLD / "unet3d_tiny_keys.json").read_text())
    m = UNet3DConditionModel(**TINY_CFG).eval()
    m.load_state_dict(seeded_unet_state_dict(shapes, 0), strict=True)
    m = m.cuda()
    for tag, (b, f, hw) in {"a": (2, 4, 8), "b": (2, 8, 16)}.items():
        x, ctx = make_inputs(b, f, hw)
        e = rel_l2(m(x.cuda(), 481, ctx.cuda()).sample, gold[f"plain_{tag}"])
        check_parity(f"golden.plain_{tag}", e, 2e-2)
    x, ctx = make_inputs(2, 4, 8, ctx_tokens=5, per_frame_ctx=True)
    check_parity("golden.per_frame_ctx", rel_l2(m(x.cuda(), 21, ctx.cuda()).sample, gold["per_frame_ctx"]), 2e-2)
    reader = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup")
    reader.set_banks({k: [t.cuda() for t in v] for k, v in make_banks(m, 16).items()})
    x, ctx = make_inputs(2, 4, 16)
    e = rel_l2(m(x.cuda(), 301, ctx.cuda()).sample, gold["with_banks"])
    check_parity("golden.with_banks", e, 2e-2)
    check_parity("golden.without_banks", rel_l2(m(x.cuda(), 301, ctx.cuda()).sample, gold["without_banks"]), 2e-2)
    h, emb, c7 = gold["blk_h"].cuda(), gold["blk_emb"].cuda(), gold["blk_ctx"].cuda()
    mods = dict(m.named_modules())
    check_parity("golden.blk_resnet", rel_l2(mods["down_blocks.1.resnets.0"](h, emb), gold["blk_resnet"]), 5e-3)
    check_parity("golden.blk_motion", rel_l2(mods["down_blocks.0.motion_modules.0"](h, None, None), gold["blk_motion"]), 5e-3)
    for blk in reader._blocks(m):
        blk._ref_mode = None
    check_parity("golden.blk_transformer",
                 rel_l2(mods["down_blocks.0.attentions.0"](h, encoder_hidden_states=c7).sample, gold["blk_transformer"]), 5e-3)


def test_cuda_unet_cold_branches_against_reference_golden_vectors():
    """forced `upsample_size` (10x10 latent: 10 -> 5 -> 3 -> 2 and back) and the class-embedding variants of
    UNet3DConditionModel.forward (unet_controlnet.py:121-128, 355-364, 400-408, 458-460) vs the reference's outputs"""
    from emote_hack_b200.unet3d import UNet3DConditionModel
    gold = torch.load(GOLD / "unet3d_tiny_cold_branches.pt")
    for tag, extra in (("odd_size", {}), ("class_table", {"num_class_embeds": 4}),
                       ("class_timestep", {"class_embed_type": "timestep"}), ("class_identity", {"class_embed_type": "identity"})):
        m = UNet3DConditionModel(**dict(TINY_CFG, **extra)).eval()
        m.load_state_dict(seeded_unet_state_dict(gold[f"{tag}_keys"], 4), strict=True)
        m = m.cuda()
        x, ctx = make_inputs(2, 2, 10 if tag == "odd_size" else 8, seed=77)
        labels = gold[f"{tag}_labels"]
        out = m(x.cuda(), 301, ctx.cuda(), class_labels=None if labels is None else labels.cuda()).sample
        assert out.shape == gold[tag].shape
        check_parity(f"golden.cold.{tag}", rel_l2(out, gold[tag]), 2e-2)
    with pytest.raises(ValueError):
        m(x.cuda(), 301, ctx.cuda())                       # class embedding configured, no labels


@pytest.fixture(scope="module")
def tiny_vae():
    from emote_hack_b200.vae import AutoencoderKL
    from oracle.vae_decoder import VAEDecoderOracle
    torch.manual_seed(1)
    vae = AutoencoderKL(block_out_channels=(64, 64, 128, 128)).eval()
    return VAEDecoderOracle(vae.state_dict()), vae.cuda()


def test_vae_decode(tiny_vae):
    oracle, vae = tiny_vae
    z = torch.randn(3, 4, 8, 8, generator=torch.Generator().manual_seed(4))
    ref = oracle.decode(z)
    out = vae.decode(z.cuda()).sample
    assert out.shape == ref.shape == (3, 3, 64, 64)
    e = rel_l2(out, ref)
    check_parity("vae.decode_tiny", e, 2e-2)


def test_vae_decode_video_and_state_dict_aliases(tiny_vae):
    from emote_hack_b200.vae import AutoencoderKL
    oracle, vae = tiny_vae
    lat = torch.randn(1, 4, 3, 8, 8, generator=torch.Generator().manual_seed(5)) * 0.18215 * 2
    ref = oracle.decode_latents(lat)
    vid, u8 = vae.decode_video(lat.cuda(), want_u8=True)
    assert vid.shape == ref.shape == (1, 3, 3, 64, 64)
    e, emax = rel_l2(vid, ref), (vid.cpu() - ref).abs().max().item()
    print(f"vae decode_video rel_l2={e:.2e} max_abs={emax:.2e}")
    check_parity("vae.decode_video_tiny", e, 2e-2)
    check_parity("vae.decode_video_tiny_maxabs", emax, 6e-2)   # [0,1] images
    assert torch.equal(u8, (vid * 255).to(torch.uint8))   # uint8 copy = the reference's truncating cast (utils/util.py:28)
    # chunked decode == batched decode; newer diffusers attention key names load
    vid2, _ = vae.decode_video(lat.cuda(), frame_chunk=2)
    assert rel_l2(vid2, vid) < 1e-5
    sd = {k.replace(".query.", ".to_q.").replace(".key.", ".to_k.").replace(".value.", ".to_v.").replace(".proj_attn.", ".to_out.0."): v
          for k, v in vae.state_dict().items()}
    vae2 = AutoencoderKL(block_out_channels=(64, 64, 128, 128)).eval()
    vae2.load_state_dict(sd)
    assert rel_l2(vae2.cuda().decode_video(lat.cuda())[0], vid) < 1e-5


def test_vae_encode_and_images2latents(tiny_vae):
    """`vae.encode(x)['latent_dist'].mean * 0.18215` of images2latents (EMOAnimationPipeline.py:402-414) on CUDA vs the
    oracle encoder (pinned on the reference's leaf modules, tests/test_oracle.py)."""
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    oracle, vae = tiny_vae
    img = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(6)) * 2 - 1
    ref = oracle.encode(img)
    out = vae.encode(img.cuda())
    dist = out["latent_dist"]
    assert dist is out.latent_dist and dist.mean.shape == (2, 4, 8, 8)
    e = rel_l2(torch.cat([dist.mean, dist.logvar], 1), ref.clamp_(min=-1e9))
    check_parity("vae.encode_moments", e, 2e-2)
    check_parity("vae.encode_mean", rel_l2(dist.mode(), ref[:, :4]), 2e-2)
    assert dist.sample().shape == (2, 4, 8, 8)
    # decoder-only checkpoints stay loadable (the encoder keeps its weights)
    missing = vae.load_state_dict({k: v for k, v in vae.state_dict().items() if not k.startswith(("encoder.", "quant_conv."))})
    assert not missing.missing_keys
    u8 = torch.randint(0, 256, (3, 32, 32, 3), generator=torch.Generator().manual_seed(7), dtype=torch.uint8)
    pipe = EMOAnimationPipeline(vae, None, DDIMScheduler())
    lat = pipe.images2latents(u8.numpy(), torch.float32)
    assert lat.shape == (3, 4, 4, 4)
    check_parity("vae.images2latents", rel_l2(lat, oracle.images2latents(u8)), 2e-2)
    with pytest.raises(ValueError):
        vae.encode(torch.zeros(1, 3, 20, 20, device="cuda"))


def test_scheduler_step_matches_oracle():
    from emote_hack_b200.pipeline import DDIMScheduler
    from oracle.ddim import DDIMOracle
    s, o = DDIMScheduler(), DDIMOracle()
    s.set_timesteps(50), o.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x, eps = torch.randn(1, 4, 3, 8, 8, generator=g), torch.randn(1, 4, 3, 8, 8, generator=g)
    for t in (981, 481, 1):
        got = s.step(eps.cuda(), t, x.cuda()).prev_sample
        assert rel_l2(got, o.step(eps, t, x)) < 1e-5
    assert s.init_noise_sigma == 1.0 and s.scale_model_input(x, 3) is x


def _oracle_denoise(unet_o, latents, ctx, steps, guidance, windows, banks=None, per_frame=False):
    """restates the reference loop EMOAnimationPipeline.py:698-823 (single process) on the oracle pieces"""
    from oracle.ddim import DDIMOracle, cfg_combine
    sch = DDIMOracle()
    f_total = latents.shape[2]
    lat = latents.clone()
    for t in sch.set_timesteps(steps).tolist():
        acc = torch.zeros(2, *lat.shape[1:])
        cnt = torch.zeros(1, 1, f_total, 1, 1)
        for c in windows:
            x = lat[:, :, c].repeat(2, 1, 1, 1, 1)
            cc = torch.cat([ctx[:f_total][c], ctx[f_total:][c]]) if per_frame else ctx
            acc[:, :, c] += unet_o(x, t, cc, banks=banks)
            cnt[:, :, c] += 1
        lat = sch.step(cfg_combine(acc, cnt, guidance), t, lat)
    return lat


@pytest.fixture(scope="module")
def tiny_pipe():
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from oracle.unet3d_port import UNet3DOracle
    torch.manual_seed(0)
    m = rerandomise_zero_inits(UNet3DConditionModel(**TINY_CFG).eval())
    o = UNet3DOracle(m.state_dict(), dict(m.config))
    return o, EMOAnimationPipeline(None, m.cuda(), DDIMScheduler())


def test_denoise_single_window(tiny_pipe):
    o, pipe = tiny_pipe
    g = torch.Generator().manual_seed(11)
    lat = torch.randn(1, 4, 4, 8, 8, generator=g)
    ctx = torch.randn(2, 7, 64, generator=g)
    ref = _oracle_denoise(o, lat, ctx, 3, 7.5, [list(range(4))])
    out = pipe.denoise(lat.cuda(), ctx.cuda(), num_inference_steps=3, guidance_scale=7.5, context_frames=16)
    e = rel_l2(out, ref)
    check_parity("denoise.single_window_3steps", e, 3e-2)


def test_ddim_inversion_matches_oracle(tiny_pipe):
    """`invert` / `next_step` (EMOAnimationPipeline.py:379-477): deterministic DDIM inversion of frame latents, against
    the same loop on the oracle UNet + the oracle's inversion step (itself a restatement of the reference's next_step)."""
    from oracle.ddim import DDIMOracle
    o, pipe = tiny_pipe
    g = torch.Generator().manual_seed(17)
    lat = torch.randn(3, 4, 8, 8, generator=g)           # f c h w
    emb = torch.randn(1, 7, 64, generator=g)
    sch = DDIMOracle()
    ts = sch.set_timesteps(5).tolist()
    ref = lat.clone()
    for i, t in enumerate(reversed(ts)):
        if i >= 3:
            continue
        eps = o(ref.permute(1, 0, 2, 3)[None], t, emb)[0].permute(1, 0, 2, 3)
        ref = sch.ddim_inversion_step(eps, t, ref)
    out, inter = pipe.invert(lat.cuda(), emb.cuda(), num_inference_steps=5, num_actual_inference_steps=3,
                             return_intermediates=True)
    assert len(inter) == 4 and out.shape == lat.shape
    e = rel_l2(out, ref)
    check_parity("denoise.inversion_3of5", e, 3e-2)
    # next_step followed by scheduler.step with the same epsilon is the identity (round trip of the two updates)
    eps = torch.randn(3, 4, 8, 8, generator=g).cuda()
    up, x0 = pipe.next_step(eps, 401, lat.cuda())
    back = pipe.scheduler.step(eps, 401, up).prev_sample
    assert rel_l2(back, lat) < 1e-5
    a_cur = float(pipe.scheduler.alphas_cumprod[201])
    assert rel_l2(x0, (lat.cuda() - (1 - a_cur) ** 0.5 * eps) / a_cur ** 0.5) < 1e-5
    with pytest.raises(TypeError):
        pipe.invert(lat.cuda(), "a prompt")


def test_cached_graph_is_rebuilt_after_weight_update(tiny_pipe):
    """The CUDA graph of the UNet step reads packed weight copies made at capture time: a load_state_dict between two
    denoise() calls must invalidate it (same inputs, new weights -> the eager result of the new weights)."""
    o, pipe = tiny_pipe
    g = torch.Generator().manual_seed(13)
    lat = torch.randn(1, 4, 4, 8, 8, generator=g).cuda()
    ctx = torch.randn(2, 7, 64, generator=g).cuda()
    kw = dict(num_inference_steps=2, guidance_scale=7.5, context_frames=16)
    before = pipe.denoise(lat.clone(), ctx, **kw)
    sd = {k: v.clone() for k, v in pipe.unet.state_dict().items()}
    try:
        with torch.no_grad():
            pipe.unet.conv_out.weight.mul_(0.5)
            pipe.unet.conv_out.bias.mul_(0.5)
        graph_after = pipe.denoise(lat.clone(), ctx, **kw)
        eager_after = pipe.denoise(lat.clone(), ctx, use_cuda_graph=False, **kw)
        assert rel_l2(graph_after, eager_after) < 1e-5
        assert rel_l2(graph_after, before) > 1e-3
    finally:
        pipe.unet.load_state_dict(sd)
    assert rel_l2(pipe.denoise(lat.clone(), ctx, **kw), before) < 1e-5


def test_graphed_single_branch_units_and_bank_refresh(tiny_pipe):
    """GraphedUNet in the three unit modes (pair / uncond / cond): the captured batch-1 branches reproduce their eager call,
    read the right bank row, and `refresh` swaps conditioning without a re-capture (what ranks do when a window's CFG
    branches are dealt to different GPUs)."""
    from emote_hack_b200 import ops
    from emote_hack_b200.pipeline import GraphedUNet, _run_unet
    o, pipe = tiny_pipe
    m = pipe.unet
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 4, 4, 8, 8, generator=g).cuda()
    ctx = torch.randn(2, 7, 64, generator=g).cuda()
    banks = {k: [v.cuda() for v in vs] for k, vs in make_banks(m, 8).items()}
    banks2 = {k: [v.cuda() for v in vs] for k, vs in make_banks(m, 8, seed=8).items()}
    t = torch.tensor([441.0], device="cuda")
    for mode, sl in (("pair", slice(0, 2)), ("uncond", slice(0, 1)), ("cond", slice(1, 2))):
        gr = GraphedUNet(m, (sl.stop - sl.start, 4, 4, 8, 8), ctx[sl], banks, x.device, mode=mode)
        for bk in (banks, banks2):
            gr.refresh(ctx[sl], bk)
            gr.lat.copy_(x[sl])
            got = gr.replay(441).clone()
            use = None if mode == "uncond" else {k: [v[0] if mode == "pair" else v[0][1:2]] for k, v in bk.items()}
            want = _run_unet(m, mode, x[sl].contiguous(), t, ctx[sl], use)
            assert rel_l2(got, want) < 1e-5, mode
    assert all(blk._ref_mode is None for blk in m.modules() if hasattr(blk, "_ref_mode"))


def test_denoise_sliding_windows_with_overlap_and_banks(tiny_pipe):
    """24 frames, windows of 8 with overlap 2 (closed loop): visit-count averaging, per-window reference banks"""
    from emote_hack_b200.pipeline import uniform
    o, pipe = tiny_pipe
    g = torch.Generator().manual_seed(12)
    lat = torch.randn(1, 4, 24, 8, 8, generator=g)
    ctx = torch.randn(2, 7, 64, generator=g)
    wins = list(uniform(0, 2, 24, 8, 1, 2))
    assert len(wins) == 4 and max(max(w) for w in wins) == 23
    banks = make_banks(pipe.unet, 8)
    ref = _oracle_denoise(o, lat, ctx, 2, 7.5, wins, banks=banks)
    out = pipe.denoise(lat.cuda(), ctx.cuda(), num_inference_steps=2, guidance_scale=7.5, context_frames=8,
                       context_overlap=2, reference_banks={k: [t.cuda() for t in v] for k, v in banks.items()})
    e = rel_l2(out, ref)
    check_parity("denoise.windows_overlap_banks", e, 3e-2)


@pytest.mark.parametrize("use_graph", [True, False])
def test_denoise_with_appearance_encoder(tiny_pipe, use_graph):
    """ReferenceNet writer inside the loop (EMOAnimationPipeline.py:711-716, 774, 788, 823): the writer runs once per
    timestep on the reference-image latents, its banks condition every window of that step; checked against the same
    loop on the oracle (writer banks from the oracle's writer mode).  Graph path = GraphedWriter + aliased banks."""
    import json
    from pathlib import Path
    from emote_hack_b200.appearance_encoder import AppearanceEncoderModel
    from emote_hack_b200.pipeline import uniform
    from oracle.ddim import DDIMOracle, cfg_combine
    from oracle.unet3d_port import UNet3DOracle
    from util_models import (APPEARANCE_TRIMMED, appearance_cfg, reader_block_names, seeded_unet_state_dict, writer_cfg)
    o, pipe = tiny_pipe
    shapes = json.loads((Path(__file__).parent / "golden" / "writer_tiny_keys.json").read_text())
    sd = {k: v for k, v in seeded_unet_state_dict(shapes, 3).items() if not k.startswith(APPEARANCE_TRIMMED)}
    enc = AppearanceEncoderModel(**appearance_cfg()).eval()
    enc.load_state_dict(sd, strict=True)
    enc = enc.cuda()
    wo = UNet3DOracle(sd, writer_cfg())
    g = torch.Generator().manual_seed(21)
    lat = torch.randn(1, 4, 12, 8, 8, generator=g)
    ctx = torch.randn(2, 7, 64, generator=g)
    ref_lat = torch.randn(1, 4, 8, 8, generator=g)
    wins = list(uniform(0, 2, 12, 8, 1, 2))
    names = reader_block_names(pipe.unet)
    # oracle loop
    sch = DDIMOracle()
    ref = lat.clone()
    for t in sch.set_timesteps(2).tolist():
        banks = {}
        wo(ref_lat.repeat(2, 1, 1, 1)[:, :, None], t, ctx, collect_banks=banks)
        banks = {n: banks[n] for n in names}
        acc = torch.zeros(2, *ref.shape[1:])
        cnt = torch.zeros(1, 1, 12, 1, 1)
        for c in wins:
            acc[:, :, c] += o(ref[:, :, c].repeat(2, 1, 1, 1, 1), t, ctx, banks=banks)
            cnt[:, :, c] += 1
        ref = sch.step(cfg_combine(acc, cnt, 7.5), t, ref)
    out = pipe.denoise(lat.cuda(), ctx.cuda(), num_inference_steps=2, guidance_scale=7.5, context_frames=8,
                       context_overlap=2, appearance_encoder=enc, ref_image_latents=ref_lat.cuda(),
                       use_cuda_graph=use_graph)
    e = rel_l2(out, ref)
    check_parity(f"denoise.appearance_encoder_graph{int(use_graph)}", e, 3e-2)
    # banks matter (otherwise this test could not see a broken hand-over)
    plain = pipe.denoise(lat.cuda(), ctx.cuda(), num_inference_steps=2, guidance_scale=7.5, context_frames=8,
                         context_overlap=2, use_cuda_graph=use_graph)
    assert rel_l2(plain, ref) > 2 * e
    with pytest.raises(ValueError):
        pipe.denoise(lat.cuda(), ctx.cuda(), num_inference_steps=1, appearance_encoder=enc)


def test_baseline_config0_full_width_plumbing_case():
    """BASELINE.json configs[0] — 1x64x64 latent, 1 frame, 2 DDIM steps, random-init SD-1.5-width UNet (the reference's
    own CPU-runnable case) — end to end on CUDA: denoise (CFG pair) + full-width VAE decode, against the oracle pieces
    evaluated in fp32 (on the GPU with TF32 off, so the check finishes in seconds)."""
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    from emote_hack_b200.unet3d import UNet3DConditionModel
    from emote_hack_b200.vae import AutoencoderKL
    from oracle.unet3d_port import UNet3DOracle
    from oracle.vae_decoder import VAEDecoderOracle
    from oracle.ddim import DDIMOracle, cfg_combine
    from util_models import FULL_CFG
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    with torch.device("cuda"):
        unet = UNet3DConditionModel(**FULL_CFG).eval()
        vae = AutoencoderKL().eval()
    rerandomise_zero_inits(unet)
    o = UNet3DOracle(unet.state_dict(), dict(unet.config), device="cuda")
    vo = VAEDecoderOracle(vae.state_dict())
    g = torch.Generator().manual_seed(1234)
    lat = torch.randn(1, 4, 1, 64, 64, generator=g).cuda()
    ctx = torch.randn(2, 77, 768, generator=g).cuda()
    sch = DDIMOracle()
    ref = lat.clone()
    for t in sch.set_timesteps(2).tolist():
        pred = o(ref.repeat(2, 1, 1, 1, 1), torch.tensor(t).cuda(), ctx)
        ref = sch.step(cfg_combine(pred, torch.ones(1, 1, 1, 1, 1, device="cuda"), 7.5), t, ref)
    pipe = EMOAnimationPipeline(vae, unet, DDIMScheduler())
    out = pipe.denoise(lat.clone(), ctx, num_inference_steps=2, guidance_scale=7.5)
    e = rel_l2(out, ref)
    want = vo.decode_latents(ref)
    video = torch.from_numpy(pipe.decode_latents(ref))
    ev = rel_l2(video, want)
    print(f"config[0] full width: 2-step denoise rel_l2={e:.2e}; VAE decode 512x512 rel_l2={ev:.2e}")
    assert video.shape == (1, 3, 1, 512, 512) and float(video.min()) >= 0 and float(video.max()) <= 1
    check_parity("config0.denoise_2steps_full_width", e, 3e-2)
    check_parity("config0.vae_decode_512", ev, 2e-2)
    del unet, vae, o, vo
    torch.cuda.empty_cache()


def test_denoise_per_frame_audio_context(tiny_pipe):
    o, pipe = tiny_pipe
    g = torch.Generator().manual_seed(13)
    lat = torch.randn(1, 4, 6, 8, 8, generator=g)
    ctx = torch.randn(12, 5, 64, generator=g)  # [uncond frames | cond frames], 5 audio tokens per frame
    ref = _oracle_denoise(o, lat, ctx, 2, 7.5, [list(range(6))], per_frame=True)
    out = pipe.denoise(lat.cuda(), ctx.cuda(), num_inference_steps=2, guidance_scale=7.5, context_frames=16)
    check_parity("denoise.per_frame_audio_ctx", rel_l2(out, ref), 3e-2)


def test_pipeline_call_returns_video(tiny_pipe, tiny_vae):
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    o, pipe = tiny_pipe
    voracle, vae = tiny_vae
    p = EMOAnimationPipeline(vae, pipe.unet, DDIMScheduler())
    g = torch.Generator(device="cuda").manual_seed(3)
    ctx = torch.randn(2, 7, 64, device="cuda", generator=g)
    lat0 = torch.randn(1, 4, 4, 8, 8, device="cuda", generator=g)
    out = p(ctx, video_length=4, height=64, width=64, num_inference_steps=2, latents=lat0.clone())
    v = out.videos
    assert v.shape == (1, 3, 4, 64, 64) and v.dtype == torch.float32 and 0 <= float(v.min()) and float(v.max()) <= 1
    ref_lat = _oracle_denoise(o, lat0.cpu(), ctx.cpu(), 2, 7.5, [list(range(4))])
    assert (v - voracle.decode_latents(ref_lat)).abs().max() < 5e-2


def test_pipeline_call_reference_keyword_surface(tiny_pipe, tiny_vae):
    """the reference's `__call__` keywords (EMOAnimationPipeline.py:543-578) are accepted with their meaning; unknown ones
    and the out-of-scope ones fail loudly instead of being swallowed"""
    from emote_hack_b200.pipeline import DDIMScheduler, EMOAnimationPipeline
    o, pipe = tiny_pipe
    voracle, vae = tiny_vae
    p = EMOAnimationPipeline(vae, pipe.unet, DDIMScheduler())
    g = torch.Generator(device="cuda").manual_seed(5)
    emb = torch.randn(2, 7, 64, device="cuda", generator=g)
    init = torch.randn(4, 4, 8, 8, device="cuda", generator=g)          # (b f) c h w, the `invert` output layout
    seen = []
    out = p(prompt=None, video_length=4, height=64, width=64, num_inference_steps=4, guidance_scale=7.5,
            negative_prompt=None, num_videos_per_prompt=1, eta=0.0, generator=None, latents=None, output_type="tensor",
            return_dict=False, callback=lambda i, t, lat: seen.append((i, t)), callback_steps=1, controlnet_condition=None,
            controlnet_conditioning_scale=1.0, context_frames=16, context_stride=1, context_overlap=4, context_batch_size=1,
            context_schedule="uniform", init_latents=init, num_actual_inference_steps=2, appearance_encoder=None,
            reference_control_writer=None, reference_control_reader=None, source_image=None, decoder_consistency=None,
            audio=None, head_rotation_speeds=None, prompt_embeddings=emb)
    assert out.shape == (1, 3, 4, 64, 64)
    assert [i for i, _ in seen] == [2, 3]                                 # num_actual_inference_steps skips the first two
    # the same two steps on the oracle
    from oracle.ddim import DDIMOracle, cfg_combine
    sch = DDIMOracle()
    lat = init.cpu().reshape(1, 4, 4, 8, 8).permute(0, 2, 1, 3, 4).contiguous()
    for t in sch.set_timesteps(4).tolist()[2:]:
        pred = o(lat.repeat(2, 1, 1, 1, 1), t, emb.cpu())
        lat = sch.step(cfg_combine(pred, torch.ones(1, 1, 4, 1, 1), 7.5), t, lat)
    assert (out - voracle.decode_latents(lat)).abs().max() < 5e-2
    with pytest.raises(TypeError):
        p(emb, video_length=4, height=64, width=64, num_inference_steps=1, not_a_reference_kwarg=1)
    with pytest.raises(NotImplementedError):
        p("a portrait", video_length=4, height=64, width=64, num_inference_steps=1)      # no text_encoder attached
    with pytest.raises(NotImplementedError):
        p(emb, video_length=4, height=64, width=64, num_inference_steps=1, controlnet_condition=[0])
    # a text_encoder front-end makes string prompts work (stand-in encoder: the CLIP model itself is out of scope)
    p2 = EMOAnimationPipeline(vae, pipe.unet, DDIMScheduler(),
                              text_encoder=lambda prompts: torch.full((len(prompts), 7, 64), 0.1 * len(prompts[0]), device="cuda"))
    v = p2("hello", video_length=4, height=64, width=64, num_inference_steps=1, negative_prompt="no").videos
    assert v.shape == (1, 3, 4, 64, 64)


def test_stochastic_ddim_eta_matches_the_published_update(tiny_pipe):
    """eta > 0 (EMOAnimationPipeline.py:553,817 `extra_step_kwargs`): x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev-s^2) eps + s z
    with s = eta sqrt((1-a_prev)/(1-a_t) (1-a_t/a_prev)) (Song et al. 2021 eq. 12/16) — scheduler.step and the fused
    CFG kernel against a direct fp64 evaluation; eta = 0 reproduces the deterministic step."""
    from emote_hack_b200 import ops
    o, pipe = tiny_pipe
    s = pipe.scheduler
    s.set_timesteps(50)
    g = torch.Generator().manual_seed(3)
    x, eps, z = (torch.randn(1, 4, 3, 8, 8, generator=g) for _ in range(3))
    for t, eta in ((981, 1.0), (481, 0.3), (21, 0.7)):
        a_t, a_p = s.alphas_for(t)
        sig = eta * ((1 - a_p) / (1 - a_t) * (1 - a_t / a_p)) ** 0.5
        xd, ed, zd = x.double(), eps.double(), z.double()
        x0 = (xd - (1 - a_t) ** 0.5 * ed) / a_t ** 0.5
        want = a_p ** 0.5 * x0 + (1 - a_p - sig ** 2) ** 0.5 * ed + sig * zd
        got = s.step(eps.cuda(), t, x.cuda(), eta=eta, variance_noise=z.cuda()).prev_sample
        assert rel_l2(got, want.float()) < 1e-5
        pair = torch.cat([eps, eps + 0.0]).cuda().contiguous()
        fused = ops.cfg_ddim_step(x.cuda().clone(), pair, None, 7.5, a_t, a_p, noise=z.cuda().contiguous(),
                                  sigma=ops.ddim_sigma(a_t, a_p, eta))
        assert rel_l2(fused, want.float()) < 1e-5
    det = s.step(eps.cuda(), 481, x.cuda()).prev_sample
    assert rel_l2(s.step(eps.cuda(), 481, x.cuda(), eta=0.0, variance_noise=z.cuda()).prev_sample, det) == 0.0
    # inside the loop: a generator makes the stochastic sampler reproducible, and it differs from eta = 0
    lat = torch.randn(1, 4, 4, 8, 8, generator=g).cuda()
    ctx = torch.randn(2, 7, 64, generator=g).cuda()
    kw = dict(num_inference_steps=3, guidance_scale=7.5, context_frames=16)
    a = pipe.denoise(lat.clone(), ctx, eta=0.5, generator=torch.Generator(device="cuda").manual_seed(1), **kw)
    b = pipe.denoise(lat.clone(), ctx, eta=0.5, generator=torch.Generator(device="cuda").manual_seed(1), **kw)
    assert rel_l2(a, b) < 1e-6 and rel_l2(a, pipe.denoise(lat.clone(), ctx, **kw)) > 1e-2


def test_long_clip_decode_is_chunked(tiny_vae):
    """decode of a long clip walks the frames in bounded chunks (default 16): same frames as one big batch"""
    oracle, vae = tiny_vae
    lat = torch.randn(1, 4, 37, 8, 8, generator=torch.Generator().manual_seed(9)).cuda() * 0.3
    whole, _ = vae.decode_video(lat, frame_chunk=37)
    chunked, u8 = vae.decode_video(lat, want_u8=True)          # default chunking: 16 + 16 + 5
    assert chunked.shape == (1, 3, 37, 64, 64) and u8.shape == chunked.shape
    assert rel_l2(chunked, whole) < 1e-5


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
