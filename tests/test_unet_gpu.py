"""Parity of the CUDA UNet3D path against the oracle port (fp32 restatement of the reference, pinned in
tests/test_oracle.py).  Tolerance protocol: tensor-core operands are bf16 (2^-9 relative
rounding, ~1.6e-3 rel-L2 per GEMM), residual stream / statistics / accumulation fp32.  Per block: rel-L2 <= 5e-3.
Whole network: rel-L2 <= 2e-2 and no worse than 1.25x the error of running the same restatement in plain eager bf16.
"""
import pytest
import torch

from util_models import (FULL_CFG, TINY_CFG, check_parity, make_banks, make_inputs, rel_l2, rerandomise_zero_inits)

pytestmark = pytest.mark.gpu


def _build(cfg, seed=0):
    from emote_hack_b200.unet3d import UNet3DConditionModel
    torch.manual_seed(seed)
    m = UNet3DConditionModel(**cfg).eval()
    return rerandomise_zero_inits(m)


def _oracle(model, **kw):
    from oracle.unet3d_port import UNet3DOracle
    return UNet3DOracle(model.state_dict(), dict(model.config), **kw)


@pytest.fixture(scope="module")
def tiny():
    m = _build(TINY_CFG)
    return m, _oracle(m), m.cuda()


def test_resnet_block(tiny):
    m, o, _ = tiny
    x, _ = make_inputs(2, 4, 8)
    g = torch.Generator().manual_seed(3)
    h = torch.randn(2, 64, 4, 8, 8, generator=g)
    emb = torch.randn(2, 256, generator=g)
    for name, inp in (("down_blocks.0.resnets.0", h), ("down_blocks.1.resnets.0", h)):
        ref = o._resnet(name, inp, emb)
        out = dict(m.named_modules())[name](inp.cuda(), emb.cuda())
        e = rel_l2(out, ref)
        check_parity(f"block.resnet.{name}", e, 5e-3)
    # concatenated (hidden, skip) input of the up blocks
    name = "up_blocks.3.resnets.2"
    a, b = torch.randn(2, 64, 4, 8, 8, generator=g), torch.randn(2, 64, 4, 8, 8, generator=g)
    ref = o._resnet(name, torch.cat([a, b], 1), emb)
    out = dict(m.named_modules())[name]((a.cuda(), b.cuda()), emb.cuda())
    check_parity("block.resnet.concat_input", rel_l2(out, ref), 5e-3)


def test_transformer3d_and_motion(tiny):
    m, o, _ = tiny
    g = torch.Generator().manual_seed(4)
    h = torch.randn(2, 64, 4, 8, 8, generator=g)
    ctx = torch.randn(2, 7, 64, generator=g)
    mods = dict(m.named_modules())
    ref = o._transformer3d("down_blocks.0.attentions.0", h, ctx, 4, None, True)
    out = mods["down_blocks.0.attentions.0"](h.cuda(), encoder_hidden_states=ctx.cuda()).sample
    e = rel_l2(out, ref)
    check_parity("block.transformer3d", e, 5e-3)
    # per-frame context (audio tokens): batch == b*f, not repeated (attention.py:118-119)
    ctx_f = torch.randn(8, 5, 64, generator=g)
    ref = o._transformer3d("down_blocks.0.attentions.0", h, ctx_f, 4, None, True)
    out = mods["down_blocks.0.attentions.0"](h.cuda(), encoder_hidden_states=ctx_f.cuda()).sample
    check_parity("block.transformer3d_per_frame_ctx", rel_l2(out, ref), 5e-3)
    ref = o._motion("down_blocks.0.motion_modules.0", h)
    out = mods["down_blocks.0.motion_modules.0"](h.cuda(), None, None)
    e = rel_l2(out, ref)
    check_parity("block.motion_module", e, 5e-3)
    # the temporal branch must actually contribute (zero-init proj_out was re-randomised)
    assert rel_l2(ref, h) > 1e-3


def test_samplers(tiny):
    m, o, _ = tiny
    g = torch.Generator().manual_seed(5)
    h = torch.randn(2, 64, 4, 8, 8, generator=g)
    mods = dict(m.named_modules())
    ref = o._conv5("down_blocks.0.downsamplers.0.conv", h, stride=2)
    out = mods["down_blocks.0.downsamplers.0"](h.cuda())
    assert out.shape == ref.shape
    check_parity("block.downsample", rel_l2(out, ref), 5e-3)
    h2 = torch.randn(2, 128, 4, 4, 4, generator=g)
    ref = o._conv5("up_blocks.1.upsamplers.0.conv", torch.nn.functional.interpolate(h2, scale_factor=(1.0, 2.0, 2.0)))
    out = mods["up_blocks.1.upsamplers.0"](h2.cuda())
    assert out.shape == ref.shape
    check_parity("block.upsample", rel_l2(out, ref), 5e-3)


@pytest.mark.parametrize("hw,f", [(8, 4), (16, 8), (32, 2)])
def test_unet_tiny_end_to_end(tiny, hw, f):
    m, o, _ = tiny
    x, ctx = make_inputs(2, f, hw)
    t = torch.tensor(481)
    ref = o(x, t, ctx)
    out = m(x.cuda(), t.cuda(), ctx.cuda()).sample
    assert out.shape == ref.shape and out.dtype == torch.float32
    e = rel_l2(out, ref)
    bf = _oracle(m, dtype=torch.bfloat16, device="cuda")(x.cuda(), t.cuda(), ctx.cuda())
    e_bf = rel_l2(bf, ref)
    print(f"unet tiny hw={hw} f={f}: rel_l2={e:.2e} (eager-bf16 restatement: {e_bf:.2e})")
    check_parity(f"unet.tiny_hw{hw}_f{f}", e, 2e-2)
    assert e < 1.25 * e_bf + 1e-3   # never worse than plain eager bf16 PyTorch on the same network
    # tuple return + python scalar timestep
    out2 = m(x.cuda(), 481, ctx.cuda(), return_dict=False)[0]
    assert rel_l2(out2, out) < 1e-5


def test_unet_tiny_reference_attention(tiny):
    """reader semantics of mutual_self_attention.py:237-258: cond half attends to [self | bank], uncond half does not"""
    from emote_hack_b200.unet3d import ReferenceAttentionControl
    m, o, _ = tiny
    x, ctx = make_inputs(2, 4, 16)
    t = torch.tensor(301)
    banks = make_banks(m, 16)
    assert len(banks) == 10
    ref = o(x, t, ctx, banks=banks)
    ref_plain = o(x, t, ctx)
    reader = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup")
    try:
        reader.set_banks({k: [v.cuda() for v in vs] for k, vs in banks.items()})
        out = m(x.cuda(), t.cuda(), ctx.cuda()).sample
        e = rel_l2(out, ref)
        print(f"bank effect={rel_l2(ref, ref_plain):.2e}")
        check_parity("unet.tiny_reference_banks", e, 2e-2)
        assert rel_l2(ref, ref_plain) > 1e-2  # banks matter
        # banks are consumed (cleared) by the forward: the next call is the plain network again
        out_plain = m(x.cuda(), t.cuda(), ctx.cuda()).sample
        check_parity("unet.tiny_after_banks_consumed", rel_l2(out_plain, ref_plain), 2e-2)
        # the unconditional half never sees the bank
        assert rel_l2(out[0], out_plain[0]) < 1e-6
    finally:
        for blk in reader._blocks(m):
            blk._ref_mode = None


def test_single_branch_calls_equal_the_cfg_pair(tiny):
    """the unit of multi-GPU work is (window x CFG branch).  A batch-1 call of one branch — the
    unconditional one without banks, the conditional one with bank row 1 and no CFG masking — reproduces its half of the
    batch-2 call (pipeline._run_unet modes "uncond" / "cond" vs "pair")."""
    from emote_hack_b200.pipeline import _run_unet
    _, o, m = tiny
    x, ctx = make_inputs(2, 4, 16)
    banks = make_banks(m, 16)
    ref = o(x, torch.tensor(301), ctx, banks=banks)                       # oracle, CFG pair with banks
    x, ctx = x.cuda(), ctx.cuda()
    banks = {k: [v.cuda() for v in vs] for k, vs in banks.items()}
    t = torch.tensor([301.0], device="cuda")
    pair = _run_unet(m, "pair", x, t, ctx, banks)
    un = _run_unet(m, "uncond", x[0:1].contiguous(), t, ctx[0:1], None)
    co = _run_unet(m, "cond", x[1:2].contiguous(), t, ctx[1:2], {k: [v[0][1:2]] for k, v in banks.items()})
    # each branch against the oracle at the operand-rounding level ...
    check_parity("unet.single_branch_uncond", rel_l2(un, ref[0:1]), 2e-2)
    check_parity("unet.single_branch_cond", rel_l2(co, ref[1:2]), 2e-2)
    # ... and against its half of the batch-2 call: not bit-identical (different row counts select different GEMM tile
    # schedules, and this random-weight network amplifies fp32-round-off differences up to the operand-rounding noise
    # floor), but far below the effect of a wrong bank row / a leaked bank (> 1e-2 below)
    check_parity("unet.single_branch_uncond_vs_pair", rel_l2(un, pair[0:1]), 2e-2)
    check_parity("unet.single_branch_cond_vs_pair", rel_l2(co, pair[1:2]), 2e-2)
    plain = _run_unet(m, "pair", x, t, ctx, None)
    assert rel_l2(plain[1:2], pair[1:2]) > 1e-2          # the bank mattered for the conditional half
    wrong = _run_unet(m, "cond", x[1:2].contiguous(), t, ctx[1:2], {k: [v[0][0:1]] for k, v in banks.items()})
    assert rel_l2(wrong, pair[1:2]) > 1e-2               # ... and so does WHICH bank row the lone branch reads
    assert all(blk._ref_mode is None and not blk.bank for blk in m.modules() if hasattr(blk, "_ref_mode"))  # disarmed


def test_in_place_weight_updates_refresh_every_packed_copy(tiny):
    """param.mul_() / copy_() (a LoRA merge, an optimiser step) must refresh the packed 16-bit copies of EVERY module kind —
    attention projections, feed-forward, convolutions, biases — not only after load_state_dict."""
    m, o, _ = tiny
    x, ctx = make_inputs(2, 2, 8)
    base = m(x.cuda(), 5, ctx.cuda()).sample
    names = ["down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight",
             "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.2.weight",
             "down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.bias",
             "down_blocks.0.motion_modules.0.temporal_transformer.transformer_blocks.0.attention_blocks.0.to_out.0.bias",
             "down_blocks.0.resnets.0.conv1.bias", "down_blocks.0.resnets.0.time_emb_proj.weight",
             "down_blocks.0.attentions.0.proj_in.bias", "time_embedding.linear_2.bias"]
    params = dict(m.named_parameters())
    for n in names:
        saved = params[n].detach().clone()
        try:
            with torch.no_grad():
                params[n].mul_(1.5).add_(0.05)
            changed = m(x.cuda(), 5, ctx.cuda()).sample
            assert rel_l2(changed, base) > 1e-5, f"in-place update of {n} was not picked up"
        finally:
            with torch.no_grad():
                params[n].copy_(saved)
    assert rel_l2(m(x.cuda(), 5, ctx.cuda()).sample, base) < 1e-5


def test_appearance_encoder_writer_banks_and_reader_update(tiny):
    """ReferenceNet writer (AppearanceEncoderModel, appearance_encoder.py) on CUDA: banks against the oracle run as a
    writer, then ReferenceAttentionControl.update() hands them to the UNet3D reader (EMOAnimationPipeline.py:711-788)."""
    import json
    from pathlib import Path
    from emote_hack_b200.appearance_encoder import AppearanceEncoderModel
    from emote_hack_b200.unet3d import ReferenceAttentionControl
    from oracle.unet3d_port import UNet3DOracle
    from util_models import (APPEARANCE_TRIMMED, appearance_cfg, reader_block_names, seeded_unet_state_dict, writer_cfg,
                             writer_inputs)
    shapes = json.loads((Path(__file__).parent / "golden" / "writer_tiny_keys.json").read_text())
    sd = {k: v for k, v in seeded_unet_state_dict(shapes, 3).items() if not k.startswith(APPEARANCE_TRIMMED)}
    enc = AppearanceEncoderModel(**appearance_cfg()).eval()
    missing, unexpected = enc.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    enc = enc.cuda()
    x, ctx = writer_inputs()
    want = {}
    UNet3DOracle(sd, writer_cfg())(x[:, :, None], 441, ctx, collect_banks=want)
    writer = ReferenceAttentionControl(enc, do_classifier_free_guidance=True, mode="write", fusion_blocks="midup")
    out = enc(x.cuda(), 441, ctx.cuda()).sample
    assert out.shape == (2, 64, 16, 16) and torch.isfinite(out).all()
    names = reader_block_names(enc)
    mods = dict(enc.named_modules())
    assert len(names) == 10
    for n in names:
        assert len(mods[n].bank) == 1
        check_parity(f"writer.bank.{n}", rel_l2(mods[n].bank[0], want[n][0]), 2e-2)   # a LayerNorm1 output deep inside the network
    # hand the banks to the video UNet's reader blocks and compare with the oracle fed the oracle's banks
    m, o, _ = tiny
    reader = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup")
    try:
        reader.update(writer)
        writer.clear()
        assert all(len(mods[n].bank) == 0 for n in names)
        xs, cs = make_inputs(2, 4, 16)
        out = m(xs.cuda(), 301, cs.cuda()).sample
        ref = o(xs, 301, cs, banks={n: want[n] for n in names})
        e = rel_l2(out, ref)
        check_parity("writer.reader_end_to_end", e, 2e-2)
    finally:
        for blk in reader._blocks(m):
            blk._ref_mode = None
        for blk in writer._blocks(enc):
            blk._ref_mode = None


def test_unet_controlnet_residuals(tiny):
    m, o, _ = tiny
    from util_models import controlnet_residuals
    x, ctx = make_inputs(2, 2, 8)
    down, mid = controlnet_residuals()
    ref = o(x, torch.tensor(10), ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
    out = m(x.cuda(), 10, ctx.cuda(), down_block_additional_residuals=[d.cuda() for d in down],
            mid_block_additional_residual=mid.cuda()).sample
    check_parity("unet.tiny_controlnet_residuals", rel_l2(out, ref), 2e-2)


def test_state_dict_reload_invalidates_packed_weights(tiny):
    m, o, _ = tiny
    x, ctx = make_inputs(2, 2, 8)
    out1 = m(x.cuda(), 5, ctx.cuda()).sample
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m.load_state_dict({k: (v * 1.01 if v.dtype.is_floating_point and "pe" not in k else v) for k, v in sd.items()})
    out2 = m(x.cuda(), 5, ctx.cuda()).sample
    assert rel_l2(out2, out1) > 1e-4
    m.load_state_dict(sd)
    # GroupNorm statistics are reduced with atomics: bit-level run-to-run differences are expected
    assert rel_l2(m(x.cuda(), 5, ctx.cuda()).sample, out1) < 1e-5


def test_reference_import_paths_run_the_cuda_modules():
    """INTEGRATION.md §1 end to end: with the package's `magicanimate` tree aliased into sys.modules, the reference's own
    import lines (EMOAnimationPipeline.py:54-56) build the B200 modules and a forward runs on the kernels."""
    import importlib
    import sys
    import emote_hack_b200.magicanimate as b200
    saved = {k: v for k, v in sys.modules.items() if k == "magicanimate" or k.startswith("magicanimate.")}
    try:
        for k in saved:
            del sys.modules[k]
        sys.modules["magicanimate"] = b200
        sys.modules["magicanimate.models"] = b200.models
        for name in ("unet_controlnet", "unet_3d_blocks", "resnet", "attention", "motion_module", "mutual_self_attention"):
            sys.modules[f"magicanimate.models.{name}"] = getattr(b200.models, name)
        UNet = importlib.import_module("magicanimate.models.unet_controlnet").UNet3DConditionModel
        from magicanimate.models.mutual_self_attention import ReferenceAttentionControl
        torch.manual_seed(0)
        m = rerandomise_zero_inits(UNet(**TINY_CFG).eval()).cuda()
        reader = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup", batch_size=1)
        x, ctx = make_inputs(2, 2, 8)
        from emote_hack_b200 import _lib
        n0 = _lib.launch_count()
        out = m(x.cuda(), torch.tensor(301).cuda(), encoder_hidden_states=ctx.cuda(), return_dict=False)[0]
        assert out.shape == (2, 4, 2, 8, 8) and torch.isfinite(out).all() and _lib.launch_count() > n0
        reader.release()
    finally:
        for k in [k for k in sys.modules if k == "magicanimate" or k.startswith("magicanimate.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_cpu_tensors_fail_loudly():
    from emote_hack_b200._lib import EmoteKernelError
    from emote_hack_b200.unet3d import ResnetBlock3D
    blk = ResnetBlock3D(in_channels=64, out_channels=64, temb_channels=256)
    with pytest.raises(EmoteKernelError):
        blk(torch.randn(1, 64, 2, 8, 8), torch.randn(1, 256))


@pytest.mark.slow
def test_unet_full_width_one_frame_pair():
    """SD-1.5 widths (1276.7 M params), 64x64 latent, 2 frames: head dims 40/80/160 and all concat widths."""
    m = _build(FULL_CFG, seed=0)
    x, ctx = make_inputs(1, 2, 64, ctx_tokens=77, ctx_dim=768)
    o = _oracle(m, device="cuda")  # fp32 on GPU (TF32 off) so the check finishes in seconds
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = o(x.cuda(), torch.tensor(981).cuda(), ctx.cuda())
    m = m.cuda()
    out = m(x.cuda(), 981, ctx.cuda()).sample
    e = rel_l2(out, ref)
    check_parity("unet.full_width_1x2f_64", e, 2e-2)


def test_unet_full_size_properties():
    """BASELINE config #2 at FULL size ([2,4,16,64,64], SD-1.5 widths, 77 text tokens) — too large for the fp32 oracle
    in a test, so checked through size-independent properties of the reference network:
    (1) samples of a batch do not interact (5-D GroupNorm is per sample, attention is per image): identical halves in,
        identical halves out, and equal to the single-sample call;
    (2) determinism: two calls agree to fp32 round-off (GroupNorm statistics are reduced in a fixed order / fp64);
    (3) the time embedding matters, the context matters (no dead inputs), and the output is finite."""
    m = _build(FULL_CFG).cuda()
    g = torch.Generator().manual_seed(99)
    x1 = torch.randn(1, 4, 16, 64, 64, generator=g).cuda()
    c1 = torch.randn(1, 77, 768, generator=g).cuda()
    x2, c2 = x1.repeat(2, 1, 1, 1, 1), c1.repeat(2, 1, 1)
    out2 = m(x2, 981, c2).sample
    assert out2.shape == (2, 4, 16, 64, 64) and torch.isfinite(out2).all()
    assert rel_l2(out2[0], out2[1]) < 1e-5
    again = m(x2, 981, c2).sample
    assert rel_l2(again, out2) < 1e-5
    single = m(x1, 981, c1).sample
    assert rel_l2(single[0], out2[0]) < 1e-5
    assert rel_l2(m(x2, 21, c2).sample, out2) > 1e-2
    other_ctx = torch.cat([c1, torch.randn(1, 77, 768, generator=g).cuda()])
    out_ctx = m(x2, 981, other_ctx).sample
    assert rel_l2(out_ctx[0], out2[0]) < 1e-5 and rel_l2(out_ctx[1], out2[1]) > 1e-3
    del m
    torch.cuda.empty_cache()


def test_unet_latent_sizes_that_need_2d_conv_tiles(tiny):
    """BASELINE config #4 geometry class (96x96 latents and their 48 / 24 / 12 levels): image rows that do not pack into
    128-pixel runs are convolved through 2-D patch tiles (4-D TMA boxes) — no im2col copy is launched."""
    from emote_hack_b200 import _lib, ops
    m, o, _ = tiny
    x, ctx = make_inputs(2, 2, 24)
    ref = o(x, 201, ctx)
    with ops.KernelProfiler() as prof:
        out = m(x.cuda(), 201, ctx.cuda()).sample
    assert "emote_im2col3x3_bf16" not in prof.summary()
    check_parity("unet.tiny_24x24_latent", rel_l2(out, ref), 2e-2)


def test_unet_32_frame_window_audio_tokens_and_banks():
    """BASELINE config #3 features together: one 32-frame window (needs temporal_position_encoding_max_len >= 32, the
    shipped 24 crashes in the reference too: motion_module.py:247), per-frame audio tokens as context, reference banks."""
    import copy
    from emote_hack_b200.unet3d import ReferenceAttentionControl, UNet3DConditionModel
    from oracle.unet3d_port import UNet3DOracle
    cfg = copy.deepcopy(TINY_CFG)
    cfg["motion_module_kwargs"]["temporal_position_encoding_max_len"] = 32
    torch.manual_seed(2)
    m = rerandomise_zero_inits(UNet3DConditionModel(**cfg).eval())
    o = UNet3DOracle(m.state_dict(), dict(m.config))
    m = m.cuda()
    x, ctx = make_inputs(2, 32, 8, ctx_tokens=5, per_frame_ctx=True)
    banks = make_banks(m, 8)
    ref = o(x, 641, ctx, banks=banks)
    reader = ReferenceAttentionControl(m, do_classifier_free_guidance=True, mode="read", fusion_blocks="midup")
    reader.set_banks({k: [v.cuda() for v in vs] for k, vs in banks.items()})
    out = m(x.cuda(), 641, ctx.cuda()).sample
    e = rel_l2(out, ref)
    check_parity("unet.tiny_f32_audio_banks", e, 2e-2)
    with pytest.raises(ValueError):  # 24-entry table, 32 frames
        m24 = rerandomise_zero_inits(UNet3DConditionModel(**TINY_CFG).eval()).cuda()
        m24(x.cuda(), 641, ctx.cuda())
