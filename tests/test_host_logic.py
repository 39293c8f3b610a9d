"""Host-side logic that needs no GPU: scheduler tables, window partitioning, and the world_size-2 reduction that
replaces the reference's gather + broadcast (gloo on CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_counter_and_partition_cover_all_frames():
    from emote_hack_b200.pipeline import uniform
    for nf, cs, ov in [(240, 16, 4), (32, 16, 4), (24, 8, 2)]:
        wins = list(uniform(0, 50, nf, cs, 1, ov))
        cnt = torch.zeros(nf)
        for w in wins:
            cnt[w] += 1
        assert cnt.min() >= 1  # every frame is denoised at least once
        for world in (2, 4, 8):
            parts = [wins[r::world] for r in range(world)]
            assert sorted(map(tuple, sum(parts, []))) == sorted(map(tuple, wins))
    assert len(list(uniform(0, 50, 240, 16, 1, 4))) == 20  # 240 f -> 20 windows
    assert len(list(uniform(0, 50, 32, 16, 1, 4))) == 3


def test_unit_partition_is_exact_balanced_and_pairs_branches():
    """plan_units: the (window x CFG-branch) units are dealt to the ranks in equal contiguous blocks — every
    unit exactly once, per-rank loads differ by at most one unit, both branches of a window on one rank fuse into a pair"""
    from emote_hack_b200.pipeline import plan_units
    for n_win in (1, 3, 4, 15, 20):
        for world in (1, 2, 3, 4, 8):
            seen, loads = [], []
            for r in range(world):
                calls = plan_units(n_win, r, world)
                units = [(w, b) for w, mode in calls for b in ((0, 1) if mode == "pair" else ((0,) if mode == "uncond" else (1,)))]
                loads.append(len(units))
                seen += units
                assert all(mode == "pair" or (w, mode) not in [(ww, "pair") for ww, _ in calls] for w, mode in calls)
            assert sorted(seen) == [(w, b) for w in range(n_win) for b in (0, 1)]
            assert max(loads) - min(loads) <= 1
    assert [len(plan_units(20, r, 8)) for r in range(8)] == [3] * 8            # 40 units -> 5 per rank = 2 pairs + 1 branch
    assert plan_units(1, 0, 2) == [(0, "uncond")] and plan_units(1, 1, 2) == [(0, "cond")]   # one clip: <= 2x
    assert plan_units(1, 2, 4) == []                                          # idle ranks still join the all-reduce
    # the reference's own split (EMOAnimationPipeline.py:757): whole windows round-robin
    assert plan_units(20, 3, 8, shard="windows") == [(3, "pair"), (11, "pair"), (19, "pair")]
    assert max(len(plan_units(20, r, 8, shard="windows")) for r in range(8)) == 3   # 3,3,3,3,2,2,2,2 -> <= 6.67x


def test_ops_refuse_cpu_tensors():
    from emote_hack_b200 import ops
    from emote_hack_b200._lib import EmoteKernelError
    with pytest.raises(EmoteKernelError):
        ops.layer_norm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nf, cs, ov, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from emote_hack_b200.pipeline import uniform
    wins = list(uniform(0, 50, nf, cs, 1, ov))
    g = torch.Generator().manual_seed(0)
    table = torch.randn(len(wins), 2, 4, cs, 2, 2, generator=g)  # stand-in per-window predictions
    from emote_hack_b200.pipeline import plan_units
    acc = torch.zeros(2, 4, nf, 2, 2)
    for wi, mode in plan_units(len(wins), rank, world):   # (window x CFG-branch) units of this rank
        b0, nb = (0, 2) if mode == "pair" else ((0, 1) if mode == "uncond" else (1, 1))
        acc[b0:b0 + nb, :, wins[wi]] += table[wi][b0:b0 + nb]
    dist.all_reduce(acc)  # the single per-step collective (replaces gather->rank0 sum->broadcast, :796-821)
    if rank == 0:
        full = torch.zeros(2, 4, nf, 2, 2)
        for wi, w in enumerate(wins):
            full[:, :, w] += table[wi]
        ret.put(bool(torch.allclose(acc, full, atol=1e-6)))
    dist.destroy_process_group()


def test_world_size_2_window_sharding_all_reduce_equals_single_process():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    for nf in (40, 28, 16):   # 4 windows (whole pairs per rank), 3 windows (one window split by branch), 1 window
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, nf, 16, 4, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret.get(timeout=10) is True


def test_reference_module_paths_resolve_to_b200_classes():
    """the reference's import paths, aliased through sys.modules, give the B200 classes."""
    import importlib
    import sys
    import emote_hack_b200.magicanimate as b200
    from emote_hack_b200 import unet3d
    saved = {k: v for k, v in sys.modules.items() if k == "magicanimate" or k.startswith("magicanimate.")}
    try:
        for k in saved:
            del sys.modules[k]
        sys.modules["magicanimate"] = b200
        sys.modules["magicanimate.models"] = b200.models
        for name in ("unet_controlnet", "unet_3d_blocks", "resnet", "attention", "motion_module", "mutual_self_attention"):
            sys.modules[f"magicanimate.models.{name}"] = getattr(b200.models, name)
        mod = importlib.import_module("magicanimate.models.unet_controlnet")
        assert mod.UNet3DConditionModel is unet3d.UNet3DConditionModel
        from magicanimate.models.mutual_self_attention import ReferenceAttentionControl
        assert ReferenceAttentionControl is unet3d.ReferenceAttentionControl
        from magicanimate.models.motion_module import get_motion_module
        assert get_motion_module is unet3d.get_motion_module
    finally:
        for k in [k for k in sys.modules if k == "magicanimate" or k.startswith("magicanimate.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_window_scheduler_matches_restatement_and_live_reference_on_random_arguments():
    """`pipeline.uniform` (product) vs `oracle.ddim.uniform_windows` (restatement) vs the reference's own
    `magicanimate/pipelines/context.py:uniform` (when /root/reference is present) over a sweep of arguments, including
    strides > 1, zero overlap and videos shorter than one window."""
    import itertools
    from emote_hack_b200.pipeline import uniform
    from oracle import ref_shim
    from oracle.ddim import uniform_windows
    live = ref_shim.load_reference_context_uniform() if ref_shim.reference_available() else None
    n = 0
    for nf, cs, stride, ov, step in itertools.product((1, 7, 16, 24, 33, 64, 240), (8, 16, 24), (1, 2, 3), (0, 2, 4), (0, 3, 17)):
        if ov >= cs:
            continue
        ours = [list(map(int, w)) for w in uniform(step, 50, nf, cs, stride, ov)]
        assert ours == uniform_windows(step, 50, nf, cs, stride, ov), (nf, cs, stride, ov, step)
        if live is not None:
            assert ours == [list(map(int, w)) for w in live(step, 50, nf, cs, stride, ov)], (nf, cs, stride, ov, step)
        assert all(0 <= f < nf for w in ours for f in w)
        n += 1
    assert n > 400


def test_weights_fingerprint_tracks_in_place_updates_and_reallocation():
    """graph-cache invalidation key of pipeline.GraphedUNet / GraphedWriter"""
    from emote_hack_b200.pipeline import _weights_fingerprint
    m = torch.nn.Linear(4, 4)
    f0 = _weights_fingerprint(m)
    assert _weights_fingerprint(m) == f0
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})     # in-place copy_: versions bump
    f1 = _weights_fingerprint(m)
    assert f1 != f0
    m.double()                                                               # storage re-allocated
    assert _weights_fingerprint(m) != f1


def test_reference_blocks_pairing_order_matches_reference_rule():
    """writer / reader blocks are paired by descending norm1 width with a stable sort (mutual_self_attention.py:585-588)"""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from util_models import TINY_CFG, appearance_cfg
    from emote_hack_b200.appearance_encoder import AppearanceEncoderModel
    from emote_hack_b200.unet3d import UNet3DConditionModel, reference_blocks
    with torch.device("meta"):
        unet, enc = UNet3DConditionModel(**TINY_CFG), AppearanceEncoderModel(**appearance_cfg())
    r, w = reference_blocks(unet), reference_blocks(enc)
    assert len(r) == len(w) == 10
    assert [b.norm1.normalized_shape for b in r] == [b.norm1.normalized_shape for b in w]
    widths = [b.norm1.normalized_shape[0] for b in r]
    assert widths == sorted(widths, reverse=True)
    rn = {id(m): n for n, m in unet.named_modules()}
    wn = {id(m): n for n, m in enc.named_modules()}
    assert [rn[id(b)] for b in r] == [wn[id(b)] for b in w]      # same block names on both sides
    assert len(reference_blocks(unet, "full")) == 16


def test_wav2vec_window_features_match_reference_loop():
    """pipeline.wav2vec_window_features against a literal restatement of the reference loop (Net.py:646-667)"""
    from emote_hack_b200.pipeline import wav2vec_window_features
    g = torch.Generator().manual_seed(0)
    for t, d, m, n in [(9, 6, 2, 2), (3, 4, 2, 2), (1, 4, 2, 2), (7, 5, 0, 3), (6, 8, 1, 0)]:
        hidden = torch.randn(1, t, d, generator=g)
        rows = []
        for f in range(t):                                       # the reference's per-frame cat / pad
            feat = hidden[0, max(f - m, 0):min(f + n + 1, t), :].flatten()
            if f - m < 0:
                feat = torch.cat((torch.zeros((m - f) * d), feat))
            if f + n + 1 > t:
                feat = torch.cat((feat, torch.zeros((f + n + 1 - t) * d)))
            rows.append(feat)
        want = torch.stack(rows)
        assert torch.equal(wav2vec_window_features(hidden, m, n, as_tokens=False), want)
        tok = wav2vec_window_features(hidden[0], m, n)
        assert tok.shape == (t, m + n + 1, d) and torch.equal(tok.reshape(t, -1), want)
    with pytest.raises(ValueError):
        wav2vec_window_features(torch.zeros(2, 3, 4, 5))
