"""Kernel-level numerics: every CUDA entry point against a plain PyTorch fp32 reference of the same op.

Tolerances: operands are rounded to the operand type (fp16 by default, bf16 with EMOTE_OPERAND=bf16) before both paths,
accumulation is fp32, so the only differences are accumulation order (fp32 outputs: plain 2e-5 asserts) and the rounding
of 16-bit outputs (`chk`: limit = 1.5 x the error measured on B200 for this operand type, tests/parity_limits.json).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from emote_hack_b200.ops import OP16   # fp16 (default) or bf16: the tensor-core operand type of this process
from util_models import chk

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops():
    from emote_hack_b200 import ops as o
    return o


def _gen(seed=0):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 320, 320), (300, 160, 192), (2, 1280, 320), (154, 640, 768),
                                   (4096, 960, 320), (1000, 136, 72)])
def test_gemm_plain(ops, M, N, K):
    g = _gen(1)
    a = torch.randn(M, K, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias=bias)
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < 2e-5
    out_b = ops.gemm(a, w, bias=bias, out_dtype=OP16)
    chk(rel_l2(out_b, ref), 4e-3)
@pytest.mark.parametrize("M,N,K,geglu", [(12500, 960, 320, False), (37900, 320, 320, False), (12500, 960, 192, False),
                                         (5000, 2560, 320, True), (4700, 2560, 64, True)])
def test_gemm_weight_stationary(ops, M, N, K, geglu):
    """Shapes routed to gemm_bres_tcgen05_kernel (K <= 320, bf16 staged-store epilogue, many row tiles per CTA),
    ragged M included; pair_mode=3 runs the same problem through the streaming kernel: results must be identical."""
    g = _gen(11)
    a = torch.randn(M, K, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    h = a.float() @ w.float().t() + bias
    if geglu:
        ref = h[:, : N // 2] * F.gelu(h[:, N // 2:])
        wp, bp = ops.pack_geglu(w.float(), bias)
    else:
        ref, wp, bp = h, w, bias
    out = ops.gemm(a, wp, bias=bp, geglu=geglu, out_dtype=OP16)
    chk(rel_l2(out, ref), 4e-3)
    stream = ops.gemm(a, wp, bias=bp, geglu=geglu, out_dtype=OP16, pair_mode=3)
    assert torch.equal(out, stream)


@pytest.mark.parametrize("M,N,K", [(2, 1280, 1280), (1, 320, 1280), (8, 1000, 320), (2, 1280, 320), (3, 64, 2560)])
def test_gemm_skinny_rows(ops, M, N, K):
    """M <= 8 goes to the weight-streaming GEMV kernel (time-embedding products); pair_mode=2 forces the tile kernel."""
    g = _gen(12)
    a = torch.randn(M, K, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    out = ops.gemm(a, w, bias=bias)
    assert rel_l2(out, ref) < 2e-5
    assert rel_l2(ops.gemm(a, w, bias=bias, pair_mode=2), out) < 2e-5
    if N % 8 == 0:
        chk(rel_l2(ops.gemm(a, w, bias=bias, out_dtype=OP16), ref), 4e-3)
    assert rel_l2(ops.gemm(a, w), ref - bias) < 2e-5


def test_gemm_epilogue_residual_rowbias_scale(ops):
    g = _gen(2)
    M, N, K = 512, 320, 640
    a = torch.randn(M, K, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    rb = torch.randn(4, N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    ref = (a.float() @ w.float().t() + bias + rb.repeat_interleave(128, 0) + res) * 0.5
    out = ops.gemm(a, w, bias=bias, row_bias=rb, rows_per_group=128, residual=res, out_scale=0.5)
    assert rel_l2(out, ref) < 2e-5
    # in-place residual (residual aliases out)
    res2 = res.clone()
    ops.gemm(a, w, bias=bias, residual=res2, out=res2)
    assert rel_l2(res2, a.float() @ w.float().t() + bias + res) < 2e-5


@pytest.mark.parametrize("C", [64, 320])
def test_gemm_geglu(ops, C):
    g = _gen(3)
    M, inner = 384, 4 * C
    a = torch.randn(M, C, device="cuda", generator=g).to(OP16)
    w = (torch.randn(2 * inner, C, device="cuda", generator=g) / math.sqrt(C)).to(OP16)
    b = torch.randn(2 * inner, device="cuda", generator=g)
    h = a.float() @ w.float().t() + b
    ref = h[:, :inner] * F.gelu(h[:, inner:])
    wp, bp = ops.pack_geglu(w.float(), b)
    out = ops.gemm(a, wp, bias=bp, geglu=True, out_dtype=OP16)
    assert out.shape == (M, inner)
    chk(rel_l2(out, ref), 4e-3)
@pytest.mark.parametrize("n_img,H,W,C,N", [(2, 8, 8, 64, 128), (2, 16, 16, 128, 64), (3, 32, 32, 64, 320),
                                           (2, 64, 64, 320, 320), (1, 128, 128, 128, 128), (1, 256, 256, 64, 3)])
def test_conv3x3_implicit(ops, n_img, H, W, C, N):
    g = _gen(4)
    x = torch.randn(n_img, H, W, C, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    wp = ops.pack_conv3x3(w.float())
    ld = N if N % 4 == 0 else 4
    out = torch.zeros(n_img * H * W, ld, device="cuda")
    ops.conv3x3(x, wp, n_img, H, W, C, bias=bias, out=out)
    assert rel_l2(out[:, :N], ref) < 2e-5


@pytest.mark.parametrize("n_img,H,W,C,N,mode", [
    (2, 96, 96, 64, 320, "plain"), (4, 48, 48, 128, 160, "res"), (6, 24, 24, 64, 128, "rowbias"), (16, 12, 12, 64, 128, "res"),
    (3, 24, 24, 64, 136, "plain"),       # images not a multiple of the 2 per sub-tile, ragged N
    (5, 12, 12, 128, 64, "res"),         # 8 images per sub-tile, 5 present
    (2, 48, 96, 64, 128, "stats"), (32, 6, 6, 64, 128, "plain"), (2, 96, 96, 512, 160, "res"),   # K = 4608: register epilogue
    (2, 40, 24, 64, 128, "rowbias_small"), (4, 24, 24, 640, 320, "stats")])
def test_conv3x3_implicit_2d_patch_tiles(ops, n_img, H, W, C, N, mode):
    """Implicit-GEMM 3x3 conv over maps whose rows do not pack into 128-pixel runs (BASELINE config #4: 96 / 48 / 24 / 12
    latents): the 128 rows of a sub-tile are a {bw, bh, images} patch, loaded, staged and stored through 4-D TMA boxes —
    no im2col copy.  Single-CTA and CTA-pair kernels, staged (K <= 4096) and register epilogues, fused residual /
    per-sample bias / GroupNorm statistics."""
    g = _gen(40)
    x = torch.randn(n_img, H, W, C, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    wp = ops.pack_conv3x3(w.float())
    M = n_img * H * W
    kw = {}
    if mode == "res":
        res = torch.randn(M, N, device="cuda", generator=g)
        kw = dict(residual=res, out_scale=0.5)
        ref = (ref + res) * 0.5
    elif mode in ("rowbias", "rowbias_small"):
        # per-sample bias: groups of whole images (the time-embedding bias of resnet.py:186-189) / of single images
        per = 2 * H * W if mode == "rowbias" else H * W
        rb = torch.randn(M // per, N, device="cuda", generator=g)
        kw = dict(row_bias=rb, rows_per_group=per)
        ref = ref + rb.repeat_interleave(per, 0)
    elif mode == "stats":
        kw = dict(stats_rows=2 * H * W)
    for pair_mode in (1, 2):
        out = ops.conv3x3(x, wp, n_img, H, W, C, bias=bias, pair_mode=pair_mode, **kw)
        assert rel_l2(out, ref) < 2e-5, pair_mode
        if mode == "stats":
            _check_colstats(out, 2 * H * W)


def test_conv3x3_narrow_channels_take_the_im2col_path(ops):
    """channel counts that are not a multiple of the 64-wide TMA box still go through the explicit im2col gather"""
    g = _gen(5)
    n_img, H, W, C, N = 2, 12, 12, 32, 128
    x = torch.randn(n_img, H, W, C, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(OP16)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    out = ops.conv3x3(x, ops.pack_conv3x3(w.float()), n_img, H, W, C)
    assert rel_l2(out, ref) < 2e-5


def test_downsample_im2col_s2(ops):
    g = _gen(6)
    n_img, H, W, C, N = 2, 16, 16, 64, 128
    x = torch.randn(n_img, H, W, C, device="cuda", generator=g)
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C)).to(OP16)
    ref = F.conv2d(x.to(OP16).float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, N)
    cols = ops.im2col_s2(x, n_img, H, W, C)
    out = ops.gemm(cols, ops.pack_conv3x3(w.float()))
    assert rel_l2(out, ref) < 2e-5


def test_upsample2x(ops):
    g = _gen(7)
    x = torch.randn(2, 8, 8, 64, device="cuda", generator=g)
    out = ops.upsample2x(x, 2, 8, 8, 64).view(2, 16, 16, 64)
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).to(OP16)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("srcs,rows_per_batch,n_batches", [((320,), 1024, 2), ((1280, 640), 256, 2), ((64,), 64, 6),
                                                          ((640, 320), 512, 2), ((128,), 4096, 1)])
def test_group_norm(ops, srcs, rows_per_batch, n_batches):
    g = _gen(8)
    rows = rows_per_batch * n_batches
    xs = [torch.randn(rows, c, device="cuda", generator=g) * 2 + 0.5 for c in srcs]
    ct = sum(srcs)
    gamma = torch.randn(ct, device="cuda", generator=g)
    beta = torch.randn(ct, device="cuda", generator=g)
    out, raw = ops.group_norm(xs, 32, rows_per_batch, n_batches, gamma, beta, 1e-5, True, want_raw=True)
    xc = torch.cat(xs, dim=1).view(n_batches, rows_per_batch, ct).permute(0, 2, 1)  # [nb, C, rows]
    ref = F.silu(F.group_norm(xc, 32, gamma, beta, 1e-5)).permute(0, 2, 1).reshape(rows, ct)
    chk(rel_l2(out, ref), 4e-3)
    assert torch.equal(raw, torch.cat(xs, dim=1).to(OP16))
    out2, _ = ops.group_norm(xs, 32, rows_per_batch, n_batches, gamma, beta, 1e-6, False)
    ref2 = F.group_norm(xc, 32, gamma, beta, 1e-6).permute(0, 2, 1).reshape(rows, ct)
    chk(rel_l2(out2, ref2), 4e-3)
def _check_colstats(out, stats_rows, ops=None):
    """the fused statistics of a GEMM output, folded per statistics batch (all channels as ONE group per column pair is not
    available from the slots directly, so fold them the way group_norm() does: 2 channels per group) against the output"""
    from emote_hack_b200 import ops as _ops
    cs, sr, ver = out._emote_colstats
    assert sr == stats_rows and ver == out._version and cs.dtype == torch.float32
    M, N = out.shape
    nb = M // stats_rows
    G = max(g for g in range(1, 65) if N % g == 0 and (N // g) % 2 == 0)
    sums = torch.empty((nb, G, 2), dtype=torch.float64, device=out.device)
    from emote_hack_b200._lib import check, load
    check(load().emote_gn_colstats_reduce(cs.data_ptr(), N, 0, N, G, stats_rows // 32, nb, sums.data_ptr(), 1,
                                          _ops._stream()), "emote_gn_colstats_reduce")
    o = out.double().view(nb, stats_rows, G, N // G)
    ref = torch.stack([o.sum((1, 3)), (o * o).sum((1, 3))], dim=-1)
    torch.testing.assert_close(sums, ref, rtol=2e-5, atol=1e-3)
    again = torch.empty_like(sums)
    check(load().emote_gn_colstats_reduce(cs.data_ptr(), N, 0, N, G, stats_rows // 32, nb, again.data_ptr(), 1,
                                          _ops._stream()), "emote_gn_colstats_reduce")
    assert torch.equal(again, sums)          # fixed summation order


@pytest.mark.parametrize("M,N,K,sr,mode", [(1024, 320, 320, 256, "res"), (1024, 320, 320, 128, "plain"),
                                           (512, 136, 72, 256, "res"), (2048, 640, 640, 1024, "rowbias"),
                                           (1024, 320, 4608, 512, "res"), (1024, 320, 4608, 512, "plain"),
                                           (896, 1280, 5120, 128, "res")])
def test_gemm_fused_gn_statistics(ops, M, N, K, sr, mode):
    """EmoteGemmArgs.colstats: per-column sum / sum of squares of the fp32 output, per block of stats_rows rows —
    staged-TMA epilogue (K <= 4096), register epilogues (K > 4096), single-CTA and CTA-pair kernels."""
    g = _gen(21)
    a = torch.randn(M, K, device="cuda", generator=g).to(OP16)
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).to(OP16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g) * 2 + 1 if mode == "res" else None
    rb = torch.randn(M // sr, N, device="cuda", generator=g) if mode == "rowbias" else None
    ref = a.float() @ w.float().t() + bias
    if res is not None:
        ref = ref + res
    if rb is not None:
        ref = ref + rb.repeat_interleave(sr, 0)
    for pair_mode in (1, 2):
        out = ops.gemm(a, w, bias=bias, residual=res, row_bias=rb, rows_per_group=sr if rb is not None else 0,
                       stats_rows=sr, pair_mode=pair_mode)
        assert rel_l2(out, ref) < 2e-5
        _check_colstats(out, sr)


def test_group_norm_from_fused_statistics(ops):
    """group_norm() consumes the statistics a conv epilogue accumulated (per frame) for a per-sample GroupNorm and for
    a concatenation with a second source that has none; results match the unfused path."""
    g = _gen(22)
    n_img, H, W, C, N = 8, 16, 16, 64, 320   # 2 samples x 4 frames, 256 rows per frame
    x = torch.randn(n_img, H, W, C, device="cuda", generator=g).to(OP16)
    wp = ops.pack_conv3x3(torch.randn(N, C, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C))
    bias = torch.randn(N, device="cuda", generator=g)
    skip = torch.randn(n_img * H * W, 320, device="cuda", generator=g)
    gamma = torch.randn(640, device="cuda", generator=g)
    beta = torch.randn(640, device="cuda", generator=g)
    fused = ops.conv3x3(x, wp, n_img, H, W, C, bias=bias, stats_rows=H * W)
    _check_colstats(fused, H * W)
    plain = fused.clone()
    assert getattr(plain, "_emote_colstats", None) is None
    for rows_pb, nb in ((4 * H * W, 2), (H * W, n_img)):
        a, _ = ops.group_norm([fused], 32, rows_pb, nb, gamma[:320], beta[:320], 1e-6, True)
        b, _ = ops.group_norm([plain], 32, rows_pb, nb, gamma[:320], beta[:320], 1e-6, True)
        assert rel_l2(a, b) < 1e-3 and (a.float() - b.float()).abs().max() < 0.07
        a, _ = ops.group_norm([fused, skip], 32, rows_pb, nb, gamma, beta, 1e-6, True)
        b, _ = ops.group_norm([plain, skip], 32, rows_pb, nb, gamma, beta, 1e-6, True)
        assert rel_l2(a, b) < 1e-3
        a, _ = ops.group_norm([skip, fused], 32, rows_pb, nb, gamma, beta, 1e-6, False)
        b, _ = ops.group_norm([skip, plain], 32, rows_pb, nb, gamma, beta, 1e-6, False)
        assert rel_l2(a, b) < 1e-3


@pytest.mark.parametrize("C_src,c_off,C_total,groups,slots,nb", [
    (320, 0, 320, 32, 2048, 2),      # 64x64x16-frame level of the UNet: 20480 (slot, column) pairs per block (1024 threads)
    (320, 640, 960, 32, 2048, 2),    # second source of a channel concat: groups straddle / miss the source
    (1280, 0, 1280, 32, 32, 2),      # 8x8 level: 1280 pairs per block (256 threads)
    (128, 0, 128, 32, 8192, 3),      # VAE decoder, one 512x512 frame per statistics batch
    (64, 0, 64, 32, 5, 7)])          # fewer pairs than threads
def test_gn_colstats_reduce_on_synthetic_slots(ops, C_src, c_off, C_total, groups, slots, nb):
    """emote_gn_colstats_reduce alone: random per-quarter slots -> (sum, sum of squares) per (batch, group) in fp64, in
    overwrite and in accumulate mode (the second source of a concatenated GroupNorm input), deterministic."""
    from emote_hack_b200._lib import check, load
    from emote_hack_b200 import ops as _ops
    g = _gen(31)
    cs = torch.randn(nb, slots, C_src, 2, device="cuda", generator=g) * 3 + 1
    cpg = C_total // groups
    ref = torch.zeros(nb, groups, 2, dtype=torch.float64, device="cuda")
    col = cs.double().sum(1)                                  # [nb, C_src, 2]
    for c in range(C_src):
        ref[:, (c_off + c) // cpg] += col[:, c]
    seed = torch.randn(nb, groups, 2, device="cuda", generator=g).double()
    for overwrite in (1, 0):
        sums = seed.clone()
        check(load().emote_gn_colstats_reduce(cs.data_ptr(), C_src, c_off, C_total, groups, slots, nb, sums.data_ptr(),
                                              overwrite, _ops._stream()), "emote_gn_colstats_reduce")
        # overwrite: groups without a column of this source become 0; accumulate: they are left alone
        torch.testing.assert_close(sums, ref if overwrite else ref + seed, rtol=1e-12, atol=1e-9)
        again = seed.clone()
        check(load().emote_gn_colstats_reduce(cs.data_ptr(), C_src, c_off, C_total, groups, slots, nb, again.data_ptr(),
                                              overwrite, _ops._stream()), "emote_gn_colstats_reduce")
        assert torch.equal(again, sums)


@pytest.mark.parametrize("C", [64, 320, 1280])
def test_layer_norm(ops, C):
    g = _gen(9)
    F_, HW, B = 4, 16, 2
    M = B * F_ * HW
    x = torch.randn(M, C, device="cuda", generator=g) * 3 + 1
    gamma = torch.randn(C, device="cuda", generator=g)
    beta = torch.randn(C, device="cuda", generator=g)
    out = ops.layer_norm(x, gamma, beta)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    chk(rel_l2(out, ref), 4e-3)
    pe = torch.randn(24, C, device="cuda", generator=g)
    out2 = ops.layer_norm(x, gamma, beta, pe=pe, rows_per_frame=HW, frames=F_)
    fidx = (torch.arange(M, device="cuda") // HW) % F_
    chk(rel_l2(out2, ref + pe[fidx]), 4e-3)
def _sdpa_ref(q, k, v, scale):
    s = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) * scale
    return torch.einsum("bhqk,bhkd->bhqd", s.softmax(-1), v.float())


@pytest.mark.parametrize("heads,d,nq,nk", [(8, 40, 256, 256), (8, 80, 64, 64), (8, 160, 64, 64), (4, 16, 64, 64),
                                           (8, 40, 100, 77), (8, 40, 4096, 4096), (4, 64, 16, 16), (2, 32, 200, 5),
                                           # short key sets with many queries -> resident-K/V streaming kernel
                                           (8, 40, 4096, 77), (8, 80, 1024, 77), (8, 160, 256, 5), (4, 64, 300, 128),
                                           (8, 40, 1000, 16), (2, 16, 257, 33)])
def test_flash_attention_fused_qkv(ops, heads, d, nq, nk):
    g = _gen(10)
    batch = 3
    C = heads * d
    self_attn = nq == nk
    if self_attn:
        qkv = torch.randn(batch, nq, 3 * C, device="cuda", generator=g).to(OP16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        kvs = (nq * 3 * C, 3 * C)
        qs = kvs
    else:
        qt = torch.randn(batch, nq, C, device="cuda", generator=g).to(OP16)
        kv = torch.randn(batch, nk, 2 * C, device="cuda", generator=g).to(OP16)
        q, k, v = qt, kv[..., :C], kv[..., C:]
        qs, kvs = (nq * C, C), (nk * 2 * C, 2 * C)
    out = torch.empty(batch, nq, C, device="cuda", dtype=OP16)
    ops.attention(q, k, v, out, batch=batch, heads=heads, head_dim=d, nq=nq, n0=nk, q_strides=qs, kv0_strides=kvs,
                  o_strides=(nq * C, C), scale=d ** -0.5)
    sp = lambda t, n: t.reshape(batch, n, heads, d).permute(0, 2, 1, 3)
    ref = _sdpa_ref(sp(q, nq), sp(k, nk), sp(v, nk), d ** -0.5).permute(0, 2, 1, 3).reshape(batch, nq, C)
    chk(rel_l2(out, ref), 6e-3)
def test_flash_attention_two_segments_cfg(ops):
    """reference attention: K/V = [self | bank], bank only visible to the conditional half, shared across frames."""
    g = _gen(11)
    heads, d, n, F_ = 8, 40, 64, 4
    C = heads * d
    batch = 2 * F_
    qkv = torch.randn(batch, n, 3 * C, device="cuda", generator=g).to(OP16)
    bank_kv = torch.randn(2, n, 2 * C, device="cuda", generator=g).to(OP16)
    out = torch.empty(batch, n, C, device="cuda", dtype=OP16)
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], out, batch=batch, heads=heads, head_dim=d, nq=n,
                  n0=n, q_strides=(n * 3 * C, 3 * C), kv0_strides=(n * 3 * C, 3 * C), o_strides=(n * C, C),
                  scale=d ** -0.5, k1=bank_kv[..., :C], v1=bank_kv[..., C:], n1=n, kv1_strides=(n * 2 * C, 2 * C),
                  kv1_batch_div=F_, kv1_first_batch=F_)
    sp = lambda t, nn_: t.reshape(t.shape[0], nn_, heads, d).permute(0, 2, 1, 3)
    q, k, v = sp(qkv[..., :C], n), sp(qkv[..., C:2 * C], n), sp(qkv[..., 2 * C:], n)
    ref_uc = _sdpa_ref(q[:F_], k[:F_], v[:F_], d ** -0.5)
    bk = sp(bank_kv[1:2, :, :C].expand(F_, -1, -1), n)
    bv = sp(bank_kv[1:2, :, C:].expand(F_, -1, -1), n)
    ref_c = _sdpa_ref(q[F_:], torch.cat([k[F_:], bk], 2), torch.cat([v[F_:], bv], 2), d ** -0.5)
    ref = torch.cat([ref_uc, ref_c]).permute(0, 2, 1, 3).reshape(batch, n, C)
    chk(rel_l2(out, ref), 6e-3)
@pytest.mark.parametrize("heads,d,F_,HW", [(8, 40, 16, 64), (8, 160, 16, 16), (4, 16, 8, 64), (8, 80, 24, 16),
                                           (8, 40, 32, 16), (4, 64, 1, 16)])
def test_temporal_attention(ops, heads, d, F_, HW):
    g = _gen(12)
    B = 2
    C = heads * d
    qkv = torch.randn(B, F_, HW, 3 * C, device="cuda", generator=g).to(OP16)
    out = ops.temporal_attention(qkv.view(-1, 3 * C), B, F_, HW, heads, d).view(B, F_, HW, C)
    sp = lambda t: t.permute(0, 2, 1, 3).reshape(B * HW, F_, heads, d).permute(0, 2, 1, 3)
    ref = _sdpa_ref(sp(qkv[..., :C]), sp(qkv[..., C:2 * C]), sp(qkv[..., 2 * C:]), d ** -0.5)
    ref = ref.permute(0, 2, 1, 3).reshape(B, HW, F_, C).permute(0, 2, 1, 3)
    chk(rel_l2(out, ref), 6e-3)
def test_tuning_knobs_change_launch_geometry_not_results(ops):
    """emote_set_tuning: heads per temporal-attention block, rows per LayerNorm block, gn_apply block count — the same bits
    under every setting (the statistics-fold knob changes the fp64 summation order and is compared by value in
    test_gn_colstats_reduce_on_synthetic_slots)."""
    from emote_hack_b200._lib import check, load
    g = _gen(33)
    B, F_, HW, heads, d = 2, 16, 64, 8, 40
    C = heads * d
    qkv = torch.randn(B * F_ * HW, 3 * C, device="cuda", generator=g).to(OP16)
    x = torch.randn(1000, 1280, device="cuda", generator=g) * 2 + 1
    gamma, beta = torch.randn(1280, device="cuda", generator=g), torch.randn(1280, device="cuda", generator=g)
    pe = torch.randn(24, 1280, device="cuda", generator=g)
    xs = x.view(4, 250, 32, 40).double()
    sums = torch.stack([xs.sum((1, 3)), (xs * xs).sum((1, 3))], dim=-1).contiguous()

    def gn_apply():
        out = torch.empty(1000, 1280, dtype=OP16, device="cuda")
        check(load().emote_gn_apply(x.data_ptr(), 1280, 0, 1280, 32, 250, 4, sums.data_ptr(), gamma.data_ptr(),
                                    beta.data_ptr(), 1e-5, 1, out.data_ptr(), None, ops._stream()), "emote_gn_apply")
        return out

    try:
        ref_t = ops.temporal_attention(qkv, B, F_, HW, heads, d)
        ref_l = ops.layer_norm(x, gamma, beta, pe=pe, rows_per_frame=10, frames=20)
        ref_g = gn_apply()
        want = F.silu(F.group_norm(x.view(4, 250, 1280).permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
        chk(rel_l2(ref_g, want.reshape(1000, 1280)), 4e-3)
        ops.set_tuning("temporal_warps", 8)
        assert torch.equal(ops.temporal_attention(qkv, B, F_, HW, heads, d), ref_t)
        for w in (1, 2, 8):
            ops.set_tuning("ln_warps", w)
            assert torch.equal(ops.layer_norm(x, gamma, beta, pe=pe, rows_per_frame=10, frames=20), ref_l)
        for tb in (1, 37, 296):
            ops.set_tuning("gn_apply_blocks", tb)
            assert torch.equal(gn_apply(), ref_g)
    finally:
        for k in ("temporal_warps", "ln_warps", "gn_apply_blocks"):
            ops.set_tuning(k, -1)


def test_softmax_rows(ops):
    g = _gen(13)
    s = torch.randn(300, 4096, device="cuda", generator=g) * 4
    out = ops.softmax_rows(s, 0.25)
    chk(rel_l2(out, (s * 0.25).softmax(-1)), 4e-3)
def test_latent_im2col_and_small_conv(ops):
    g = _gen(14)
    B, Cl, F_, H, W, N = 2, 4, 3, 8, 8, 64
    lat = torch.randn(B, Cl, F_, H, W, device="cuda", generator=g)
    w = torch.randn(N, Cl, 3, 3, device="cuda", generator=g) * 0.2
    bias = torch.randn(N, device="cuda", generator=g)
    cols = ops.latent_im2col(lat)
    out = ops.gemm(cols, ops.pack_conv3x3_small(w), bias=bias)
    x2d = lat.permute(0, 2, 1, 3, 4).reshape(B * F_, Cl, H, W)
    ref = F.conv2d(x2d.to(OP16).float(), w.to(OP16).float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    assert rel_l2(out, ref) < 2e-5
    # with pre-scale + pointwise linear (VAE post_quant_conv folded in front)
    pw = torch.randn(Cl, Cl, device="cuda", generator=g)
    pb = torch.randn(Cl, device="cuda", generator=g)
    cols2 = ops.latent_im2col(lat, pre_scale=1 / 0.18215, pw_weight=pw, pw_bias=pb)
    z = F.conv2d(x2d / 0.18215, pw.view(Cl, Cl, 1, 1), pb)
    ref2 = F.conv2d(z.to(OP16).float(), w.to(OP16).float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    out2 = ops.gemm(cols2, ops.pack_conv3x3_small(w), bias=bias)
    chk(rel_l2(out2, ref2), 3e-3)
def test_layout_and_misc(ops):
    g = _gen(15)
    B, Cc, F_, H, W = 2, 4, 3, 8, 8
    x = torch.randn(B, Cc, F_, H, W, device="cuda", generator=g)
    tok = ops.ncfhw_to_tokens(x)
    assert torch.equal(tok.view(B, F_, H, W, Cc), x.permute(0, 2, 3, 4, 1))
    assert torch.equal(ops.tokens_to_ncfhw(tok, B, Cc, F_, H, W), x)
    y = torch.randn_like(x)
    assert torch.equal(ops.add_f32(x, y), x + y)
    chk(rel_l2(ops.silu_bf16(x), F.silu(x)), 4e-3)
    wide = torch.randn(64, 128, device="cuda", generator=g)
    dst = torch.zeros(64, 256, device="cuda", dtype=OP16)
    ops.cast_bf16(wide, dst, c_offset=128)
    assert torch.equal(dst[:, 128:], wide.to(OP16)) and dst[:, :128].abs().sum() == 0


def test_timestep_embedding(ops):
    t = torch.tensor([981.0, 1.0, 500.0], device="cuda")
    out = ops.timestep_embedding(t, 320, True, 0.0)
    half = 160
    e = torch.exp(-math.log(10000.0) * torch.arange(half, device="cuda", dtype=torch.float32) / half)
    emb = t[:, None] * e[None]
    ref = torch.cat([torch.cos(emb), torch.sin(emb)], -1)
    assert (out.float() - ref).abs().max() < 1.5e-2  # bf16 output + large-argument sin/cos


def test_window_gather_scatter_and_fill(ops):
    """emote_gather_frames / emote_scatter_add_frames / emote_fill_f32 against torch indexing: the `latents[:, :, c]`
    gather + CFG repeat, the per-frame audio-token gather and `noise_pred[:, :, c] += pred` of the denoise loop, bit-exact
    (wrapping windows, strided windows, single-branch offsets)."""
    g = _gen(31)
    C, F_, H, W = 4, 24, 8, 8
    lat = torch.randn(1, C, F_, H, W, device="cuda", generator=g)
    for win in ([0, 1, 2, 3, 4, 5, 6, 7], [20, 21, 22, 23, 0, 1, 2, 3], [1, 5, 9, 13, 17, 21]):
        idx = torch.tensor(win, dtype=torch.int32, device="cuda")
        for nb in (1, 2):
            dst = torch.empty(nb, C, len(win), H, W, device="cuda")
            ops.gather_frames(lat, dst, idx, nb * C, F_, H * W, src_mod=C)
            assert torch.equal(dst, lat[:, :, win].repeat(nb, 1, 1, 1, 1))
        ctx = torch.randn(2 * F_, 5, 64, device="cuda", generator=g)          # [uncond frames | cond frames]
        for nb, b0 in ((2, 0), (1, 0), (1, 1)):
            cdst = torch.empty(nb * len(win), 5, 64, device="cuda")
            ops.gather_frames(ctx, cdst, idx, nb, F_, 5 * 64, src_mod=2, src_off=b0)
            want = torch.cat([ctx[(b0 + b) * F_:(b0 + b + 1) * F_][win] for b in range(nb)])
            assert torch.equal(cdst, want)
        acc = torch.randn(2, C, F_, H, W, device="cuda", generator=g)
        for nb, b0 in ((2, 0), (1, 0), (1, 1)):
            pred = torch.randn(nb, C, len(win), H, W, device="cuda", generator=g)
            want = acc.clone()
            want[b0:b0 + nb, :, win] += pred
            ops.scatter_add_frames(pred, acc, idx, nb * C, F_, H * W, dst_off=b0 * C)
            assert torch.equal(acc, want)
    t = torch.empty(5, device="cuda")
    assert torch.equal(ops.fill_f32(t, 481.0), torch.full((5,), 481.0, device="cuda"))
    with pytest.raises(Exception):
        ops.gather_frames(lat, torch.empty(1, C, 2, H, W, device="cuda"), torch.zeros(3, dtype=torch.int32, device="cuda"),
                          C, F_, H * W, src_mod=C)


def test_ddim_step_plain_and_accumulator_clear(ops):
    g = _gen(17)
    x = torch.randn(3, 4, 8, 8, device="cuda", generator=g)                 # any rank: the plain step is elementwise
    eps = torch.randn(3, 4, 8, 8, device="cuda", generator=g)
    a_t, a_p = 0.41, 0.63
    want = a_p ** 0.5 * (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5 + (1 - a_p) ** 0.5 * eps
    assert rel_l2(ops.ddim_step(x.clone(), eps, a_t, a_p), want) < 1e-6
    lat = torch.randn(1, 4, 6, 8, 8, device="cuda", generator=g)
    npred = torch.randn(2, 4, 6, 8, 8, device="cuda", generator=g)
    keep = ops.cfg_ddim_step(lat.clone(), npred.clone(), None, 7.5, a_t, a_p)
    acc = npred.clone()
    out = ops.cfg_ddim_step(lat.clone(), acc, None, 7.5, a_t, a_p, zero_noise_pred=True)
    assert torch.equal(out, keep) and acc.abs().sum() == 0            # consumed and cleared for the next timestep
    with pytest.raises(Exception):
        ops.cfg_ddim_step(lat.clone(), npred, None, 7.5, a_t, a_p, sigma=0.9)    # sigma^2 > 1 - a_prev / no noise tensor


def test_cfg_ddim_step(ops):
    g = _gen(16)
    lat = torch.randn(1, 4, 6, 8, 8, device="cuda", generator=g)
    npred = torch.randn(2, 4, 6, 8, 8, device="cuda", generator=g)
    cnt = torch.tensor([1, 2, 1, 1, 2, 1], device="cuda", dtype=torch.float32)
    a_t, a_p, gs = 0.37, 0.52, 7.5
    e = npred / cnt.view(1, 1, 6, 1, 1)
    eps = e[0:1] + gs * (e[1:2] - e[0:1])
    x0 = (lat - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    ref = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps
    out = ops.cfg_ddim_step(lat.clone(), npred, cnt, gs, a_t, a_p)
    assert rel_l2(out, ref) < 1e-5


def test_video_grid_frames_match_reference_writer_input(ops, tmp_path):
    """emote_video_grid_u8 / save_videos_grid: byte-exact against the frames the reference's save_videos_grid
    (utils/util.py:21-33) handed to imageio (tests/golden/video_grid.pt), and against the numpy oracle on a 512 x 512 clip"""
    from pathlib import Path
    import numpy as np
    from emote_hack_b200.magicanimate.utils import util
    from oracle import video_grid
    gold = torch.load(Path(__file__).parent / "golden" / "video_grid.pt")
    for case in gold["cases"]:
        v = torch.rand(*case["shape"], generator=torch.Generator().manual_seed(case["seed"]))
        v = v * 2 - 1 if case["rescale"] else v
        got = util.video_frames_u8(v.cuda(), case["rescale"], case["n_rows"])
        assert got.shape == case["frames"].shape and torch.equal(got.cpu(), case["frames"]), case["shape"]
    v = torch.rand(2, 3, 4, 512, 512, generator=torch.Generator().manual_seed(5))
    v[0, 0, 0, 0, :4] = torch.tensor([1.0, 0.0, 254.999 / 255, 1e-9])
    got = util.video_frames_u8(v.cuda(), False, 6).cpu().numpy()
    assert np.array_equal(got, video_grid.video_frames_u8(v.numpy(), False, 6))
    assert util.save_videos_grid(v.cuda(), str(tmp_path / "out" / "clip.npy")) in ("numpy", "imageio")
    back = util.video2images(str(tmp_path / "out" / "clip.npy"), step=1, length=8)
    assert len(back) == 4 and np.array_equal(np.stack(back), got)
    util.save_images_grid(v[:, :, :1].cuda(), str(tmp_path / "grid.png"))
    from PIL import Image
    png = np.asarray(Image.open(tmp_path / "grid.png"))
    assert np.array_equal(png, video_grid.video_frames_u8(v[:, :, :1].numpy(), False, 8)[0])
    with pytest.raises(Exception):
        ops.video_grid_u8(torch.zeros(1, 2, 1, 4, 4, device="cuda"))


def test_vae_postprocess(ops):
    g = _gen(17)
    tok = torch.randn(2 * 64, 4, device="cuda", generator=g)
    of, ou = ops.vae_postprocess(tok, 2, 8, 8, want_f32=True, want_u8=True)
    ref = (tok[:, :3].view(2, 64, 3).permute(0, 2, 1).reshape(2, 3, 8, 8) / 2 + 0.5).clamp(0, 1)
    assert torch.allclose(of, ref, atol=1e-6)
    assert torch.equal(ou, (of * 255).to(torch.uint8))   # the reference's truncating cast (utils/util.py:28)
    assert (ou.float() - ref * 255).abs().max() <= 1.0 + 1e-3


def test_errors_are_loud(ops):
    from emote_hack_b200._lib import EmoteKernelError
    a = torch.zeros(16, 12, device="cuda", dtype=OP16)
    w = torch.zeros(16, 12, device="cuda", dtype=OP16)
    with pytest.raises(EmoteKernelError):
        ops.gemm(a, w)  # K not a multiple of 8
    with pytest.raises(EmoteKernelError):
        ops.gemm(a.cpu(), w)


@pytest.mark.parametrize("batch,nq,nk,ramp", [(2, 4096, 4096, False), (3, 300, 300, False), (1, 64, 64, False),
                                              (2, 1024, 1024, True), (1, 200, 333, False)])
def test_attention_wide_512_single_head(ops, batch, nq, nk, ramp):
    """emote_attention_wide_bf16: the VAE mid-block attention (one 512-dim head, orig_attention.py:360-376) as a tcgen05
    flash kernel batched over images — fused-QKV strides, ragged tails, and rising scores that force the lazy rescale."""
    g = _gen(24)
    C = 512
    if nq == nk:
        qkv = torch.randn(batch, nq, 3 * C, device="cuda", generator=g)
        if ramp:   # keys whose scores rise along the key axis: every tile beats the previous maximum by more than 2^8
            u = torch.nn.functional.normalize(torch.randn(C, device="cuda", generator=g), dim=0)
            qkv[..., :C] += 6.0 * u * C ** 0.25
            qkv[..., C:2 * C] += torch.linspace(0, 40, nk, device="cuda")[None, :, None] * u * C ** 0.25
        qkv = qkv.to(OP16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        qs = kvs = (nq * 3 * C, 3 * C)
    else:
        q = torch.randn(batch, nq, C, device="cuda", generator=g).to(OP16)
        kv = torch.randn(batch, nk, 2 * C, device="cuda", generator=g).to(OP16)
        k, v = kv[..., :C], kv[..., C:]
        qs, kvs = (nq * C, C), (nk * 2 * C, 2 * C)
    out = torch.zeros(batch, nq, C, device="cuda", dtype=OP16)
    ops.attention_wide(q, k, v, out, batch=batch, nq=nq, nk=nk, q_strides=qs, kv_strides=kvs, o_strides=(nq * C, C),
                       scale=C ** -0.5)
    ref = _sdpa_ref(q[:, None].float(), k[:, None].float(), v[:, None].float(), C ** -0.5)[:, 0]
    assert torch.isfinite(out.float()).all()
    chk(rel_l2(out, ref), 6e-3)


def test_flash_attention_tcgen05_lazy_rescale_with_growing_scores(ops):
    """The tcgen05 kernel keeps O in TMEM and rescales it only when a row's running maximum has grown by more than 2^8:
    keys whose scores rise steadily along the key axis (every tile beats the previous maximum by far more than that) force
    the rescale path on every tile, a row-dependent ramp makes the warp-uniform decision fire for rows that do not need
    it, and a flat tail exercises the no-rescale path after many rescales."""
    g = _gen(23)
    batch, heads, d, n = 2, 2, 40, 1024
    C = heads * d
    u = torch.nn.functional.normalize(torch.randn(d, device="cuda", generator=g), dim=0)
    ramp = torch.cat([torch.linspace(0, 60, n - 256, device="cuda"), torch.full((256,), 60.0, device="cuda")])
    k = ramp[:, None] * u[None] + 0.3 * torch.randn(n, d, device="cuda", generator=g)
    qscale = torch.linspace(0.5, 4.0, n, device="cuda")                       # rows see ramps of different steepness
    q = qscale[:, None] * u[None] * d ** 0.5 + 0.3 * torch.randn(n, d, device="cuda", generator=g)
    v = torch.randn(n, d, device="cuda", generator=g)
    rep = lambda t: t[None, :, None, :].expand(batch, n, heads, d).reshape(batch, n, C).to(OP16).contiguous()
    qq, kk, vv = rep(q), rep(k), rep(v)
    out = torch.zeros(batch, n, C, device="cuda", dtype=OP16)
    ops.attention(qq, kk, vv, out, batch=batch, heads=heads, head_dim=d, nq=n, n0=n, q_strides=(n * C, C),
                  kv0_strides=(n * C, C), o_strides=(n * C, C), scale=d ** -0.5, impl="tc")
    sp = lambda t: t.reshape(batch, n, heads, d).permute(0, 2, 1, 3)
    ref = _sdpa_ref(sp(qq), sp(kk), sp(vv), d ** -0.5).permute(0, 2, 1, 3).reshape(batch, n, C)
    assert torch.isfinite(out.float()).all()
    chk(rel_l2(out, ref), 6e-3)


@pytest.mark.parametrize("heads,d,nq,nk", [(8, 40, 256, 256), (8, 40, 1024, 1024), (8, 80, 256, 256), (2, 40, 200, 300),
                                           (8, 40, 4096, 4096), (4, 80, 1024, 1024), (1, 40, 64, 64),
                                           (8, 160, 256, 256), (12, 64, 499, 499), (2, 160, 100, 333), (3, 64, 1024, 77)])
def test_flash_attention_tcgen05(ops, heads, d, nq, nk):
    g = _gen(20)
    batch = 3
    C = heads * d
    if nq == nk:
        qkv = torch.randn(batch, nq, 3 * C, device="cuda", generator=g).to(OP16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        qs = kvs = (nq * 3 * C, 3 * C)
    else:
        qt = torch.randn(batch, nq, C, device="cuda", generator=g).to(OP16)
        kv = torch.randn(batch, nk, 2 * C, device="cuda", generator=g).to(OP16)
        q, k, v = qt, kv[..., :C], kv[..., C:]
        qs, kvs = (nq * C, C), (nk * 2 * C, 2 * C)
    out = torch.zeros(batch, nq, C, device="cuda", dtype=OP16)
    ops.attention(q, k, v, out, batch=batch, heads=heads, head_dim=d, nq=nq, n0=nk, q_strides=qs, kv0_strides=kvs,
                  o_strides=(nq * C, C), scale=d ** -0.5, impl="tc")
    sp = lambda t, n: t.reshape(batch, n, heads, d).permute(0, 2, 1, 3)
    ref = _sdpa_ref(sp(q, nq), sp(k, nk), sp(v, nk), d ** -0.5).permute(0, 2, 1, 3).reshape(batch, nq, C)
    e = rel_l2(out, ref)
    print(f"tc attention heads={heads} d={d} nq={nq} nk={nk}: rel_l2={e:.2e}")
    chk(e, 6e-3)


def test_flash_attention_tcgen05_two_segments_cfg(ops):
    g = _gen(21)
    heads, d, n, F_ = 8, 40, 256, 4
    C = heads * d
    batch = 2 * F_
    qkv = torch.randn(batch, n, 3 * C, device="cuda", generator=g).to(OP16)
    bank_kv = torch.randn(2, n, 2 * C, device="cuda", generator=g).to(OP16)
    outs = []
    for impl in ("tc", "mma"):
        out = torch.zeros(batch, n, C, device="cuda", dtype=OP16)
        ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], out, batch=batch, heads=heads, head_dim=d,
                      nq=n, n0=n, q_strides=(n * 3 * C, 3 * C), kv0_strides=(n * 3 * C, 3 * C), o_strides=(n * C, C),
                      scale=d ** -0.5, k1=bank_kv[..., :C], v1=bank_kv[..., C:], n1=n, kv1_strides=(n * 2 * C, 2 * C),
                      kv1_batch_div=F_, kv1_first_batch=F_, impl=impl)
        outs.append(out)
    chk(rel_l2(outs[0], outs[1]), 6e-3)