"""Dev tool: sweeps of the emote_set_tuning knobs on the UNet's shapes, each timed as launches replayed from one CUDA graph
— "gn_reduce" (1 = flat statistics fold, 0 = per-slot walk), "gn_apply_blocks" (target block count of gn_apply),
"ln_warps" (rows per LayerNorm block), "temporal_warps" (heads per temporal-attention block) — then one full-size UNet
call under the settings named on the command line (default: library defaults vs the per-slot fold)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from emote_hack_b200 import _lib, ops  # noqa: E402
from emote_hack_b200._lib import check  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda")


def graph_time(fn, n=40, reps=5):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)   # us per launch


# (C_src, c_offset, C_total, slots_per_batch, n_batches): 5-D GroupNorm of the resnets (2 samples x 16 frames), the per-frame
# GroupNorm of Transformer3DModel (32 frames), second source of an up-block concat
shapes = [(320, 0, 320, 2048, 2), (640, 0, 640, 512, 2), (1280, 0, 1280, 128, 2), (1280, 0, 1280, 32, 2),
          (320, 0, 320, 128, 32), (640, 0, 640, 32, 32), (1280, 0, 1280, 8, 32), (320, 320, 640, 2048, 2),
          (640, 1280, 1920, 128, 2)]
print(f"operand={_lib.OPERAND}")
for (cs, off, ct, slots, nb) in shapes:
    sl = torch.randn(nb, slots, cs, 2, device=dev)
    res = {}
    for mode in (0, 1):
        ops.set_tuning("gn_reduce", mode)
        sums = torch.zeros(nb, 32, 2, dtype=torch.float64, device=dev)
        fn = lambda: check(lib.emote_gn_colstats_reduce(sl.data_ptr(), cs, off, ct, 32, slots, nb, sums.data_ptr(), 1,
                                                        torch.cuda.current_stream().cuda_stream), "reduce")
        us = graph_time(fn)
        res[mode] = (us, sums.clone())
    d = (res[0][1] - res[1][1]).abs().max().item()
    print(f"reduce C_src={cs} off={off} C_total={ct} slots={slots} batches={nb}: per-slot {res[0][0]:.2f} us, flat {res[1][0]:.2f} us, "
          f"max |diff| {d:.2e}")

# gn_apply on the same levels (rows_per_batch, n_batches, C); -1 = library default (2368 blocks, <= 4 load batches per block)
for (rows, nb, c) in [(65536, 2, 320), (16384, 2, 640), (4096, 2, 1280), (1024, 2, 1280), (4096, 32, 320), (1024, 32, 640),
                      (256, 32, 1280), (64, 32, 1280)]:
    x = torch.randn(rows * nb, c, device=dev)
    gamma, beta = torch.randn(c, device=dev), torch.randn(c, device=dev)
    sums = torch.empty(nb, 32, 2, dtype=torch.float64, device=dev)
    xs = x.view(nb, rows, 32, c // 32).double()
    sums[..., 0], sums[..., 1] = xs.sum((1, 3)), (xs * xs).sum((1, 3))
    out = torch.empty(rows * nb, c, dtype=ops.OP16, device=dev)
    fn = lambda: check(lib.emote_gn_apply(x.data_ptr(), c, 0, c, 32, rows, nb, sums.data_ptr(), gamma.data_ptr(),
                                          beta.data_ptr(), 1e-5, 1, out.data_ptr(), None,
                                          torch.cuda.current_stream().cuda_stream), "apply")
    gb = rows * nb * c * 6 / 1e9
    line = []
    for tb in (-1, 2368, 1184, 592, 296):
        ops.set_tuning("gn_apply_blocks", tb)
        us = graph_time(fn, n=10)
        line.append(f"{tb}: {us:.2f} us ({gb / us * 1e6:.0f} GB/s)")
    ops.set_tuning("gn_apply_blocks", -1)
    print(f"gn_apply rows/batch={rows} batches={nb} C={c}: " + ", ".join(line))

# LayerNorm (M, C, temporal PE?) x rows per block
for (M, c, pe_on) in [(131072, 320, False), (131072, 320, True), (32768, 640, False), (32768, 640, True), (8192, 1280, False),
                      (8192, 1280, True), (2048, 1280, False)]:
    x = torch.randn(M, c, device=dev)
    gamma, beta = torch.randn(c, device=dev), torch.randn(c, device=dev)
    pe = torch.randn(24, c, device=dev) if pe_on else None
    fn = lambda: ops.layer_norm(x, gamma, beta, pe=pe, rows_per_frame=M // 32, frames=16)
    gb = M * c * 6 / 1e9
    line = []
    for w in (2, 4, 8):
        ops.set_tuning("ln_warps", w)
        us = graph_time(fn, n=10)
        line.append(f"{w} warps: {us:.2f} us ({gb / us * 1e6:.0f} GB/s)")
    ops.set_tuning("ln_warps", -1)
    print(f"layernorm M={M} C={c} pe={pe_on}: " + ", ".join(line))

# temporal attention (B, F, HW, heads, d)
for (B, F_, HW, heads, d) in [(2, 16, 4096, 8, 40), (2, 16, 1024, 8, 80), (2, 16, 256, 8, 160), (2, 16, 64, 8, 160)]:
    C_ = heads * d
    qkv = torch.randn(B * F_ * HW, 3 * C_, device=dev).to(ops.OP16)
    out = torch.empty(B * F_ * HW, C_, dtype=ops.OP16, device=dev)
    fn = lambda: check(lib.emote_temporal_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, F_, HW, heads, d, d ** -0.5,
                                                         torch.cuda.current_stream().cuda_stream), "temporal")
    gb = B * F_ * HW * C_ * 8 / 1e9
    line, outs = [], {}
    for w in (4, 8):
        ops.set_tuning("temporal_warps", w)
        us = graph_time(fn, n=10)
        outs[w] = out.clone()
        line.append(f"{w} heads/block: {us:.2f} us ({gb / us * 1e6:.0f} GB/s)")
    ops.set_tuning("temporal_warps", -1)
    print(f"temporal B={B} F={F_} HW={HW} heads={heads} d={d}: " + ", ".join(line) +
          f", equal={torch.equal(outs[4], outs[8])}")

if "--no-unet" not in sys.argv:
    from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
    from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402
    torch.manual_seed(0)
    with torch.device(dev):
        unet = UNet3DConditionModel(**FULL_CFG).eval()
    rerandomise_zero_inits(unet)
    x = torch.randn(2, 4, 16, 64, 64, device=dev); ctx = torch.randn(2, 77, 768, device=dev)
    # settings to compare at UNet level: "k=v,k=v" arguments; "" = library defaults
    settings = [a for a in sys.argv[1:] if not a.startswith("--")] or ["", "gn_reduce=0"]
    keys = ("gn_reduce", "gn_apply_blocks", "ln_warps", "temporal_warps")
    outs = {}
    for st in settings * 2:
        for k in keys:
            ops.set_tuning(k, -1 if k != "gn_reduce" else 1)
        for kv in filter(None, st.split(",")):
            k, v = kv.split("=")
            ops.set_tuning(k, int(v))
        for _ in range(2):
            unet(x, 981, ctx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            o = unet(x, 981, ctx).sample
        e1.record(); torch.cuda.synchronize()
        outs[st] = o.clone()
        print(f"UNet call [{st or 'defaults'}]: {e0.elapsed_time(e1) / 5:.2f} ms")
    base = outs[settings[0]]
    for st in settings[1:]:
        print(f"[{st}] vs [{settings[0] or 'defaults'}] rel diff", ((outs[st] - base).norm() / base.norm()).item())
    for k in keys:
        ops.set_tuning(k, -1 if k != "gn_reduce" else 1)
