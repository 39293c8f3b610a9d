"""Dev tool: run the tiny UNet twice with identical inputs, report the first modules whose outputs differ."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from util_models import TINY_CFG, make_inputs, rel_l2, rerandomise_zero_inits  # noqa: E402
from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402

torch.manual_seed(0)
m = rerandomise_zero_inits(UNet3DConditionModel(**TINY_CFG).eval()).cuda()
x, ctx = make_inputs(2, 4, 16)
x, ctx = x.cuda(), ctx.cuda()
runs = []
for r in range(2):
    rec = []
    hooks = []
    for name, mod in m.named_modules():
        if name == "":
            continue
        def hook(mod, inp, out, name=name):
            o = out.sample if hasattr(out, "sample") else out
            if isinstance(o, tuple):
                o = o[0]
            if torch.is_tensor(o):
                rec.append((name, o.detach().float().clone()))
        hooks.append(mod.register_forward_hook(hook))
    out = m(x, 481, ctx).sample
    torch.cuda.synchronize()
    for h in hooks:
        h.remove()
    runs.append(rec)
n = 0
for (n1, a), (n2, b) in zip(*runs):
    assert n1 == n2
    e = rel_l2(a, b)
    if e > 1e-6:
        print(f"DIFF {n1}: {e:.3e}")
        n += 1
        if n > 12:
            break
print("modules recorded", len(runs[0]), "differing", n)
