#!/usr/bin/env python
"""Build-container measurement (needs /root/reference): seconds of ONE conditional-branch UNet3D call [1,4,16,64,64]
(SD-1.5 widths, motion modules on, 77 text tokens, fp32, all host threads) through
  (a) the UNTOUCHED reference modules (imported via oracle/ref_shim.py, unsliced attention), and
  (b) the oracle port bench.py times as its CPU arm,
with identical seeded weights, plus the output agreement.  Shows that the port is a fair stand-in for the reference's CPU
path on the GPU box, where /root/reference does not exist.  Writes profiles/r02_cpu_port_vs_reference.json."""
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import ref_shim  # noqa: E402
from oracle.unet3d_port import UNet3DOracle  # noqa: E402
from util_models import FULL_CFG, seeded_unet_state_dict  # noqa: E402


def main(frames: int):
    threads = len(os.sched_getaffinity(0))
    torch.set_num_threads(threads)
    shapes = json.loads((ROOT / "tests" / "golden" / "unet3d_full_keys.json").read_text())
    sd = seeded_unet_state_dict(shapes, seed=0)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 4, frames, 64, 64, generator=g)
    ctx = torch.randn(1, 77, 768, generator=g)
    U = ref_shim.load_reference_unet_class()
    ref = U(**FULL_CFG).eval()
    ref.load_state_dict(sd, strict=True)
    # no set_attention_slice here: with motion modules on, the reference's sliced path allocates
    # [b*d*heads, d, c/heads] (motion_module.py:276,323 passes the PRE-rearrange sequence length to
    # orig_attention.py:686-690), 21 GB per temporal attention at one frame — the unsliced scores (8.6 GB fp32 per
    # softmax operand at 16 frames) fit the build container's 62 GB
    with torch.no_grad():
        ref(x[:, :, :1], torch.tensor(981), ctx)                      # warm the thread pool / allocator
        t0 = time.perf_counter(); want = ref(x, torch.tensor(981), ctx).sample; t_ref = time.perf_counter() - t0
    del ref
    port = UNet3DOracle(sd, FULL_CFG, attention_slice_bytes=1 << 30)
    port(x[:, :, :1], 981, ctx)
    t0 = time.perf_counter(); got = port(x, 981, ctx); t_port = time.perf_counter() - t0
    rel = ((got - want).norm() / want.norm()).item()
    out = {"what": f"one UNet3D branch call [1,4,{frames},64,64], SD-1.5 widths + motion modules, ctx [1,77,768], CPU fp32",
           "threads": threads, "reference_seconds": round(t_ref, 2), "port_seconds": round(t_port, 2),
           "port_over_reference": round(t_port / t_ref, 3), "rel_l2_port_vs_reference": rel,
           "reference": "magicanimate/models/unet_controlnet.py UNet3DConditionModel via oracle/ref_shim.py (unsliced attention)",
           "port": "oracle/unet3d_port.py UNet3DOracle(attention_slice_bytes=1<<30)"}
    (ROOT / "profiles" / "r02_cpu_port_vs_reference.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
