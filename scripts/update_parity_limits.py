#!/usr/bin/env python
"""Rewrites tests/parity_limits.json from the parity logs of `pytest -m gpu` runs on a B200:

    EMOTE_PARITY_LOG=gpurun_out/parity_fp16.log python -m pytest tests -m gpu -q
    EMOTE_OPERAND=bf16 EMOTE_PARITY_LOG=gpurun_out/parity_bf16.log python -m pytest tests -m gpu -q
    python scripts/update_parity_limits.py gpurun_out/parity_fp16.log gpurun_out/parity_bf16.log

limit = 1.5 x the largest error measured for the key (floored at 2e-6: fp32 round-off level checks), per operand type.
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
FACTOR, FLOOR = 1.5, 2e-6


def main(paths):
    out_path = ROOT / "tests" / "parity_limits.json"
    table = json.loads(out_path.read_text()) if out_path.exists() else {}
    seen = {}
    for p in paths:
        for line in Path(p).read_text().splitlines():
            f = line.split()
            if len(f) != 4:
                continue
            op, key, err = f[0], f[1], float(f[2])
            if key.startswith("info."):
                continue
            seen.setdefault(op, {})
            seen[op][key] = max(seen[op].get(key, 0.0), err)
    for op, keys in seen.items():
        table.setdefault(op, {})
        for key, err in keys.items():
            table[op][key] = {"measured": float(f"{err:.4g}"), "limit": float(f"{max(FACTOR * err, FLOOR):.4g}")}
    out_path.write_text(json.dumps(table, indent=1, sort_keys=True) + "\n")
    print(f"{out_path}: " + ", ".join(f"{op}: {len(v)} keys" for op, v in table.items()))


if __name__ == "__main__":
    main(sys.argv[1:])
