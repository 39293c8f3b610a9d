"""TEST INFRASTRUCTURE ONLY — CPU restatement of the frame loop of `save_videos_grid`
(/root/reference/magicanimate/utils/util.py:21-30) and of the latent interpolation helpers (:125-141).

Parity pinned: tests/golden/video_grid.pt holds the frames the reference function itself handed to `imageio.mimsave`
(oracle/make_golden.py `util` runs the untouched reference file with a recording stand-in for the absent `imageio`),
and the outputs of its `linear` / `slerp`.

numpy only.  `torchvision.utils.make_grid` (third-party, torchvision 0.26 in this image; unpinned in the reference's
requirements) is restated from its documented behaviour: nrow images per row, `padding` zero pixels between and around
the cells, a single image returned unchanged, one channel replicated to three.
"""
from __future__ import annotations

import math

import numpy as np


def make_grid(images: np.ndarray, nrow: int = 8, padding: int = 2) -> np.ndarray:
    """[n, c, h, w] -> [3|c, Hg, Wg] (torchvision.utils.make_grid with pad_value 0, normalize False)"""
    if images.shape[1] == 1:
        images = np.concatenate([images] * 3, axis=1)
    n, c, h, w = images.shape
    if n == 1:
        return images[0]
    xmaps = min(nrow, n)
    ymaps = int(math.ceil(n / xmaps))
    ch, cw = h + padding, w + padding
    grid = np.zeros((c, ch * ymaps + padding, cw * xmaps + padding), dtype=images.dtype)
    for k in range(n):
        y, x = divmod(k, xmaps)
        grid[:, y * ch + padding: y * ch + padding + h, x * cw + padding: x * cw + padding + w] = images[k]
    return grid


def video_frames_u8(videos: np.ndarray, rescale: bool = False, n_rows: int = 6) -> np.ndarray:
    """util.py:22-30: [b, c, t, h, w] fp32 -> uint8 [t, Hg, Wg, 3]"""
    frames = []
    for f in range(videos.shape[2]):
        x = make_grid(videos[:, :, f], nrow=n_rows).transpose(1, 2, 0)
        if rescale:
            x = (x + np.float32(1.0)) / np.float32(2.0)
        frames.append((x * np.float32(255)).astype(np.int32).astype(np.uint8))   # truncate, keep the low 8 bits
    return np.stack(frames)


def linear(v0: np.ndarray, v1: np.ndarray, t: float) -> np.ndarray:
    """util.py:125-126"""
    return (1.0 - t) * v0 + t * v1


def slerp(v0: np.ndarray, v1: np.ndarray, t: float, dot_threshold: float = 0.9995) -> np.ndarray:
    """util.py:128-141: spherical interpolation of two tensors treated as flat vectors; linear when nearly parallel"""
    dot = float(((v0 / np.linalg.norm(v0)) * (v1 / np.linalg.norm(v1))).sum())
    if abs(dot) > dot_threshold:
        return linear(v0, v1, t)
    omega = math.acos(dot)
    return (math.sin((1.0 - t) * omega) * v0 + math.sin(t * omega) * v1) / math.sin(omega)
