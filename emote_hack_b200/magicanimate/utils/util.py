"""Frames -> file: the step after the hot path, mirroring `magicanimate/utils/util.py`.

`save_videos_grid` (util.py:21-33) builds every frame on the device with one kernel (`emote_video_grid_u8`:
torchvision.make_grid layout + optional (x+1)/2 + the truncating uint8 cast) and copies the uint8 frames to the host once
(the reference moves fp32 frames, loops over them in Python and casts with numpy).  The container behind the writer is
not part of the numerics: imageio when it is importable (what the reference calls), else OpenCV (`.mp4` / `.avi`) or
Pillow (`.gif`).  The latent interpolation helpers (util.py:118-141) are here because `interpolate_latents` uses them.
"""
import os
from typing import List, Optional

import numpy as np
import torch

from ... import ops


def video_frames_u8(videos: torch.Tensor, rescale: bool = False, n_rows: int = 6) -> torch.Tensor:
    """[b, c, t, h, w] float (c = 1 or 3) -> uint8 [t, Hg, Wg, 3] on the device — the frames util.py:22-30 would append."""
    if not videos.is_cuda:
        raise ops._lib.EmoteKernelError("video_frames_u8: expected a CUDA tensor (there is no CPU path)")
    return ops.video_grid_u8(videos.float().contiguous(), nrow=n_rows, padding=2, rescale=rescale)


def _write_frames(path: str, frames: np.ndarray, fps: float) -> str:
    """frames uint8 [t, H, W, 3] RGB -> file; returns the writer used."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    try:
        import imageio
        imageio.mimsave(path, list(frames), fps=fps)
        return "imageio"
    except ImportError:
        pass
    ext = os.path.splitext(path)[1].lower()
    if ext == ".gif":
        from PIL import Image
        imgs = [Image.fromarray(f) for f in frames]
        imgs[0].save(path, save_all=True, append_images=imgs[1:], duration=max(1, int(round(1000.0 / fps))), loop=0)
        return "pillow"
    if ext == ".npy":
        np.save(path, frames)
        return "numpy"
    import cv2
    t, h, w, _ = frames.shape
    fourcc = cv2.VideoWriter_fourcc(*("mp4v" if ext in (".mp4", ".m4v", ".mov") else "MJPG"))
    wr = cv2.VideoWriter(path, fourcc, float(fps), (w, h))
    if not wr.isOpened():
        raise RuntimeError(f"no video writer for {path!r} (imageio is not installed and OpenCV cannot open this container)")
    for f in frames:
        wr.write(np.ascontiguousarray(f[:, :, ::-1]))   # OpenCV takes BGR
    wr.release()
    return "opencv"


def save_videos_grid(videos: torch.Tensor, path: str, rescale: bool = False, n_rows: int = 6, fps: int = 25) -> str:
    """util.py:21-33.  `videos` [b, c, t, h, w] in [0, 1] (or [-1, 1] with rescale) on the device."""
    frames = video_frames_u8(videos, rescale, n_rows)
    host = torch.empty(frames.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(frames)
    return _write_frames(path, host.numpy(), fps)


def save_images_grid(images: torch.Tensor, path: str) -> None:
    """util.py:35-42: [b, c, 1, h, w] -> one PNG/JPEG grid (make_grid's default nrow = 8)."""
    assert images.shape[2] == 1   # no time dimension
    grid = video_frames_u8(images, False, 8)[0].cpu().numpy()
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    from PIL import Image
    Image.fromarray(grid).save(path)


def video2images(path: str, step: int = 4, length: int = 16, start: int = 0) -> List[np.ndarray]:
    """util.py:102-108: every `step`-th frame (RGB uint8 arrays), at most `length` of them."""
    try:
        import imageio
        frames = [np.array(f) for f in imageio.get_reader(path)]
    except ImportError:
        if os.path.splitext(path)[1].lower() == ".npy":
            frames = list(np.load(path))
        else:
            import cv2
            cap, frames = cv2.VideoCapture(path), []
            while True:
                ok, f = cap.read()
                if not ok:
                    break
                frames.append(np.ascontiguousarray(f[:, :, ::-1]))
            cap.release()
    return frames[start::step][:length]


def images2video(video, path: str, fps: int = 8) -> None:
    """util.py:111-113."""
    _write_frames(path, np.stack([np.asarray(f, dtype=np.uint8) for f in video]), fps)


# ---- latent interpolation (util.py:116-141)
def linear(v1: torch.Tensor, v2: torch.Tensor, t: float) -> torch.Tensor:
    return (1.0 - t) * v1 + t * v2


def slerp(v0: torch.Tensor, v1: torch.Tensor, t: float, DOT_THRESHOLD: float = 0.9995) -> torch.Tensor:
    dot = ((v0 / v0.norm()) * (v1 / v1.norm())).sum()
    if dot.abs() > DOT_THRESHOLD:
        return linear(v0, v1, t)
    omega = dot.acos()
    return (((1.0 - t) * omega).sin() * v0 + (t * omega).sin() * v1) / omega.sin()


tensor_interpolation: Optional[object] = None


def get_tensor_interpolation_method():
    return tensor_interpolation


def set_tensor_interpolation_method(is_slerp: bool) -> None:
    global tensor_interpolation
    tensor_interpolation = slerp if is_slerp else linear
