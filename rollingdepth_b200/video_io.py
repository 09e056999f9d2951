"""Frame loading for `RollingDepthPipeline.__call__` (host-side adjacent code).

Semantics of the reference's rollingdepth/video_io.py:38-137 -- decode, per-frame float32, `resize_max_res`
(aspect-preserving, antialiased) and (x / 255) * 2 - 1 -- with PyAV when it is installed and OpenCV otherwise
(PyAV is not part of this image).  `forward()` takes tensors and never touches this module.
"""
from __future__ import annotations

from os import PathLike
from typing import List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def _interp_mode(resample_method: str) -> str:
    return {"BILINEAR": "bilinear", "BICUBIC": "bicubic", "NEAREST": "nearest", "NEAREST_EXACT": "nearest-exact"}[
        resample_method.upper()]


def resize_frames(img: torch.Tensor, size: List[int], resample_method: str = "BILINEAR") -> torch.Tensor:
    """torchvision.transforms.functional.resize(img, size, interpolation, antialias=True) on tensors."""
    mode = _interp_mode(resample_method)
    if mode.startswith("nearest"):
        return F.interpolate(img, size=tuple(size), mode=mode)
    if img.device.type == "cpu" and img.dtype in (torch.float16, torch.bfloat16):
        # torch has no 16-bit antialiased resampling kernel on CPU (the fp16 presets return fp16 host tensors):
        # resample in fp32 and round once
        return F.interpolate(img.float(), size=tuple(size), mode=mode, align_corners=False, antialias=True).to(img.dtype)
    return F.interpolate(img, size=tuple(size), mode=mode, align_corners=False, antialias=True)


def resize_max_res(img: torch.Tensor, max_edge_resolution: int, resample_method: str = "BILINEAR") -> torch.Tensor:
    """video_io.py:38-68."""
    assert 4 == img.dim(), f"Invalid input shape {img.shape}"
    oh, ow = img.shape[-2:]
    f = min(max_edge_resolution / ow, max_edge_resolution / oh)
    return resize_frames(img, [int(oh * f), int(ow * f)], resample_method)


def _iter_frames_rgb(input_path):
    try:
        import av  # type: ignore
        av_open = av.open
    except (ImportError, AttributeError):
        av_open = None
    if av_open is not None:
        container = av_open(input_path)
        stream = container.streams.video[0]
        stream.thread_type = "AUTO"
        try:
            for frame in container.decode(stream):
                yield frame.to_ndarray(format="rgb24")
        finally:
            container.close()
        return
    import cv2  # OpenCV fallback decoder (same frames, BGR -> RGB)
    cap = cv2.VideoCapture(str(input_path))
    if not cap.isOpened():
        raise RuntimeError(f"cannot open {input_path}")
    try:
        while True:
            ok, bgr = cap.read()
            if not ok:
                break
            yield cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    finally:
        cap.release()


def load_video_frames(input_path: PathLike, start_frame: int = 0, frame_count: int = 0, processing_res: int = 0,
                      resample_method: str = "BILINEAR", verbose: bool = False) -> Tuple[torch.Tensor, torch.Size]:
    """video_io.py:71-137 -> ([N, 3, H, W] float32 in [-1, 1], original (H, W))."""
    assert start_frame >= 0
    end_before = start_frame + frame_count if frame_count > 0 else np.inf
    frames, original_res = [], None
    for i, arr in enumerate(_iter_frames_rgb(input_path)):
        if i >= end_before:
            break
        if i < start_frame:
            continue
        frame = torch.from_numpy(np.ascontiguousarray(arr)).float()
        if original_res is None:
            original_res = frame.shape[:2]
        frame = frame.permute(2, 0, 1)[None]
        if processing_res > 0:
            frame = resize_max_res(frame, processing_res, resample_method)
        frames.append((frame / 255.0) * 2.0 - 1.0)
    if not frames:
        raise RuntimeError(f"No frame is loaded from {input_path}")
    return torch.cat(frames, dim=0), original_res
