"""Seeded synthetic inputs for the perception-embedding path.

There is no network, so benchmarks and parity tests run on synthetic 1080p frames:

* ``noise_frames``  (G0) iid uniform uint8 — worst case for the resize filter, used for kernel parity.
* ``scene_frames``  (G1) integer-only procedural "gameplay scenes" — frame ``i`` belongs to scene
  ``i // frames_per_scene``; every pixel is a pure integer function of (i, y, x, c), so the numpy
  (host) and torch (device) generators produce identical bytes on any rank.

Both return ``uint8 [N, H, W, 3]`` in the HWC RGB layout decoded video frames have in the reference
(scripts/extract_features.py:230-264 yields PIL RGB images).
"""
from __future__ import annotations

import numpy as np
import torch

_M32 = 0xFFFFFFFF


def noise_frames(n: int, h: int = 1080, w: int = 1920, seed: int = 1000) -> torch.Tensor:
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, generator=g)


def _scene_constants(scene: int):
    rng = np.random.default_rng(2000 + int(scene))
    base = rng.integers(32, 161, size=3)
    gx = rng.integers(-64, 65, size=3)
    gy = rng.integers(-64, 65, size=3)
    return base.astype(np.int64), gx.astype(np.int64), gy.astype(np.int64)


def _mix32_np(v: np.ndarray) -> np.ndarray:
    v = v & _M32
    v ^= v >> 16
    v = (v * 0x85EBCA6B) & _M32
    v ^= v >> 13
    v = (v * 0xC2B2AE35) & _M32
    v ^= v >> 16
    return v


def scene_frames_np(start: int, n: int, h: int = 1080, w: int = 1920, frames_per_scene: int = 30) -> np.ndarray:
    """Frames ``start .. start+n-1`` of the G1 stream, numpy/host."""
    out = np.empty((n, h, w, 3), np.uint8)
    y = np.arange(h, dtype=np.int64)[:, None, None]
    x = np.arange(w, dtype=np.int64)[None, :, None]
    c = np.arange(3, dtype=np.int64)[None, None, :]
    for k in range(n):
        i = start + k
        s = i // frames_per_scene
        base, gx, gy = _scene_constants(s)
        grad = (x * gx[None, None, :] + y * gy[None, None, :]) >> 10
        shift = 4 * (i % frames_per_scene)  # the checkerboard scrolls, a sprite moves: frames of a scene differ
        checker = 96 * (((((x + shift) >> 6) + (y >> 6) + s) & 1))
        sx, sy = (i * 137) % max(w - 256, 1), (i * 89) % max(h - 256, 1)
        sprite = 80 * ((x >= sx) & (x < sx + 256) & (y >= sy) & (y < sy + 256)) * (1 - 2 * (c == (i % 3)))
        noise = (_mix32_np(((i * h + y) * w + x) * 3 + c) & 31) - 16
        out[k] = np.clip(base[None, None, :] + grad + checker + sprite + noise, 0, 255).astype(np.uint8)
    return out


def _mix32_t(v: torch.Tensor) -> torch.Tensor:
    v = v & _M32
    v = v ^ (v >> 16)
    v = (v * 0x85EBCA6B) & _M32
    v = v ^ (v >> 13)
    v = (v * 0xC2B2AE35) & _M32
    v = v ^ (v >> 16)
    return v


def scene_frames(start: int, n: int, h: int = 1080, w: int = 1920, frames_per_scene: int = 30,
                 device: str | torch.device = "cpu") -> torch.Tensor:
    """Same stream as :func:`scene_frames_np`, generated with torch integer ops on ``device``."""
    dev = torch.device(device)
    y = torch.arange(h, dtype=torch.int64, device=dev)[:, None, None]
    x = torch.arange(w, dtype=torch.int64, device=dev)[None, :, None]
    c = torch.arange(3, dtype=torch.int64, device=dev)[None, None, :]
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    for k in range(n):
        i = start + k
        s = i // frames_per_scene
        base, gx, gy = _scene_constants(s)
        base_t = torch.as_tensor(base, device=dev)[None, None, :]
        gx_t = torch.as_tensor(gx, device=dev)[None, None, :]
        gy_t = torch.as_tensor(gy, device=dev)[None, None, :]
        grad = (x * gx_t + y * gy_t) >> 10  # arithmetic shift == floor, like numpy
        shift = 4 * (i % frames_per_scene)
        checker = 96 * ((((x + shift) >> 6) + (y >> 6) + s) & 1)
        sx, sy = (i * 137) % max(w - 256, 1), (i * 89) % max(h - 256, 1)
        inside = ((x >= sx) & (x < sx + 256) & (y >= sy) & (y < sy + 256)).to(torch.int64)
        sprite = 80 * inside * (1 - 2 * (c == (i % 3)).to(torch.int64))
        noise = (_mix32_t(((i * h + y) * w + x) * 3 + c) & 31) - 16
        out[k] = (base_t + grad + checker + sprite + noise).clamp_(0, 255).to(torch.uint8)
    return out


def timestamps(n: int, fps: float = 1.0, start: int = 0) -> np.ndarray:
    """float64 timestamps ``t_i = i / fps`` (reference: timestamp = frame_idx / video_fps,
    scripts/extract_features.py:252-258)."""
    return (np.arange(start, start + n, dtype=np.float64)) / float(fps)


def region_boxes(n: int, H: int = 1080, W: int = 1920, seed: int = 4000) -> list[tuple[int, int, int, int]]:
    """Seeded detection boxes (x1, y1, x2, y2) of a frame: UI-like wide strips, tall panels and near-square sprites from
    a small set of shapes, so several detections share a target size — the synthetic stand-in for SAM detections
    (scripts/extract_features.py:552-583 encodes one per call)."""
    rng = np.random.default_rng(seed)
    shapes = [(1000, 440), (480, 900), (520, 520), (800, 200), (300, 300), (1200, 600), (240, 640), (640, 360)]
    boxes = []
    for i in range(n):
        w, h = shapes[i % len(shapes)]
        x1, y1 = int(rng.integers(0, W - w)), int(rng.integers(0, H - h))
        boxes.append((x1, y1, x1 + w, y1 + h))
    return boxes
