"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the masked-region variant of the path:
`SigLIPSemanticEncoder.encode_masked_regions` (reference: src/perception/siglip_semantic_encoder.py:485-562),
reached from `scripts/extract_features.py:551-585` when SAM detections exist.  Steps and where their
arithmetic lives:

  * mask -> expanded bounding box -> crop: `RegionExtractor.extract_masked_region` (:301-344), restated literally.
  * NaFlex target size: `AspectPreservingResizer.compute_optimal_size` (:97-135), restated literally.
  * `image.resize((w, h), Image.Resampling.BICUBIC)` (:155): Pillow (12.x here; `uv.lock:2734` pins 12.0.0)
    `src/libImaging/Resample.c`: `precompute_coeffs` (double coefficients, filter support scaled for
    down-sampling, a = -0.5), `normalize_coeffs_8bpc` (fixed point, PRECISION_BITS = 32 - 8 - 2 = 22),
    `ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc` with a uint8 intermediate,
    `clip8((2^21 + sum px * k) >> 22)`.  Restated in numpy integer arithmetic below.
  * `/255`, ImageNet mean / std in fp32 (:357-365), zero padding to the batch maximum (:527-537), cast to the
    model dtype (:539).
  * `SigLIPEncoder.forward` (:246-289): `get_image_features` + `vision_model(...)`.  With the published
    729-position checkpoint this raises for every non-square region (HF adds a [729, D] position table to a
    different token count: HF:models/siglip/modeling_siglip.py:176-186), so `extract_features.py:584` skips
    them; SURVEY.md §8 f.4 asks for the variant "with interpolate_pos_encoding fixed", i.e. HF's own
    `interpolate_pos_encoding=True` (HF :137-174: bicubic, align_corners=False, torch's A = -0.75 cubic
    convolution with border-clamped taps: ATen native/UpSample.h `cubic_convolution1/2`,
    `area_pixel_compute_source_index`).  Square 378 x 378 regions take the untouched table either way.
  * `_pool_features` (:426-443) and the REN `projection` head, fp32 (:416-420, :547).

Pinned by tests/test_region_cpu.py against Pillow live (bit-exact), against `torch.nn.functional.interpolate`
live, and against tests/golden/golden_regions.npz, which tests/golden/make_golden_regions.py produced by running
the reference's OWN `encode_masked_regions` with an HF `SiglipVisionModel` injected (fp32, CPU).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import siglip_ref

PRECISION_BITS = 32 - 8 - 2
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------------- Pillow bicubic
def _bicubic(x: float) -> float:
    """Resample.c `bicubic_filter` (a = -0.5)."""
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int):
    """Resample.c `precompute_coeffs` (box = the whole axis) + `normalize_coeffs_8bpc`.
    Returns (xmin[out], count[out], k[out, ksize] int32)."""
    filterscale = scale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int64)
    count = np.zeros(out_size, np.int64)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = [_bicubic((x + lo - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        xmin[xx], count[xx] = lo, n
    return xmin, count, kk


def _resample_last_axis(img: np.ndarray, out_size: int) -> np.ndarray:
    """img uint8 [..., n, C] resampled along axis -2 (one `ImagingResample*_8bpc` pass)."""
    n = img.shape[-2]
    xmin, count, kk = pil_coeffs(n, out_size)
    acc = np.full(img.shape[:-2] + (out_size, img.shape[-1]), 1 << (PRECISION_BITS - 1), np.int64)
    for t in range(kk.shape[1]):
        live = count > t
        idx = np.where(live, xmin + t, 0)
        acc += img[..., idx, :].astype(np.int64) * np.where(live, kk[:, t], 0).astype(np.int64)[:, None]
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bicubic_u8(img_hwc: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """`PIL.Image.fromarray(img).resize((out_w, out_h), Image.Resampling.BICUBIC)` for an RGB uint8 image:
    horizontal pass first (over the source rows the vertical pass reads; filtering all of them gives the same
    bytes), then the vertical pass (Resample.c `ImagingResampleInner`).  An axis that keeps its size gets the
    identity taps (bicubic(0) = 1, bicubic(+-1) = bicubic(+-2) = 0), which is Pillow's skipped pass."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    tmp = _resample_last_axis(img_hwc, out_w)                                   # [H, out_w, C]
    return _resample_last_axis(tmp.transpose(1, 0, 2), out_h).transpose(1, 0, 2)  # [out_h, out_w, C]


# --------------------------------------------------------------------------------------------- host geometry
def compute_optimal_size(original_h: int, original_w: int, base: int = 384, min_res: int = 128, max_res: int = 768,
                         patch: int = 14) -> tuple[int, int]:
    """`AspectPreservingResizer.compute_optimal_size` (reference :97-135).  Returns (target_h, target_w)."""
    aspect_ratio = original_w / original_h
    if aspect_ratio >= 1:
        target_w = min(max_res, max(min_res, base))
        target_h = int(target_w / aspect_ratio)
        target_h = max(min_res, target_h)
    else:
        target_h = min(max_res, max(min_res, base))
        target_w = int(target_h * aspect_ratio)
        target_w = max(min_res, target_w)
    target_h = (target_h // patch) * patch
    target_w = (target_w // patch) * patch
    return max(patch, target_h), max(patch, target_w)


def extract_bbox(frame_shape, mask: np.ndarray, expand_ratio: float = 0.1) -> tuple[int, int, int, int]:
    """The bounding box of `RegionExtractor.extract_masked_region` (reference :318-338)."""
    ys, xs = np.where(mask)
    if len(xs) == 0:
        h, w = frame_shape[:2]
        cx, cy = w // 2, h // 2
        size = min(h, w) // 4
        return (cx - size, cy - size, cx + size, cy + size)
    x_min, x_max = xs.min(), xs.max()
    y_min, y_max = ys.min(), ys.max()
    width = x_max - x_min
    height = y_max - y_min
    x_min = max(0, int(x_min - width * expand_ratio))
    y_min = max(0, int(y_min - height * expand_ratio))
    x_max = min(frame_shape[1], int(x_max + width * expand_ratio))
    y_max = min(frame_shape[0], int(y_max + height * expand_ratio))
    return (x_min, y_min, x_max, y_max)


def prepare_region_tensor(region_hwc: np.ndarray, preserve_aspect_ratio: bool = True, base: int = 384,
                          min_res: int = 128, max_res: int = 768) -> tuple[torch.Tensor, float]:
    """`RegionExtractor.prepare_region_tensor` (reference :346-367): resize, /255, ImageNet normalisation, CHW fp32."""
    h, w = region_hwc.shape[:2]
    aspect_ratio = w / h
    if preserve_aspect_ratio:
        th, tw = compute_optimal_size(h, w, base, min_res, max_res)
    else:
        th = tw = base
    resized = region_hwc if (tw, th) == (w, h) else pil_resize_bicubic_u8(region_hwc, tw, th)
    t = torch.from_numpy(resized.copy()).float().permute(2, 0, 1) / 255.0
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    return (t - mean) / std, aspect_ratio


# --------------------------------------------------------------------------------------------- position table
def _cubic_coeffs(t: float, A: float = -0.75):
    """ATen UpSample.h `get_cubic_upsample_coefficients`."""
    def c1(x):  # |x| <= 1
        return ((A + 2) * x - (A + 3)) * x * x + 1

    def c2(x):  # 1 < |x| < 2
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def _axis_taps(in_size: int, out_size: int):
    """Per output index: 4 clamped source indices and 4 weights (align_corners=False, scale = in / out)."""
    scale = in_size / out_size
    idx = np.zeros((out_size, 4), np.int64)
    wgt = np.zeros((out_size, 4), np.float64)
    for o in range(out_size):
        src = scale * (o + 0.5) - 0.5
        i0 = math.floor(src)
        wgt[o] = _cubic_coeffs(src - i0)
        idx[o] = [min(max(i0 - 1 + j, 0), in_size - 1) for j in range(4)]
    return idx, wgt


def interpolate_pos(pos: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """HF `SiglipVisionEmbeddings.interpolate_pos_encoding` (HF :137-174): pos [g*g, D] -> [gh*gw, D]."""
    n, D = pos.shape
    g = int(n ** 0.5)
    if gh == g and gw == g:
        return pos
    p = pos.double().numpy().reshape(g, g, D)
    iy, wy = _axis_taps(g, gh)
    ix, wx = _axis_taps(g, gw)
    rows = (p[:, ix, :] * wx[None, :, :, None]).sum(2)            # [g, gw, D]
    out = (rows[iy, :, :] * wy[:, :, None, None]).sum(1)          # [gh, gw, D]
    return torch.from_numpy(out.reshape(gh * gw, D)).to(pos.dtype)


# --------------------------------------------------------------------------------------------- the whole path
def ren_projection(psd: dict, x: torch.Tensor) -> torch.Tensor:
    """The `projection` head (reference :416-420): Linear -> nn.GELU() (erf) -> Linear, fp32; keys "0.*", "2.*"."""
    h = siglip_ref.gelu_erf(x.float() @ psd["0.weight"].float().T + psd["0.bias"].float())
    return h @ psd["2.weight"].float().T + psd["2.bias"].float()


def pool_features(sequence: torch.Tensor, pooled: torch.Tensor, strategy: str) -> torch.Tensor:
    """`_pool_features` (reference :426-443)."""
    if strategy == "mean":
        return sequence.mean(dim=1)
    if strategy == "max":
        return sequence.max(dim=1)[0]
    return pooled


def encode_masked_regions(sd: dict, psd: dict, frame: np.ndarray, masks, heads: int, patch: int = 14, eps: float = 1e-6,
                          pool_strategy: str = "mean", batch_size: int = 16, base: int = 384, min_res: int = 128,
                          max_res: int = 768, preserve_aspect_ratio: bool = True, detail: dict | None = None):
    """Reference :485-562 with `interpolate_pos_encoding=True`.  Returns a list of
    (entity_id, embedding fp32 [D], bbox, aspect_ratio)."""
    prepared, meta = [], []
    for entity_id, mask in masks:
        bbox = extract_bbox(frame.shape, mask)
        x1, y1, x2, y2 = bbox
        t, ar = prepare_region_tensor(frame[y1:y2, x1:x2], preserve_aspect_ratio, base, min_res, max_res)
        prepared.append(t)
        meta.append((entity_id, bbox, ar))
    out = []
    for i in range(0, len(prepared), batch_size):
        chunk = prepared[i:i + batch_size]
        max_h = max(t.shape[1] for t in chunk)
        max_w = max(t.shape[2] for t in chunk)
        batch = torch.stack([torch.nn.functional.pad(t, (0, max_w - t.shape[2], 0, max_h - t.shape[1])) for t in chunk])
        gh, gw = max_h // patch, max_w // patch
        pre = "vision_model." if any(k.startswith("vision_model.") for k in sd) else ""
        pos = interpolate_pos(sd[pre + "embeddings.position_embedding.weight"].float(), gh, gw)
        seams: dict = {}
        pooled = siglip_ref.vision_forward(sd, batch, heads, patch, eps, seams=seams, pos=pos)
        feats = pool_features(seams["last_hidden_state"], pooled, pool_strategy)
        emb = ren_projection(psd, feats)
        if detail is not None:
            detail.setdefault("pixel_values", []).append(batch)
            detail.setdefault("features", []).append(feats)
        for j, e in enumerate(emb):
            out.append((meta[i + j][0], e, meta[i + j][1], meta[i + j][2]))
    return out
