"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of the reference's VideoMAE clip route (scripts/extract_features.py:335-403):
`VideoMAEImageProcessor(pil_frames)` -> `VideoMAEModel(**inputs).last_hidden_state.mean(dim=1)`.
The arithmetic lives in HuggingFace transformers (5.5.0 installed here and restated):
  models/videomae/image_processing_videomae.py:35-105  shortest-edge resize (uint8, antialias) -> center crop ->
                                                        fused rescale + normalize (class default mean = std = 0.5;
                                                        MCG-NJU/videomae-base's preprocessor_config.json: ImageNet)
  models/videomae/modeling_videomae.py:80-91            fixed sinusoid position table
  :120-178 tubelet Conv3d patch embedding   :209-268 self-attention (q_bias / v_bias, key without bias)
  :335-385 pre-LN encoder layer (exact-erf GELU, eps 1e-12)   :407-475 optional final LayerNorm.
Pinned against the HF modules by tests/test_oracle_cpu.py and tests/golden/golden_videomae.npz.
"""
from __future__ import annotations

import numpy as np
import torch

from . import preprocess_ref
from .siglip_ref import _ln, _mha, gelu_erf


def resize_geometry(H: int, W: int, shortest_edge: int = 224, crop: int = 224):
    """(out_h, out_w, crop_y0, crop_x0) of the processor for an H x W frame: the short side becomes
    `shortest_edge`, the long side int(short * long / short_in) (HF image_transforms.get_resize_output_image_size),
    then a centered crop x crop window (torchvision center_crop: int(round((size - crop) / 2)))."""
    if H <= W:
        out_h, out_w = shortest_edge, int(shortest_edge * W / H)
    else:
        out_h, out_w = int(shortest_edge * H / W), shortest_edge
    return out_h, out_w, int(round((out_h - crop) / 2.0)), int(round((out_w - crop) / 2.0))


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def pixel_values(frames_hwc: np.ndarray, shortest_edge: int = 224, crop: int = 224, resample: int = 2,
                 image_mean=IMAGENET_MEAN, image_std=IMAGENET_STD) -> np.ndarray:
    """uint8 [N,H,W,3] -> float32 [N,3,crop,crop] == VideoMAEImageProcessor(frames).pixel_values[0] with the
    constants of MCG-NJU/videomae-base's preprocessor_config.json (pass 0.5 / 0.5 for the class defaults)."""
    _, H, W, _ = frames_hwc.shape
    out_h, out_w, y0, x0 = resize_geometry(H, W, shortest_edge, crop)
    u8 = preprocess_ref.resize_u8(frames_hwc, out_h, out_w, resample)[:, :, y0:y0 + crop, x0:x0 + crop]
    sub, div = preprocess_ref.fused_sub_div(image_mean, image_std)
    return ((u8.astype(np.float32) - sub[None, :, None, None]) / div[None, :, None, None]).astype(np.float32)


def tubelet_patches(pv: torch.Tensor, frames: int, patch: int, tubelet: int) -> torch.Tensor:
    """[clips*frames,3,H,W] -> [clips, T, 3*tubelet*p*p]: im2col of Conv3d(kernel = stride = (tubelet,p,p)) after the
    permute to (B,C,T,H,W); token order (t/tubelet, y/p, x/p), column order (c, kt, ky, kx)."""
    n, C, H, W = pv.shape
    clips, gh, gw, gt = n // frames, H // patch, W // patch, frames // tubelet
    x = pv.reshape(clips, gt, tubelet, C, gh, patch, gw, patch)
    x = x.permute(0, 1, 4, 6, 3, 2, 5, 7)  # clip, tt, py, px, c, kt, ky, kx
    return x.reshape(clips, gt * gh * gw, C * tubelet * patch * patch)


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.from_numpy(table.astype(np.float32))


def encoder_forward(sd: dict, pv: torch.Tensor, frames: int, heads: int, patch: int, tubelet: int,
                    eps: float = 1e-12, seams: dict | None = None, dtype=torch.float32) -> torch.Tensor:
    """pixel_values [clips*frames,3,S,S] -> mean-pooled clip embedding [clips, D]."""
    pre = "videomae." if any(k.startswith("videomae.") for k in sd) else ""
    W = lambda n: sd[pre + n].to(dtype)  # noqa: E731
    cols = tubelet_patches(pv.to(dtype), frames, patch, tubelet)
    wp = W("embeddings.patch_embeddings.projection.weight")
    D = wp.shape[0]
    h = cols @ wp.reshape(D, -1).T + W("embeddings.patch_embeddings.projection.bias")
    h = h + sinusoid_table(cols.shape[1], D).to(dtype)[None]
    if seams is not None:
        seams["patches"] = cols
        seams["embeddings"] = h
    n_layers = 1 + max(int(k[len(pre):].split(".")[2]) for k in sd if k.startswith(pre + "encoder.layer."))
    zero = torch.zeros(D, dtype=dtype)
    for i in range(n_layers):
        p = f"encoder.layer.{i}."
        a = p + "attention.attention."
        qb = W(a + "q_bias") if (pre + a + "q_bias") in sd else zero
        vb = W(a + "v_bias") if (pre + a + "v_bias") in sd else zero
        y = _ln(h, W(p + "layernorm_before.weight"), W(p + "layernorm_before.bias"), eps)
        att = _mha(y, W(a + "query.weight"), qb, W(a + "key.weight"), zero, W(a + "value.weight"), vb, heads)
        h = h + (att @ W(p + "attention.output.dense.weight").T + W(p + "attention.output.dense.bias"))
        y = _ln(h, W(p + "layernorm_after.weight"), W(p + "layernorm_after.bias"), eps)
        y = gelu_erf(y @ W(p + "intermediate.dense.weight").T + W(p + "intermediate.dense.bias"))
        h = h + (y @ W(p + "output.dense.weight").T + W(p + "output.dense.bias"))
        if seams is not None:
            seams[f"layer{i}"] = h
    if (pre + "layernorm.weight") in sd:
        h = _ln(h, W("layernorm.weight"), W("layernorm.bias"), eps)
    if seams is not None:
        seams["last_hidden_state"] = h
    return h.mean(dim=1)
