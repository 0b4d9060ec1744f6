"""Golden vectors of the masked-region variant, produced by the REFERENCE'S OWN
`SigLIPSemanticEncoder.encode_masked_regions` (src/perception/siglip_semantic_encoder.py:485-562) with an HF
`SiglipVisionModel` injected (fp32, CPU) whose position table is interpolated for non-square grids
(`interpolate_pos_encoding=True` — the fix the survey names; with the published 729-position checkpoint the reference
raises for every non-square region).  Run in the build container:

    python tests/golden/make_golden_regions.py

Output: tests/golden/golden_regions.npz
  mid_*     a 140-pixel / 10 x 10-position tower (hd 72, 2 layers): 7 regions of a 270 x 480 frame, in batches of 16
            (mixed grids, zero padding) and of 1, for the pooling strategies mean / cls / max
  so_*      the full so400m tower: 3 regions of a 1080p G1 frame (landscape 12 x 27, portrait 27 x 15, square 27 x 27 = the
            untouched position table, the one case the reference encodes as published),
            one per call like scripts/extract_features.py:568, pooling "mean" (the reference default)
Also the resized bytes of each region as Pillow produced them (SHA-256) and the reference's bbox / aspect ratio.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")
from gameplay_vision_llm_b200 import synth  # noqa: E402
from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_ren_projection_state_dict,  # noqa: E402
                                                synth_siglip_state_dict)
from make_golden import hf_model  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

MID_SPEC = SiglipVisionSpec(hidden=216, intermediate=400, layers=2, heads=3, image=140, patch=14)
MID_CFG = dict(base_resolution=140, min_resolution=42, max_resolution=280)
# (x1, y1, x2, y2) rectangles set True in the masks: wide, tall, square-ish, tiny, full-width strip, empty, near-square
MID_RECTS = [(40, 30, 300, 120), (200, 10, 260, 250), (100, 100, 200, 200), (5, 5, 12, 9), (0, 200, 480, 230), None,
             (300, 60, 420, 175)]
SO_RECTS = [(300, 200, 1300, 640), (900, 60, 1380, 960), (400, 300, 920, 820)]


def rect_mask(shape, rect):
    m = np.zeros(shape[:2], np.bool_)
    if rect is not None:
        x1, y1, x2, y2 = rect
        m[y1:y2, x1:x2] = True
    return m


class Shim(torch.nn.Module):
    """What `AutoModel.from_pretrained` would hand the reference, with the position table interpolated."""

    def __init__(self, vm):
        super().__init__()
        self.vm = vm

    @property
    def device(self):
        return torch.device("cpu")

    def get_image_features(self, pixel_values=None, **kw):
        return self.vm(pixel_values=pixel_values, interpolate_pos_encoding=True).pooler_output

    def vision_model(self, pixel_values=None, **kw):
        return self.vm(pixel_values=pixel_values, interpolate_pos_encoding=True)


def reference_encoder(spec, sd, psd, pool, batch_size, cfg_kw):
    from perception.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder
    cfg = NaFlexConfig(device="cpu", dtype=torch.float32, use_amp=False, use_tf32=False, embedding_dim=spec.hidden,
                       pool_strategy=pool, batch_size=batch_size, **cfg_kw)
    enc = SigLIPSemanticEncoder(cfg)
    enc.encoder._model = Shim(hf_model(spec, sd))
    enc.encoder._processor = None
    enc.projection.load_state_dict(psd)
    return enc


def resized_sha(enc, frame, masks):
    out = []
    for _, mask in masks:
        region, _ = enc.region_extractor.extract_masked_region(frame, mask)
        resized, _ = enc.region_extractor.resizer.resize_with_aspect_ratio(region)
        a = np.asarray(resized)
        out.append(np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8))
    return np.stack(out)


def main():
    torch.set_num_threads(os.cpu_count())
    gold = {}
    # ---- mid tower ----
    spec = MID_SPEC
    sd = synth_siglip_state_dict(spec, seed=0)
    psd = synth_ren_projection_state_dict(spec.hidden, seed=3)
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    masks = [(f"e{i}", rect_mask(frame.shape, r)) for i, r in enumerate(MID_RECTS)]
    for pool in ("mean", "cls", "max"):
        for bs in (16, 1):
            enc = reference_encoder(spec, sd, psd, pool, bs, MID_CFG)
            with torch.no_grad():
                res = enc.encode_masked_regions(frame, masks)
            gold[f"mid_{pool}_bs{bs}"] = torch.stack([r.embedding for r in res]).numpy()
    gold["mid_bbox"] = np.array([r.original_bbox for r in res], np.int64)
    gold["mid_aspect"] = np.array([r.aspect_ratio for r in res], np.float64)
    gold["mid_resized_sha"] = resized_sha(enc, frame, masks)
    # encode_with_context (scipy dilation) on one mask
    enc = reference_encoder(spec, sd, psd, "mean", 16, MID_CFG)
    with torch.no_grad():
        a, b = enc.encode_with_context(frame, masks[2][1], context_radius=30)
    gold["mid_context"] = torch.stack([a.embedding, b.embedding]).numpy()
    gold["mid_context_bbox"] = np.array([a.original_bbox, b.original_bbox], np.int64)

    # ---- so400m ----
    spec = SiglipVisionSpec.so400m()
    sd = synth_siglip_state_dict(spec, seed=0)
    psd = synth_ren_projection_state_dict(spec.hidden, seed=3)
    frame = synth.scene_frames_np(40, 1)[0]
    masks = [(f"s{i}", rect_mask(frame.shape, r)) for i, r in enumerate(SO_RECTS)]
    enc = reference_encoder(spec, sd, psd, "mean", 16, {})
    with torch.no_grad():
        res = [enc.encode_masked_regions(frame, [mk])[0] for mk in masks]  # one call per detection (:568)
        res_batched = enc.encode_masked_regions(frame, masks)               # one padded batch of 3
    gold["so_mean_single"] = torch.stack([r.embedding for r in res]).numpy()
    gold["so_mean_batched"] = torch.stack([r.embedding for r in res_batched]).numpy()
    gold["so_bbox"] = np.array([r.original_bbox for r in res], np.int64)
    gold["so_aspect"] = np.array([r.aspect_ratio for r in res], np.float64)
    gold["so_resized_sha"] = resized_sha(enc, frame, masks)
    gold["weight_checksum"] = np.array([float(sum(v.double().sum() for v in sd.values()))])
    np.savez_compressed(os.path.join(OUT, "golden_regions.npz"), **gold)
    for k, v in gold.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    main()
