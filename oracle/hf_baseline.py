"""ORACLE / CPU BASELINE (test + bench infrastructure only — never on the product path).

Times the reference's CPU route for the metric (BASELINE.json configs[0]):
`SiglipImageProcessor(384x384, mean=std=0.5)` -> `SiglipVisionModel(so400m).pooler_output` ->
`MultiModalProjector(1152, 4096)`, fp32, batch 8, `torch.inference_mode()`, all host threads.

kind = "reference": the third-party modules the reference itself calls (HF transformers + torch.nn;
    src/perception/siglip_semantic_encoder.py:195-204,474-481, src/agent_core/qwen_reasoning_core.py:1009-1013)
    are importable and are what is timed.  (The reference's own thin wrappers are Python files under
    /root/reference, which does not exist on the GPU box; they add no arithmetic.)
kind = "port": transformers is missing -> the oracle's restatement (oracle/preprocess_ref.py +
    oracle/siglip_ref.py) is timed instead.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch


def run(n_frames: int = 8, batch: int = 8, warmup_batches: int = 1, frame_hw=(1080, 1920)) -> dict:
    from gameplay_vision_llm_b200 import synth
    from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_projector_state_dict,
                                                    synth_siglip_state_dict)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = SiglipVisionSpec.so400m()
    sd = synth_siglip_state_dict(spec, seed=0)
    psd = synth_projector_state_dict(spec.hidden, 4096, seed=1)
    H, W = frame_hw
    n_total = n_frames + warmup_batches * batch
    frames = synth.scene_frames_np(0, n_total, H, W)
    kind = "reference"
    try:
        from PIL import Image
        from transformers import SiglipImageProcessor, SiglipVisionConfig, SiglipVisionModel
        cfg = SiglipVisionConfig(hidden_size=spec.hidden, intermediate_size=spec.intermediate,
                                 num_hidden_layers=spec.layers, num_attention_heads=spec.heads, image_size=spec.image,
                                 patch_size=spec.patch, layer_norm_eps=spec.eps, hidden_act=spec.act)
        model = SiglipVisionModel(cfg).eval().float()
        model.load_state_dict(sd, strict=True)
        proc = SiglipImageProcessor(size={"height": spec.image, "width": spec.image}, resample=2, image_mean=[0.5] * 3,
                                    image_std=[0.5] * 3)
        projector = torch.nn.Sequential(torch.nn.Linear(spec.hidden, 4096), torch.nn.GELU(), torch.nn.Linear(4096, 4096))
        projector.load_state_dict({k[len("net."):]: v for k, v in psd.items()})

        def step(fr):
            pv = proc(images=[Image.fromarray(f) for f in fr], return_tensors="pt")["pixel_values"]
            return projector(model(pixel_values=pv).pooler_output)
    except ImportError:
        kind = "port"
        from oracle import preprocess_ref, siglip_ref

        def step(fr):
            pv = torch.from_numpy(preprocess_ref.pixel_values(fr, spec.image, spec.image, 2))
            return siglip_ref.projector_forward(psd, siglip_ref.vision_forward(sd, pv, spec.heads, spec.patch, spec.eps))

    with torch.inference_mode():
        for i in range(warmup_batches):
            step(frames[i * batch:(i + 1) * batch])
        off = warmup_batches * batch
        t0 = time.perf_counter()
        done = 0
        while done < n_frames:
            n = min(batch, n_frames - done)
            out = step(frames[off + done:off + done + n])
            done += n
        dt = time.perf_counter() - t0
    assert out.shape[-1] == 4096
    return {"frames_per_s": n_frames / dt, "seconds": dt, "cores": cores, "kind": kind,
            "sample": f"{n_frames} synthetic 1080p frames in batches of {batch} after {warmup_batches} warm-up batch(es); "
                      f"HF SiglipImageProcessor + SiglipVisionModel(so400m, random init) + projector, fp32, "
                      f"torch {torch.__version__}, {cores} threads"}


def run_regions(n_regions: int = 4, warmup_regions: int = 1, frame_hw=(1080, 1920)) -> dict:
    """The reference's masked-region route on the host cores, one detection per call: bbox mask -> expanded box -> PIL
    bicubic resize -> ImageNet normalisation -> HF SiglipVisionModel (fp32, `interpolate_pos_encoding=True`) ->
    mean over the tokens -> REN projection (src/perception/siglip_semantic_encoder.py:485-562)."""
    from PIL import Image

    from gameplay_vision_llm_b200 import synth
    from gameplay_vision_llm_b200.synth import region_boxes
    from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_ren_projection_state_dict,
                                                    synth_siglip_state_dict)
    from oracle import region_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = SiglipVisionSpec.so400m()
    sd = synth_siglip_state_dict(spec, seed=0)
    psd = synth_ren_projection_state_dict(spec.hidden, seed=3)
    H, W = frame_hw
    frame = synth.scene_frames_np(40, 1, H, W)[0]
    kind = "reference"
    try:
        from transformers import SiglipVisionConfig, SiglipVisionModel
        cfg = SiglipVisionConfig(hidden_size=spec.hidden, intermediate_size=spec.intermediate,
                                 num_hidden_layers=spec.layers, num_attention_heads=spec.heads, image_size=spec.image,
                                 patch_size=spec.patch, layer_norm_eps=spec.eps, hidden_act=spec.act)
        model = SiglipVisionModel(cfg).eval().float()
        model.load_state_dict(sd, strict=True)
        projection = torch.nn.Sequential(torch.nn.Linear(spec.hidden, spec.hidden), torch.nn.GELU(),
                                         torch.nn.Linear(spec.hidden, spec.hidden))
        projection.load_state_dict(psd)

        def one(box):
            mask = np.zeros((H, W), np.bool_)
            mask[box[1]:box[3], box[0]:box[2]] = True
            x1, y1, x2, y2 = region_ref.extract_bbox(frame.shape, mask)
            region = Image.fromarray(frame[y1:y2, x1:x2])
            th, tw = region_ref.compute_optimal_size(region.size[1], region.size[0])
            t = torch.from_numpy(np.array(region.resize((tw, th), Image.Resampling.BICUBIC))).float().permute(2, 0, 1) / 255.0
            t = (t - torch.tensor(region_ref.IMAGENET_MEAN).view(3, 1, 1)) / torch.tensor(region_ref.IMAGENET_STD).view(3, 1, 1)
            seq = model(pixel_values=t[None], interpolate_pos_encoding=True).last_hidden_state
            return projection(seq.mean(dim=1))
    except ImportError:
        kind = "port"

        def one(box):
            mask = np.zeros((H, W), np.bool_)
            mask[box[1]:box[3], box[0]:box[2]] = True
            return region_ref.encode_masked_regions(sd, psd, frame, [("r", mask)], spec.heads, spec.patch, spec.eps)[0][1][None]

    boxes = region_boxes(n_regions + warmup_regions, H, W)
    with torch.inference_mode():
        for b in boxes[:warmup_regions]:
            one(b)
        t0 = time.perf_counter()
        for b in boxes[warmup_regions:]:
            out = one(b)
        dt = time.perf_counter() - t0
    assert out.shape[-1] == spec.hidden
    return {"regions_per_s": n_regions / dt, "seconds": dt, "cores": cores, "kind": kind,
            "sample": f"{n_regions} seeded detections of one synthetic 1080p frame, one call each, after {warmup_regions} "
                      f"warm-up; PIL bicubic + HF SiglipVisionModel(so400m, random init, interpolate_pos_encoding) + mean "
                      f"pool + REN projection, fp32, torch {torch.__version__}, {cores} threads"}
