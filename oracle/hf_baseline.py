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
