"""Generates tests/golden/golden_videomae.npz by running the REAL third-party implementation the reference calls
for its VideoMAE clip route (scripts/extract_features.py:335-403): HF `VideoMAEImageProcessor` + `VideoMAEModel`
on seeded synthetic frames with the synthetic weights of gameplay_vision_llm_b200.weights (seed 2).

    python tests/golden/make_golden_videomae.py          (build container, CPU)
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gameplay_vision_llm_b200 import synth  # noqa: E402
from gameplay_vision_llm_b200.weights import (VideoMAESpec, synth_projector_state_dict,  # noqa: E402
                                                synth_videomae_state_dict)

OUT = os.path.dirname(os.path.abspath(__file__))


def hf_model(spec, sd):
    from transformers import VideoMAEConfig, VideoMAEModel
    cfg = VideoMAEConfig(image_size=spec.image, patch_size=spec.patch, num_frames=spec.frames, tubelet_size=spec.tubelet,
                         hidden_size=spec.hidden, num_hidden_layers=spec.layers, num_attention_heads=spec.heads,
                         intermediate_size=spec.intermediate, hidden_act=spec.act, layer_norm_eps=spec.eps,
                         qkv_bias=True, use_mean_pooling=not spec.final_norm)
    m = VideoMAEModel(cfg).eval()
    m.load_state_dict(sd, strict=True)
    return m.float()


IMAGENET_MEAN, IMAGENET_STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def hf_processor(size, checkpoint_constants=True):
    """checkpoint_constants: the values of MCG-NJU/videomae-base's preprocessor_config.json (ImageNet mean / std,
    resample 2) — what `VideoMAEImageProcessor.from_pretrained` builds in the reference; False = the class defaults
    (0.5 / 0.5), which only apply to a config that carries no values."""
    from transformers import VideoMAEImageProcessor
    kw = {"image_mean": IMAGENET_MEAN, "image_std": IMAGENET_STD} if checkpoint_constants else {}
    return VideoMAEImageProcessor(size={"shortest_edge": size}, crop_size={"height": size, "width": size}, resample=2, **kw)


def main():
    from PIL import Image
    torch.set_num_threads(os.cpu_count())
    out = {}
    # ---- preprocessing: the class-default processor on 1080p frames (224 x 398 resize, crop columns [87, 311)) ----
    frames = np.concatenate([synth.scene_frames_np(0, 2), synth.noise_frames(2, seed=1001).numpy()], 0)
    pv = hf_processor(224, False)([Image.fromarray(f) for f in frames], return_tensors="pt")["pixel_values"][0].numpy()
    u8 = np.rint(pv * 127.5 + 127.5).astype(np.uint8)
    assert np.array_equal(((u8.astype(np.float32) - 127.5) / 127.5).astype(np.float32), pv)
    out["pre1080_sha"] = np.frombuffer(bytes.fromhex(hashlib.sha256(u8.tobytes()).hexdigest()), np.uint8)
    out["pre1080_rows"] = u8[:, :, ::28, :].copy()
    # the same frames with the checkpoint's constants: fp32 pixel_values, full-tensor hash + sampled rows
    pv = hf_processor(224)([Image.fromarray(f) for f in frames], return_tensors="pt")["pixel_values"][0].numpy()
    out["pre1080_imagenet_sha"] = np.frombuffer(bytes.fromhex(hashlib.sha256(np.ascontiguousarray(pv).tobytes()).hexdigest()), np.uint8)
    out["pre1080_imagenet_rows"] = pv[:, :, ::28, :].copy()

    # ---- tiny encoder: every seam from HF, frames small enough for a pure-CPU run ----
    spec = VideoMAESpec.tiny()
    sd = synth_videomae_state_dict(spec, seed=2)
    m = hf_model(spec, sd)
    small = synth.noise_frames(2 * spec.frames, 123, 211, seed=9).numpy()
    pv = hf_processor(spec.image)([Image.fromarray(f) for f in small[:spec.frames]], return_tensors="pt")["pixel_values"]
    pv2 = hf_processor(spec.image)([Image.fromarray(f) for f in small[spec.frames:]], return_tensors="pt")["pixel_values"]
    pv = torch.cat([pv, pv2], 0)  # [2 clips, frames, 3, S, S]
    with torch.no_grad():
        o = m(pixel_values=pv, output_hidden_states=True)
    # the use_mean_pooling=True variant (fine-tuned checkpoints): VideoMAEModel has no final LayerNorm
    import dataclasses
    spec_nf = dataclasses.replace(spec, final_norm=False)
    sd_nf = synth_videomae_state_dict(spec_nf, seed=2)
    assert "layernorm.weight" in sd and "layernorm.weight" not in sd_nf
    with torch.no_grad():
        o_nf = hf_model(spec_nf, sd_nf)(pixel_values=pv)
    out.update(tiny_nofinal_last_hidden_state=o_nf.last_hidden_state.numpy(),
               tiny_nofinal_pooled=o_nf.last_hidden_state.mean(dim=1).numpy())
    out.update(tiny_pixel_values=pv.numpy(), tiny_embeddings=o.hidden_states[0].numpy(),
               tiny_layer0=o.hidden_states[1].numpy(), tiny_last_hidden_state=o.last_hidden_state.numpy(),
               tiny_pooled=o.last_hidden_state.mean(dim=1).numpy())

    # ---- base encoder: one 16-frame 1080p clip (two scenes), pooled + projected ----
    spec = VideoMAESpec.base()
    sd = synth_videomae_state_dict(spec, seed=2)
    m = hf_model(spec, sd)
    clip = np.concatenate([synth.scene_frames_np(0, 8), synth.scene_frames_np(30, 8)], 0)
    pv = hf_processor(224)([Image.fromarray(f) for f in clip], return_tensors="pt")["pixel_values"]
    with torch.no_grad():
        pooled = m(pixel_values=pv).last_hidden_state.mean(dim=1)
    psd = synth_projector_state_dict(spec.hidden, 4096, seed=3)
    proj = torch.nn.Sequential(torch.nn.Linear(spec.hidden, 4096), torch.nn.GELU(), torch.nn.Linear(4096, 4096))
    proj.load_state_dict({k[len("net."):]: v for k, v in psd.items()})
    with torch.no_grad():
        projected = proj(pooled)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    out.update(base_pooled=pooled.numpy(), base_projected=projected.numpy(), base_weight_checksum=np.array([wsum]))
    print("base pooled stats", pooled.mean().item(), pooled.std().item())
    np.savez_compressed(os.path.join(OUT, "golden_videomae.npz"), **out)
    print("wrote golden_videomae.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
