"""Generates the golden vectors under tests/golden/ by running the REAL third-party implementations the
reference calls (HF transformers SiglipVisionModel / SiglipImageProcessor, torchvision resize,
torch.nn modules) on seeded inputs.  Run in the build container (CPU):

    python tests/golden/make_golden.py

Outputs (small, committed):
  golden_preprocess.npz   torchvision/HF uint8 resize + pixel_values samples for seeded frames
  golden_tiny.npz         HF SiglipVisionModel (tiny spec) seams + pooled + projector on seeded frames
  golden_so400m.npz       HF SiglipVisionModel at the full so400m spec: pooled / projected embeddings of
                          4 G1 frames + top-k over them, with the synthetic weights of
                          gameplay_vision_llm_b200.weights (seed 0 / 1)
Also drives the reference's own classes (injected model) to pin the call sequence.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gameplay_vision_llm_b200 import synth  # noqa: E402
from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_projector_state_dict,  # noqa: E402
                                                synth_siglip_state_dict)

OUT = os.path.dirname(os.path.abspath(__file__))


def hf_model(spec, sd):
    from transformers import SiglipVisionConfig, SiglipVisionModel
    cfg = SiglipVisionConfig(hidden_size=spec.hidden, intermediate_size=spec.intermediate,
                             num_hidden_layers=spec.layers, num_attention_heads=spec.heads, image_size=spec.image,
                             patch_size=spec.patch, layer_norm_eps=spec.eps, hidden_act=spec.act)
    m = SiglipVisionModel(cfg).eval()
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.float()


def hf_processor(size, resample):
    from transformers import SiglipImageProcessor
    return SiglipImageProcessor(size={"height": size, "width": size}, resample=resample, image_mean=[0.5] * 3,
                                image_std=[0.5] * 3)


def ref_projector(enc, psd):
    """The reference's own MultiModalProjector class (src/agent_core/qwen_reasoning_core.py:995)."""
    try:
        sys.path.insert(0, "/root/reference/src")
        from agent_core.qwen_reasoning_core import MultiModalProjector
    except Exception:
        class MultiModalProjector(torch.nn.Module):  # same definition, used only if the reference is absent
            def __init__(self, e, l=4096):
                super().__init__()
                self.net = torch.nn.Sequential(torch.nn.Linear(e, l), torch.nn.GELU(), torch.nn.Linear(l, l))

            def forward(self, x):
                return self.net(x)
    p = MultiModalProjector(enc, psd["net.0.weight"].shape[0])
    p.load_state_dict(psd)
    return p.float().eval()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from PIL import Image
    torch.set_num_threads(os.cpu_count())
    # ---- preprocessing goldens: torchvision path through the HF processor ----
    pre = {}
    for name, frames in (("noise", synth.noise_frames(2, seed=1000).numpy()), ("scene", synth.scene_frames_np(0, 2))):
        for rs in (2, 3):
            pv = hf_processor(384, rs)(images=[Image.fromarray(f) for f in frames], return_tensors="pt")["pixel_values"].numpy()
            u8 = np.rint(pv * 127.5 + 127.5).astype(np.uint8)
            assert np.array_equal(((u8.astype(np.float32) - 127.5) / 127.5).astype(np.float32), pv)
            pre[f"{name}_rs{rs}_sha"] = np.frombuffer(bytes.fromhex(sha(u8)), np.uint8)
            pre[f"{name}_rs{rs}_rows"] = u8[:, :, ::48, :].copy()  # 8 of 384 rows per channel, all columns
    small = synth.noise_frames(3, 123, 211, seed=7).numpy()
    pv = hf_processor(56, 2)(images=[Image.fromarray(f) for f in small], return_tensors="pt")["pixel_values"].numpy()
    pre["small_123x211_to56_rs2"] = np.rint(pv * 127.5 + 127.5).astype(np.uint8)
    np.savez_compressed(os.path.join(OUT, "golden_preprocess.npz"), **pre)

    # ---- tiny tower: every seam from HF ----
    spec = SiglipVisionSpec.tiny()
    sd = synth_siglip_state_dict(spec, seed=0)
    m = hf_model(spec, sd)
    frames = synth.noise_frames(3, 123, 211, seed=7).numpy()
    pv = hf_processor(spec.image, 2)(images=[Image.fromarray(f) for f in frames], return_tensors="pt")["pixel_values"]
    with torch.no_grad():
        o = m(pixel_values=pv, output_hidden_states=True)
    psd = synth_projector_state_dict(spec.hidden, 256, seed=1)
    with torch.no_grad():
        proj = ref_projector(spec.hidden, psd)(o.pooler_output)
    np.savez_compressed(
        os.path.join(OUT, "golden_tiny.npz"), pixel_values=pv.numpy(), embeddings=o.hidden_states[0].numpy(),
        layer0=o.hidden_states[1].numpy(), layer1=o.hidden_states[2].numpy(),
        last_hidden_state=o.last_hidden_state.numpy(), pooled=o.pooler_output.numpy(), projected=proj.numpy())

    # ---- full so400m: pooled + projected for 4 scene frames (two scenes) ----
    spec = SiglipVisionSpec.so400m()
    sd = synth_siglip_state_dict(spec, seed=0)
    m = hf_model(spec, sd)
    frames = np.concatenate([synth.scene_frames_np(0, 2), synth.scene_frames_np(30, 2)], 0)
    pv = hf_processor(384, 2)(images=[Image.fromarray(f) for f in frames], return_tensors="pt")["pixel_values"]
    with torch.no_grad():
        o = m(pixel_values=pv)
    psd = synth_projector_state_dict(spec.hidden, 4096, seed=1)
    with torch.no_grad():
        proj = ref_projector(spec.hidden, psd)(o.pooler_output)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    np.savez_compressed(
        os.path.join(OUT, "golden_so400m.npz"), frame_ids=np.array([0, 1, 30, 31]), pooled=o.pooler_output.numpy(),
        projected=proj.numpy(), tokens_row0=o.last_hidden_state[:, 0].numpy(), weight_checksum=np.array([wsum]))
    print("pooled stats", o.pooler_output.mean().item(), o.pooler_output.std().item())
    pn = torch.nn.functional.normalize(proj, dim=1)
    print("projected pairwise cosine\n", (pn @ pn.T).numpy())

    # ---- the reference's own encode_image call sequence with the model injected ----
    try:
        sys.path.insert(0, "/root/reference/src")
        from perception.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder

        class Shim(torch.nn.Module):  # transformers>=5 returns an output object; the reference expects a tensor
            def __init__(self, vm):
                super().__init__()
                self.vm = vm

            @property
            def device(self):
                return torch.device("cpu")

            def get_image_features(self, pixel_values=None, **kw):
                return self.vm(pixel_values=pixel_values).pooler_output

        enc = SigLIPSemanticEncoder(NaFlexConfig(device="cpu", dtype=torch.float32))
        enc.encoder._model = Shim(m)
        enc.encoder._processor = hf_processor(384, 2)
        e0 = enc.encode_image(Image.fromarray(frames[0]))
        print("reference encode_image vs golden pooled[0]: max abs diff", (e0 - o.pooler_output[0]).abs().max().item())
        assert torch.allclose(e0, o.pooler_output[0], atol=1e-5)
    except ImportError as exc:  # reference tree not mounted
        print("reference classes not importable:", exc)


if __name__ == "__main__":
    main()
