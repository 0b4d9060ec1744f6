"""Golden vectors for BASELINE.json configs[0] at full size: HF fp32 `SiglipImageProcessor` -> `SiglipVisionModel`
(so400m spec, synthetic weights seed 0) -> the reference's `MultiModalProjector` (seed 1) on

  * the 64 G1 scene frames 0..63 (the config's own batch),
  * 8 G0 iid-noise frames (seed 1000; the worst case for softmax / LayerNorm statistics),
  * 16 held-out G1 query frames 3000..3015 (scenes 100+, none in the index; SURVEY.md §8d "seed 3000"),

plus the float64 cosine top-16 of the 16 queries over the 64 projected index rows with the score margins
between consecutive ranks.  ~3 minutes of CPU in the build container:

    python tests/golden/make_golden_config1.py

Output: golden_config1.npz (fp32 pooled [88,1152] / projected [88,4096], top-k indices, margins).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import hf_model, hf_processor, ref_projector  # noqa: E402

from gameplay_vision_llm_b200 import synth  # noqa: E402
from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_projector_state_dict,  # noqa: E402
                                                synth_siglip_state_dict)
from oracle import siglip_ref  # noqa: E402

N_INDEX, N_NOISE, N_QUERY, QUERY_START, TOP_K = 64, 8, 16, 3000, 16


def frames_for(name: str) -> np.ndarray:
    if name == "index":
        return synth.scene_frames_np(0, N_INDEX)
    if name == "noise":
        return synth.noise_frames(N_NOISE, seed=1000).numpy()
    return synth.scene_frames_np(QUERY_START, N_QUERY)


def main():
    from PIL import Image
    torch.set_num_threads(os.cpu_count())
    spec = SiglipVisionSpec.so400m()
    sd = synth_siglip_state_dict(spec, seed=0)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    model = hf_model(spec, sd)
    proc = hf_processor(384, 2)
    proj = ref_projector(spec.hidden, synth_projector_state_dict(spec.hidden, 4096, seed=1))
    pooled, projected = [], []
    for name in ("index", "noise", "query"):
        frames = frames_for(name)
        for i0 in range(0, len(frames), 8):  # batch 8, like the reference CPU arm
            pv = proc(images=[Image.fromarray(f) for f in frames[i0:i0 + 8]], return_tensors="pt")["pixel_values"]
            with torch.no_grad():
                p = model(pixel_values=pv).pooler_output
                pooled.append(p.numpy())
                projected.append(proj(p).numpy())
            print(name, i0, flush=True)
    pooled, projected = np.concatenate(pooled), np.concatenate(projected)
    idx_rows, q_rows = projected[:N_INDEX], projected[N_INDEX + N_NOISE:]
    scores, idx, _ = siglip_ref.cosine_topk(idx_rows, q_rows, TOP_K)
    # margins between consecutive ranks, and between rank 16 and the best row left out
    full = siglip_ref.cosine_scores(idx_rows, q_rows) if hasattr(siglip_ref, "cosine_scores") else None
    if full is None:
        a = idx_rows.astype(np.float64)
        q = q_rows.astype(np.float64)
        full = (q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-12)) @ (
            a / np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-12)).T
    srt = -np.sort(-full, axis=1)
    margins = srt[:, :TOP_K] - srt[:, 1:TOP_K + 1]
    print("min margin inside the top-16 (incl. rank 16 vs 17):", margins.min())
    np.savez_compressed(os.path.join(HERE, "golden_config1.npz"), pooled=pooled.astype(np.float32),
                        projected=projected.astype(np.float32), topk_idx=idx.astype(np.int32),
                        topk_scores=scores.astype(np.float64), margins=margins, weight_checksum=np.array([wsum]),
                        layout=np.array([N_INDEX, N_NOISE, N_QUERY, QUERY_START, TOP_K]))


if __name__ == "__main__":
    main()
