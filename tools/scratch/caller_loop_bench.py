"""The reference's caller-level function on its own data model: run_siglip_encoder([(ts, PIL 1080p frame)], ...)
(scripts/extract_features.py:502-610, fallback branch) — frames/s from a list of PIL images to CPU embeddings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from PIL import Image
from gameplay_vision_llm_b200 import synth
from gameplay_vision_llm_b200.pipeline import run_siglip_encoder
from gameplay_vision_llm_b200.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_siglip_state_dict
spec = SiglipVisionSpec.so400m()
sd = synth_siglip_state_dict(spec, seed=0)
N = 128
frames = [(i / 1.0, Image.fromarray(f)) for i, f in enumerate(synth.scene_frames_np(0, N))]
for bs in (16, 64):
    enc = SigLIPSemanticEncoder(NaFlexConfig(device="cuda:0", state_dict=sd, batch_size=bs, fold_layernorm=True))
    run_siglip_encoder(frames[:bs], "cuda:0", encoder=enc)
    torch.cuda.synchronize(); t = time.perf_counter()
    out = run_siglip_encoder(frames, "cuda:0", encoder=enc)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    assert len(out) == N
    print(f"run_siglip_encoder, {N} PIL 1080p frames, batch_size {bs}: {N / dt:.1f} frames/s ({dt / N * 1e3:.2f} ms per frame)")
