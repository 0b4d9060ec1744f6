"""In-step A/B on ONE box: the batch-64 embedding step with the attention kernel launched with persistent CTAs (ships)
against the same kernel launched with one work item per CTA (round 1's schedule), alternating, 24 steps per pass.
    python tools/scratch/attn_instep_ab.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gameplay_vision_llm_b200 import _lib, synth
from gameplay_vision_llm_b200.pipeline import EmbeddingPipeline
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_projector_state_dict, synth_siglip_state_dict

dev = torch.device("cuda:0")
spec = SiglipVisionSpec.so400m()
sd, psd = synth_siglip_state_dict(spec, seed=0), synth_projector_state_dict(spec.hidden, 4096, seed=1)
pipe = EmbeddingPipeline(sd, psd, spec, dev, batch=64)
frames = torch.cat([synth.scene_frames(j, 16, 1080, 1920, device=dev) for j in range(0, 64, 16)])
lib = _lib.lib()
lib.gvl_debug_set_attn_ctas_per_sm.argtypes = [ctypes.c_int]
lib.gvl_debug_set_attn_ctas_per_sm.restype = None


def run(steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        pipe.embed(frames)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


run(6)
for n in (2, 0, 2, 0, 2, 0):
    lib.gvl_debug_set_attn_ctas_per_sm(n)
    run(3)
    ms = run(24)
    print(f"attention {'persistent CTAs (2 per SM)' if n else 'one item per CTA      '}: {ms:.3f} ms per step, {64e3 / ms:.1f} frames/s")
lib.gvl_debug_set_attn_ctas_per_sm(2)
