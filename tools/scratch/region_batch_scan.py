"""How much does a larger ragged pass buy?  16 / 32 / 64 detections in one pass (same frame, masks repeated)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gameplay_vision_llm_b200 import synth
from gameplay_vision_llm_b200.siglip_semantic_encoder import BoxMask, NaFlexConfig, SigLIPSemanticEncoder
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_ren_projection_state_dict, synth_siglip_state_dict
from oracle import hf_baseline
spec = SiglipVisionSpec.so400m()
enc = SigLIPSemanticEncoder(NaFlexConfig(device="cuda:0", state_dict=synth_siglip_state_dict(spec, seed=0), batch_size=16))
enc.projection.load_state_dict(synth_ren_projection_state_dict(spec.hidden, seed=3))
fd = torch.from_numpy(synth.scene_frames_np(0, 1)[0]).cuda()
base = [(f"d{i}", BoxMask((1080, 1920), y1, y2, x1, x2)) for i, (x1, y1, x2, y2) in enumerate(synth.region_boxes(16))]
for rep in (1, 2, 4, 8):
    masks = base * rep
    for _ in range(3): enc.encode_regions_individually(fd, masks, max_tokens=65536)
    torch.cuda.synchronize(); t = time.perf_counter()
    n = 10
    for _ in range(n): enc.encode_regions_individually(fd, masks, max_tokens=65536)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n
    print(f"{16 * rep:4d} detections per pass: {dt * 1e3:.2f} ms, {16 * rep / dt:.0f} regions/s")
