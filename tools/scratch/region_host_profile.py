"""Where does the host time of encode_regions_individually go?  (per 16-detection 1080p frame)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gameplay_vision_llm_b200 import ops, synth
from gameplay_vision_llm_b200.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_ren_projection_state_dict, synth_siglip_state_dict
from oracle import hf_baseline
spec = SiglipVisionSpec.so400m()
enc = SigLIPSemanticEncoder(NaFlexConfig(device="cuda:0", state_dict=synth_siglip_state_dict(spec, seed=0), batch_size=16))
enc.projection.load_state_dict(synth_ren_projection_state_dict(spec.hidden, seed=3))
frame = synth.scene_frames_np(0, 1)[0]
fd = torch.from_numpy(frame).cuda()
masks = []
for i, (x1, y1, x2, y2) in enumerate(synth.region_boxes(16)):
    m = np.zeros((1080, 1920), np.bool_); m[y1:y2, x1:x2] = True; masks.append((f"d{i}", m))
for _ in range(3): enc.encode_regions_individually(fd, masks)
def T(fn, n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
ms, _ = T(lambda: enc.encode_regions_individually(fd, masks)); print(f"whole call          {ms:.3f} ms")
ms, prep = T(lambda: enc._prepare_regions(fd, masks)); print(f"_prepare_regions    {ms:.3f} ms")
ms, _ = T(lambda: [enc.region_extractor.region_bbox(frame.shape, m) for _, m in masks]); print(f"  16 x region_bbox  {ms:.3f} ms")
_, boxes, sizes, meta = prep
order = sorted(range(16), key=lambda i: sizes[i])
ms, (patches, _) = T(lambda: ops.region_patches(fd, [boxes[i] for i in order], [sizes[i] for i in order], None, enc._region_lut, patch=14, ld=592)); print(f"region_patches      {ms:.3f} ms")
groups = {}
for i in order: groups.setdefault(sizes[i], []).append(i)
shapes = [(len(v), k[0] // 14, k[1] // 14) for k, v in groups.items()]
m = enc.encoder._model
M = sum(b * gh * gw for b, gh, gw in shapes)
ws = m._ws_bytes(ops.siglip_ragged_workspace_bytes(m.pack, M, 16))
ms, (pooled, tokens) = T(lambda: ops.siglip_forward_ragged(m.pack, patches, shapes, workspace=ws, return_tokens=True)); print(f"forward_ragged      {ms:.3f} ms (host+device)")
t = time.perf_counter()
for _ in range(20): ops.siglip_forward_ragged(m.pack, patches, shapes, workspace=ws, return_tokens=True)
print(f"  host enqueue only {(time.perf_counter() - t) / 20 * 1e3:.3f} ms"); torch.cuda.synchronize()
def pool():
    feats, r0 = [], 0
    for b, gh, gw in shapes:
        feats.append(ops.mean_tokens(tokens[r0:r0 + b * gh * gw], b, gh * gw)); r0 += b * gh * gw
    return torch.cat(feats, 0)
ms, feats = T(pool); print(f"pooling             {ms:.3f} ms")
ms, _ = T(lambda: enc.projection(feats).cpu()); print(f"projection + .cpu() {ms:.3f} ms")
