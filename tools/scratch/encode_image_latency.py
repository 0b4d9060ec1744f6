"""Per-frame latency of the reference's own loop (scripts/extract_features.py:593-603): encode_image(PIL) + .cpu()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from PIL import Image
from gameplay_vision_llm_b200 import _lib, ops, synth
from gameplay_vision_llm_b200.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_siglip_state_dict
spec = SiglipVisionSpec.so400m()
enc = SigLIPSemanticEncoder(NaFlexConfig(device="cuda:0", state_dict=synth_siglip_state_dict(spec, seed=0)))
frames = [Image.fromarray(f) for f in synth.scene_frames_np(0, 8)]
for f in frames[:3]: enc.encode_image(f).cpu()
torch.cuda.synchronize(); t = time.perf_counter(); n = 40
for i in range(n): enc.encode_image(frames[i % 8]).cpu()
dt = (time.perf_counter() - t) / n
print(f"encode_image + .cpu(): {dt * 1e3:.3f} ms per frame = {1 / dt:.1f} frames/s")
m = enc.encoder._model
arr = np.asarray(frames[0])
def T(fn, n=40):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
ms, _ = T(lambda: np.asarray(frames[0].convert("RGB"), dtype=np.uint8)); print(f"  PIL -> numpy           {ms:.3f} ms")
ms, dev = T(lambda: torch.from_numpy(arr[None]).to("cuda:0")); print(f"  H2D 6.2 MB (pageable)  {ms:.3f} ms")
ms, pv = T(lambda: ops.preprocess(dev, 384, 384, 2, layout=ops.LAYOUT_F32_CHW)); print(f"  preprocess -> f32 CHW  {ms:.3f} ms")
ms, pt = T(lambda: ops.patchify(pv, 14, 592)); print(f"  patchify               {ms:.3f} ms")
l0 = _lib.launch_count()
ms, emb = T(lambda: m.forward_patches(pt)); print(f"  tower forward (B = 1)  {ms:.3f} ms, {(_lib.launch_count() - l0) // 40} launches")
ms, _ = T(lambda: emb.cpu()); print(f"  .cpu()                 {ms:.3f} ms")
# batch via encode_frames for comparison
ms, _ = T(lambda: enc.encode_frames(np.stack([np.asarray(f) for f in frames])).cpu(), 10); print(f"encode_frames(8 frames) {ms:.3f} ms = {8 / ms * 1e3:.0f} frames/s")
