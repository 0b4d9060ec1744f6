"""Experiment: does running two half-batch pipelines on two streams (GEMMs capped to part of the machine, so the other
stream's attention / GEMM co-runs) beat one batch-64 pipeline under the 1 kW power cap?
    GVL_GEMM_CLUSTERS=37 python tools/scratch/overlap_bench.py --streams 2 --batch 32"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gameplay_vision_llm_b200 import synth
from gameplay_vision_llm_b200.pipeline import EmbeddingPipeline
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_projector_state_dict, synth_siglip_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
spec = SiglipVisionSpec.so400m()
sd, psd = synth_siglip_state_dict(spec, seed=0), synth_projector_state_dict(spec.hidden, 4096, seed=1)
pipes = [EmbeddingPipeline(sd, psd, spec, dev, batch=a.batch) for _ in range(a.streams)]
streams = [torch.cuda.Stream(device=dev) for _ in range(a.streams)]
frames = [torch.cat([synth.scene_frames(64 * i + j, 16, 1080, 1920, device=dev) for j in range(0, a.batch, 16)]) for i in range(a.streams)]
def run(n):
    for s in range(n):
        for p, st, f in zip(pipes, streams, frames):
            with torch.cuda.stream(st):
                p.embed(f)
run(3); torch.cuda.synchronize()
t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
run(a.steps)
for st in streams: torch.cuda.current_stream().wait_stream(st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"streams {a.streams} batch {a.batch} clusters {os.environ.get('GVL_GEMM_CLUSTERS','all')}: {a.steps * a.streams * a.batch / ms * 1e3:.1f} frames/s ({ms / a.steps:.2f} ms per round)")
