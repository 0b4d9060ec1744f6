"""Tuning aid: attention time vs sequence length (same 128-row tiles) -> per-block steady-state cost and per-CTA start-up."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gameplay_vision_llm_b200 import ops
DEV = "cuda:0"
def timeit(fn, reps=10, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
H, hd = 16, 72
for B, T in [(64, 729), (64, 768), (32, 1536), (16, 3072), (128, 384), (256, 192), (64, 128)]:
    qkv = torch.randn(B * T, 3 * H * hd, device=DEV).to(torch.bfloat16)
    out = torch.empty(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, B, T, H, hd, out=out))
    tiles = B * H * ((T + 127) // 128); blocks = (T + 63) // 64
    cyc_per_cta = ms * 1e-3 * 1.965e9 * 148 * 2 / tiles
    print(f"B={B} T={T}: {ms*1e3:.1f} us  tiles={tiles} blocks/tile={blocks}  ~{cyc_per_cta:.0f} cycles per CTA (2 resident/SM)  {cyc_per_cta/blocks:.0f}/block")
    del qkv, out
