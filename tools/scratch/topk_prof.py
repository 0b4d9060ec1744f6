"""Tuning aid: per-kernel-family device time of the 72k x 4096 retrieval (gvl_prof events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gameplay_vision_llm_b200 import _lib, ops
DEV = "cuda:0"
N, D, Q, k = 72000, 4096, 128, 16
g = torch.Generator(device=DEV).manual_seed(5)
centers = torch.randn(N // 60, D, device=DEV, generator=g)
index = (centers.repeat_interleave(60, 0) + 0.35 * torch.randn(N, D, device=DEV, generator=g)).to(torch.bfloat16)
qs = torch.randint(0, N // 60, (Q,), device=DEV, generator=g)
queries = (centers[qs] + 0.35 * torch.randn(Q, D, device=DEV, generator=g)).to(torch.bfloat16)
inv = ops.row_inv_norm(index)
for mode, name in ((ops.TOPK_TENSOR, "tensor"),):
    for _ in range(3): ops.topk_cosine(index, queries, k, mode=mode, inv_norm=inv)
    torch.cuda.synchronize()
    _lib.prof_enable(True)
    for _ in range(10): ops.topk_cosine(index, queries, k, mode=mode, inv_norm=inv)
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    for fam, v in _lib.prof_summary().items():
        print(f"{name}: {fam:14s} {v['ms'] / 10 * 1e3:8.1f} us per search ({v['launches'] // 10} launches)")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.topk_cosine(index, queries, k, mode=mode, inv_norm=inv)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: whole call {e0.elapsed_time(e1) / 10 * 1e3:.1f} us  ({N * D * 2 / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9:.0f} GB/s of index)")
