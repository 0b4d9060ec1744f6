"""Tuning aid: practical read-only HBM bandwidth on a (72000, 4096) bf16 matrix (torch reduction, our row-norm kernel)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gameplay_vision_llm_b200 import ops
DEV = "cuda:0"
N, D = 72000, 4096
x = torch.randn(N, D, device=DEV).to(torch.bfloat16)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
b = N * D * 2
xi = x.view(torch.int32)
for name, fn in (("torch int32 sum", lambda: xi.sum()), ("torch bf16 amax", lambda: x.amax()),
                 ("gvl row_inv_norm", lambda: ops.row_inv_norm(x)),
                 ("torch copy (r+w)", lambda: x.clone())):
    ms = timeit(fn)
    tot = b * (2 if "copy" in name else 1)
    print(f"{name:20s} {ms*1e3:8.1f} us  {tot/ms/1e6:7.0f} GB/s")
