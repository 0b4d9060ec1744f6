"""Same GEMM through this library and through cuBLAS (torch.matmul), for an ncu metric comparison per launch:
    ncu --metrics <list> --csv --log-file out.csv python tools/scratch/gemm_vs_cublas_ncu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gameplay_vision_llm_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N, K) in ((8192, 8192, 8192), (46656, 3456, 1152), (46656, 1152, 4304)):
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(N, K, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        ops.gemm(a, w, out=out)
    for _ in range(2):
        torch.matmul(a, w.t(), out=out)
    torch.cuda.synchronize()
