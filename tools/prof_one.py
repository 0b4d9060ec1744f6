#!/usr/bin/env python
"""Run ONE kernel case a few times (for `ncu -s 1 -c 1` captures): python tools/prof_one.py qkv|out|fc1|fc2|attn|pre|ln"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gameplay_vision_llm_b200 import ops, synth  # noqa: E402

DEV = "cuda:0"
M = 64 * 729
SHAPES = {"patch": (M, 1152, 592, 0, 729), "qkv": (M, 3456, 1152, 0, 0), "out": (M, 1152, 1152, 0, 1),
          "fc1": (M, 4304, 1152, 1, 0), "fc2": (M, 1152, 4304, 0, 1), "kv": (M, 2304, 1152, 0, 0)}


def main():
    what = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if what in SHAPES:
        m, n, k, act, res = SHAPES[what]
        a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
        w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
        bias = torch.randn(n, device=DEV)
        out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        r, mod = None, 0
        if res == 1:
            r = out
        elif res > 1:
            r, mod = torch.randn(res, n, device=DEV).to(torch.bfloat16), res
        fn = lambda: ops.gemm(a, w, bias, r, res_row_mod=mod, act=act, out=out)  # noqa: E731
    elif what == "attn":
        qkv = torch.randn(M, 3 * 1152, device=DEV).to(torch.bfloat16)
        out = torch.empty(M, 1152, device=DEV, dtype=torch.bfloat16)
        fn = lambda: ops.attention(qkv, 64, 729, 16, 72, out=out)  # noqa: E731
    elif what == "ln":
        x = torch.randn(M, 1152, device=DEV).to(torch.bfloat16)
        g, b = torch.ones(1152, device=DEV), torch.zeros(1152, device=DEV)
        out = torch.empty_like(x)
        fn = lambda: ops.layernorm(x, g, b, 1e-6, out=out)  # noqa: E731
    elif what == "pre":
        frames = synth.noise_frames(64, seed=1).to(DEV)
        out = torch.empty(M, 592, device=DEV, dtype=torch.bfloat16)
        fn = lambda: ops.preprocess(frames, 384, 384, 2, out=out)  # noqa: E731
    else:
        raise SystemExit(f"unknown case {what}")
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print("done", what)


if __name__ == "__main__":
    main()
