#!/usr/bin/env python
"""Timeline of one attention CTA (SM-clock stamps written by attention_sdb.cu when a trace buffer is set): two consecutive
work items of the same persistent CTA, block by block, and the O read-out between them.

The stamps are compiled out of the shipped library (they cost ~10 % of the kernel).  Build a traced copy and point the
tool at it:
    cd gameplay_vision_llm_b200/csrc && mkdir -p /tmp/bt && for f in *.cu; do nvcc -gencode arch=compute_100a,code=sm_100a \
        -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DGVL_ATTN_TRACE -c $f -o /tmp/bt/${f%.cu}.o; done && \
        nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgvl_trace.so /tmp/bt/*.o -cudart static -ldl
    GVL_TRACE_LIB=$PWD/gameplay_vision_llm_b200/libgvl_trace.so python tools/attn_trace.py
-> per key block: softmax warp 0 phases, MMA thread phases, TMA refill time (cycles); per item: o_done wait + read-out."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gameplay_vision_llm_b200 import _lib, ops  # noqa: E402

if os.environ.get("GVL_TRACE_LIB"):  # a library built with EXTRA=-DGVL_ATTN_TRACE (the stamps are compiled out of the shipped one)
    _lib.LIB_PATH = os.environ["GVL_TRACE_LIB"]

B, T, H, hd = 64, 729, 16, 72
qkv = torch.randn(B * T, 3 * H * hd, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, H * hd, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(32 * 16, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.attention(qkv, B, T, H, hd, out=out)
torch.cuda.synchronize()
lib = _lib.lib()
lib.gvl_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
lib.gvl_debug_set_attn_trace(trace.data_ptr())
ops.attention(qkv, B, T, H, hd, out=out)
torch.cuda.synchronize()
lib.gvl_debug_set_attn_trace(None)
t = trace.cpu().view(32, 16).numpy()
t0 = t[0, 0]
print("rows 0-12: the traced item (x 1, head 3, image 17); rows 16-28: the item the same CTA runs next")
print("block | softmax: wait_S  ld  max+exp  st_wait  arrive | next-block gap || MMA: idle->p_full  issue_PV  issue_S(j+2) || TMA kv_empty at")
for base in (0, 16):
    for j in range(12):
        r = t[base + j]
        nxt = t[base + j + 1, 0] - r[5] if j + 1 < 12 else 0
        print(f"{j:5d} | start {r[0]-t0:7d}: {r[1]-r[0]:6d} {r[2]-r[1]:5d} {r[3]-r[2]:7d} {r[4]-r[3]:7d} {r[5]-r[4]:6d} | {nxt:5d} || "
              f"wait from {r[8]-t0:7d}: {r[9]-r[8]:6d} {r[10]-r[9]:6d} {r[11]-r[10]:6d} || {r[12]-t0:7d}")
    f = t[base + 12]
    print(f"finalise: reached {f[0]-t0:7d}, o_done wait {f[1]-f[0]:6d}, L + O read-out + stores {f[2]-f[1]:6d}, done at {f[2]-t0:7d}")
    print("block period (softmax warp 0):", [int(t[base + j + 1, 0] - t[base + j, 0]) for j in range(11)])
print("item period (block 0 to block 0):", int(t[16, 0] - t[0, 0]))
