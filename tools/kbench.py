#!/usr/bin/env python
"""Per-kernel microbenchmark at the batch-64 shapes of the headline workload (CUDA events, isolated kernels).
Not the bench of record (bench.py is) — a tuning aid: python tools/kbench.py [gemm|attn|ln|pre|all]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gameplay_vision_llm_b200 import ops, synth  # noqa: E402

DEV = "cuda:0"
M = 64 * 729


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gemm_cases():
    shapes = [("patch", M, 1152, 592, 0, 729), ("qkv", M, 3456, 1152, 0, 0), ("out", M, 1152, 1152, 0, 1),
              ("fc1", M, 4304, 1152, 1, 0), ("fc2", M, 1152, 4304, 0, 1), ("kv", M, 2304, 1152, 0, 0),
              ("proj1", 64, 4096, 1152, 2, 0), ("proj2", 64, 4096, 4096, 0, 0)]
    tot = 0.0
    for name, m, n, k, act, res in shapes:
        a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
        w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
        bias = torch.randn(n, device=DEV)
        out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        r, mod = None, 0
        if res == 1:
            r = out  # in-place residual like the tower
        elif res > 1:
            r, mod = torch.randn(res, n, device=DEV).to(torch.bfloat16), res
        ms = timeit(lambda: ops.gemm(a, w, bias, r, res_row_mod=mod, act=act, out=out))
        tf = 2.0 * m * n * k / ms / 1e9
        per_step = {"qkv": 27, "out": 27, "fc1": 27, "fc2": 27}.get(name, 1)
        tot += ms * per_step
        print(f"gemm {name:6s} M={m} N={n} K={k} act={act} res={res}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s  (x{per_step}/step)")
    print(f"gemm total per step ~ {tot:.2f} ms")


def gemm_probe_cases():
    """Epilogue-cost probes: the same shapes with / without activation and residual (run under GVL_GEMM_BN=192/128
    for the tile-width variants; the environment knob is read once per process)."""
    for name, n, k, act, res in [("fc1", 4304, 1152, 1, 0), ("fc1-noact", 4304, 1152, 0, 0), ("out", 1152, 1152, 0, 1),
                                 ("out-nores", 1152, 1152, 0, 0), ("fc2", 1152, 4304, 0, 1), ("fc2-nores", 1152, 4304, 0, 0),
                                 ("qkv", 3456, 1152, 0, 0), ("qkv-f32out", 3456, 1152, 0, -1)]:
        a = torch.randn(M, k, device=DEV).to(torch.bfloat16)
        w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
        bias = torch.randn(n, device=DEV)
        out = torch.empty(M, n, device=DEV, dtype=torch.float32 if res < 0 else torch.bfloat16)
        ms = timeit(lambda: ops.gemm(a, w, bias, out if res == 1 else None, act=act, out=out))
        print(f"gemm {name:10s} N={n} K={k} BN={os.environ.get('GVL_GEMM_BN', 'auto')}: {ms*1e3:8.1f} us  "
              f"{2.0*M*n*k/ms/1e9:7.1f} TFLOP/s")


def gemm_fused_cases():
    """The LayerNorm-fusion variants at the layer shapes (producer = stats_out, consumer = ln_stats)."""
    D = 1152
    slots = ops.gemm_stats_slots(D)
    stats = torch.zeros(M, slots, 2, device=DEV)
    x = torch.randn(M, D, device=DEV).to(torch.bfloat16)
    for name, n, k, act, mode in [("qkv+ln", 3456, D, 0, "c"), ("fc1+ln", 4304, D, 1, "c"), ("out+st", D, D, 0, "p"),
                                  ("fc2+st", D, 4304, 0, "p")]:
        a = torch.randn(M, k, device=DEV).to(torch.bfloat16)
        w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
        bias = torch.randn(n, device=DEV)
        out = torch.empty(M, n, device=DEV, dtype=torch.bfloat16) if mode == "c" else x
        c1 = torch.randn(n, device=DEV)
        if mode == "c":
            fn = lambda: ops.gemm(a, w, bias, act=act, out=out, ln_stats=stats, ln_c1=c1, ln_dim=D, ln_eps=1e-6)  # noqa: E731
        else:
            fn = lambda: ops.gemm(a, w, bias, residual=x, out=x, stats_out=stats)  # noqa: E731
        ms = timeit(fn)
        print(f"gemm {name:7s} N={n} K={k}: {ms*1e3:8.1f} us  {2.0*M*n*k/ms/1e9:7.1f} TFLOP/s")


def sm_clock_during(fn, reps):
    """median SM clock (MHz) and power (W) sampled through NVML while fn runs reps times back to back"""
    import threading
    import pynvml
    pynvml.nvmlInit()
    hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
    clk, pwr, stop = [], [], threading.Event()

    def poll():
        while not stop.is_set():
            clk.append(pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM))
            pwr.append(pynvml.nvmlDeviceGetPowerUsage(hdl) / 1e3)
            stop.wait(0.002)

    th = threading.Thread(target=poll)
    th.start()
    ms = timeit(fn, reps=reps, warm=3)
    stop.set()
    th.join()
    clk.sort()
    return ms, clk[len(clk) // 2], max(pwr)


def attn_case():
    B, T, H, hd = 64, 729, 16, 72
    qkv = torch.randn(B * T, 3 * H * hd, device=DEV).to(torch.bfloat16)
    out = torch.empty(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, B, T, H, hd, out=out))
    print(f"attention B={B} T={T}: {ms*1e3:.1f} us  {4.0*B*H*T*T*hd/ms/1e9:.1f} TFLOP/s  (x27/step = {ms*27:.2f} ms)")
    import ctypes

    from gameplay_vision_llm_b200 import _lib
    lib = _lib.lib()
    lib.gvl_debug_set_attn_ctas_per_sm.argtypes = [ctypes.c_int]
    lib.gvl_debug_set_attn_ctas_per_sm.restype = None
    for n in (0, 1, 2, 0, 2):  # A/B in one process: 0 = one work item per CTA (round 1's launch), 2 = persistent (ships)
        lib.gvl_debug_set_attn_ctas_per_sm(n)
        ms = timeit(lambda: ops.attention(qkv, B, T, H, hd, out=out), reps=50)
        print(f"attention CTAs/SM={n} ({'one item per CTA' if n == 0 else 'persistent'}): {ms*1e3:.1f} us")
    lib.gvl_debug_set_attn_ctas_per_sm(2)
    ms, mhz, watts = sm_clock_during(lambda: ops.attention(qkv, B, T, H, hd, out=out), reps=600)
    print(f"attention sustained (600 launches): {ms*1e3:.1f} us at a median {mhz} MHz SM clock, peak {watts:.0f} W")


def ln_case():
    x = torch.randn(M, 1152, device=DEV).to(torch.bfloat16)
    g = torch.ones(1152, device=DEV)
    b = torch.zeros(1152, device=DEV)
    out = torch.empty_like(x)
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-6, out=out))
    print(f"layernorm {M}x1152: {ms*1e3:.1f} us  {M*1152*4/ms/1e6:.0f} GB/s  (x56/step = {ms*56:.2f} ms)")


def pre_case():
    frames = synth.noise_frames(64, seed=1).to(DEV)
    out = torch.empty(64 * 729, 592, device=DEV, dtype=torch.bfloat16)
    for rs, path, name in ((2, ops.PRE_PATH_AUTO, "streaming 5:1"), (2, ops.PRE_PATH_PLANAR, "planar"),
                           (3, ops.PRE_PATH_AUTO, "planar")):
        with ops.preprocess_path(path):
            ms = timeit(lambda: ops.preprocess(frames, 384, 384, rs, out=out))
        gb = 64 * (1080 * 1920 * 3 + 729 * 588 * 2) / ms / 1e6
        print(f"preprocess 64x1080p rs={rs} [{name}]: {ms*1e3:.1f} us  {gb:.0f} GB/s "
              f"({gb/6552.6*100:.1f}% of measured HBM peak)")


def sustained_cases():
    """Ours vs cuBLAS (torch.matmul, no epilogue) on the tower's GEMM shapes, each run back to back for ~1.5 s so the
    power governor settles (MEASURED_PEAKS.json's 'sustained' figure is the same procedure on 8192^3)."""
    for name, n, k, act, res in [("qkv", 3456, 1152, 0, 0), ("fc1", 4304, 1152, 1, 0), ("fc2", 1152, 4304, 0, 1),
                                 ("out", 1152, 1152, 0, 1), ("sq8192", 8192, 8192, 0, 0)]:
        m = 8192 if name == "sq8192" else M
        a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
        w = (torch.randn(n, k, device=DEV) / math.sqrt(k)).to(torch.bfloat16)
        bias = torch.randn(n, device=DEV)
        out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
        wt = w.t()
        flops = 2.0 * m * n * k
        probe = timeit(lambda: ops.gemm(a, w, bias, out if res else None, act=act, out=out), reps=5)
        reps = max(50, int(1500 / probe))
        ms, mhz, watts = sm_clock_during(lambda: ops.gemm(a, w, bias, out if res else None, act=act, out=out), reps)
        ms2, mhz2, watts2 = sm_clock_during(lambda: torch.matmul(a, wt, out=out), reps)
        line = (f"sustained {name:7s} M={m} N={n} K={k}: ours {flops/ms/1e9:7.1f} TFLOP/s @ {mhz} MHz {watts:.0f} W | "
                f"cuBLAS (no bias/act/residual) {flops/ms2/1e9:7.1f} TFLOP/s @ {mhz2} MHz {watts2:.0f} W")
        if res:  # the same work: D = A.W^T + C with C = D in place (beta = 1), still without the bias
            ms3, mhz3, watts3 = sm_clock_during(lambda: torch.addmm(out, a, wt, out=out), reps)
            line += f" | cuBLAS + residual (addmm, beta=1) {flops/ms3/1e9:7.1f} TFLOP/s @ {mhz3} MHz {watts3:.0f} W"
        print(line)


def topk_case():
    """configs[4] retrieval: top-16 of 128 queries over a 72 000 x 4096 bf16 index, per-kernel-family breakdown."""
    from gameplay_vision_llm_b200 import _lib
    N, D, Q, k = 72000, 4096, 128, 16
    index = torch.randn(N, D, device=DEV).to(torch.bfloat16)
    queries = torch.randn(Q, D, device=DEV).to(torch.bfloat16)
    inv = ops.row_inv_norm(index)
    for name, mode, kw in (("scan", ops.TOPK_SCAN, {}), ("tensor", ops.TOPK_TENSOR, {}),
                           ("tensor+cached", ops.TOPK_TENSOR, {"inv_norm": inv})):
        ms = timeit(lambda: ops.topk_cosine(index, queries, k, mode=mode, **kw), reps=5, warm=2)
        _lib.prof_enable(True)
        ops.topk_cosine(index, queries, k, mode=mode, **kw)
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        parts = ", ".join(f"{f} {v['ms']*1e3:.0f} us x{v['launches']}" for f, v in _lib.prof_summary().items())
        print(f"topk {name:14s}: {ms*1e3:8.1f} us  ({N*D*2/ms/1e6:.0f} GB/s of index)  [{parts}]")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gemm", "all"):
        gemm_cases()
    if what == "gemmx":
        gemm_probe_cases()
    if what in ("gemmf", "all"):
        gemm_fused_cases()
    if what in ("attn", "all"):
        attn_case()
    if what in ("ln", "all"):
        ln_case()
    if what in ("pre", "all"):
        pre_case()
    if what == "sustained":
        sustained_cases()
    if what in ("topk", "all"):
        topk_case()
