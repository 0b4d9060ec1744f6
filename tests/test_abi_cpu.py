"""The C-ABI library loads without a GPU, exports every symbol include/gvl.h declares, and its host-only
entry points agree with the oracle.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from gameplay_vision_llm_b200 import _lib, ops
from oracle import preprocess_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gvl.h")).read()
    return sorted(set(re.findall(r"GVL_API[^;(]*?\b(gvl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 15
    handle = _lib.lib()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/gvl.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert handle.gvl_abi_version() == _lib.ABI_VERSION == int(re.search(r"#define GVL_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "gvl.h")).read()).group(1))
    assert isinstance(_lib.launch_count(), int)


@pytest.mark.parametrize("in_size,out_size", [(1920, 384), (1080, 384), (1080, 224), (211, 56), (1280, 384), (400, 399)])
@pytest.mark.parametrize("rs", [2, 3])
def test_resize_taps_match_oracle(in_size, out_size, rs):
    xmin, xsize, w, prec = ops.resize_taps(in_size, out_size, rs)
    o_xmin, o_xsize, o_w, o_prec = preprocess_ref.resize_taps(in_size, out_size, rs)
    assert prec == o_prec
    assert np.array_equal(xmin, o_xmin) and np.array_equal(xsize, o_xsize)
    assert w.shape == o_w.shape and np.array_equal(w, o_w)


def test_argument_errors_are_reported_not_crashed():
    handle = _lib.lib()
    rc = handle.gvl_gemm_bf16(None, 8, None, 8, None, None, 0, 0, None, 8, 0, 1, 8, 8, 0, None)
    assert rc != 0 and b"null pointer" in handle.gvl_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "gvl_gemm_bf16")
    buf = (ctypes.c_int32 * 4)()
    w = (ctypes.c_int16 * 4)()
    p, u = ctypes.c_int(0), ctypes.c_int(0)
    assert handle.gvl_resize_taps(100, 4, 2, 1, buf, buf, w, ctypes.byref(p), ctypes.byref(u)) != 0  # max_taps too small


def test_ops_refuse_cpu_tensors():
    import torch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.preprocess(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_varlen_attention_tile_table_is_built_on_the_host():
    """gvl_attention_varlen_tiles (host only): one {first token, length, first query row, 0} entry per 128-row query tile
    of every item of a ragged batch, longest items first (a tile's cost grows with its item's key count, so the tail of
    the launch is short tiles), equal lengths in input order; the size query and the error path need no device."""
    lib = _lib.lib()
    toks = np.array([130, 729, 40, 729, 128, 257], np.int32)
    tp = toks.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    n = ctypes.c_int(0)
    assert lib.gvl_attention_varlen_tiles(len(toks), tp, None, ctypes.byref(n)) == 0
    want_n = int(sum(-(-t // 128) for t in toks))
    assert n.value == want_n == 2 + 6 + 1 + 6 + 1 + 3
    table = np.full((n.value, 4), -7, np.int32)
    assert lib.gvl_attention_varlen_tiles(len(toks), tp, table.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                          ctypes.byref(n)) == 0
    starts = np.concatenate([[0], np.cumsum(toks)[:-1]])
    order = sorted(range(len(toks)), key=lambda i: -int(toks[i]))  # stable: the two 729-token items keep their order
    want = [(int(starts[i]), int(toks[i]), q0, 0) for i in order for q0 in range(0, int(toks[i]), 128)]
    assert table.tolist() == [list(w) for w in want]
    # every token row is covered by exactly one tile
    covered = np.zeros(int(toks.sum()), np.int32)
    for tok0, length, q0, _ in table.tolist():
        covered[tok0 + q0:tok0 + min(q0 + 128, length)] += 1
    assert (covered == 1).all()
    bad = np.array([5, 0, 3], np.int32)
    assert lib.gvl_attention_varlen_tiles(3, bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), None, ctypes.byref(n)) != 0
    assert b"item 1" in lib.gvl_last_error()
