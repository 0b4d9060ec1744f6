"""ctypes binding of libgvl_sm100a.so (include/gvl.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C gameplay_vision_llm_b200/csrc`.
There is no fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int16, c_int32, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# GVL_LIB_PATH: A/B runs of two builds of the same library (tuning aid; the ABI version is still checked)
LIB_PATH = os.environ.get("GVL_LIB_PATH") or os.path.join(_HERE, "libgvl_sm100a.so")

c_float_p = POINTER(c_float)
ABI_VERSION = 11  # include/gvl.h GVL_ABI_VERSION


class VitLayer(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
        "c1_qkv", "c1_fc1")]


class VitWeights(ctypes.Structure):
    _fields_ = (
        [(n, c_int32) for n in ("D", "I", "H", "hd", "L", "T", "patch_k", "patch_ld")]
        + [("eps", c_float), ("act", c_int32)]
        + [("w_patch", c_void_p), ("b_patch", c_void_p), ("pos", c_void_p), ("layers", POINTER(VitLayer)),
           ("post_g", c_void_p), ("post_b", c_void_p), ("probe_q", c_void_p), ("w_kv", c_void_p), ("b_kv", c_void_p),
           ("w_ho", c_void_p), ("b_ho", c_void_p), ("hln_g", c_void_p), ("hln_b", c_void_p), ("w_hfc1", c_void_p),
           ("b_hfc1", c_void_p), ("w_hfc2", c_void_p), ("b_hfc2", c_void_p), ("fold_ln", c_int32), ("c1_kv", c_void_p)]
    )


class RaggedGroup(ctypes.Structure):
    _fields_ = [("B", c_int32), ("T", c_int32), ("pos", c_void_p)]


class GemmFusion(ctypes.Structure):
    _fields_ = [("stats_out", c_void_p), ("ln_stats", c_void_p), ("ln_slots", c_int32), ("ln_dim", c_int32),
                ("ln_c1", c_void_p), ("ln_eps", c_float)]


# name -> (restype, argtypes); every symbol include/gvl.h declares
SIGNATURES = {
    "gvl_last_error": (c_char_p, []),
    "gvl_abi_version": (c_int, []),
    "gvl_launch_count": (c_ulonglong, []),
    "gvl_check_device": (c_int, [c_int]),
    "gvl_prof_enable": (c_int, [c_int]),
    "gvl_prof_summary": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(c_ulonglong), POINTER(ctypes.c_double)]),
    "gvl_resize_taps": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int16),
                                POINTER(c_int), POINTER(c_int)]),
    "gvl_preprocess_path": (c_int, [c_int]),
    "gvl_preprocess_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float_p, c_float_p, c_void_p,
                                  c_int, c_int, c_int, c_void_p]),
    "gvl_preprocess_u8_crop": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_float_p, c_float_p, c_void_p, c_int, c_void_p]),
    "gvl_preprocess_u8_crop_band": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_float_p, c_float_p, c_void_p, c_int, c_void_p]),
    "gvl_copy_band_h2d": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_longlong, c_void_p]),
    "gvl_patchify_tubelet_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gvl_mean_tokens_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "gvl_videomae_workspace_bytes": (c_size_t, [POINTER(VitWeights), c_int]),
    "gvl_videomae_forward": (c_int, [POINTER(VitWeights), c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_int, c_void_p,
                                     c_void_p]),
    "gvl_patchify_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gvl_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "gvl_gemm_stats_slots": (c_int, [c_int]),
    "gvl_ln_finalize": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "gvl_gemm_bf16_fused": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                    c_int, c_int, c_int, c_int, c_int, POINTER(GemmFusion), c_void_p]),
    "gvl_layernorm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                   c_void_p]),
    "gvl_attention_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "gvl_attention_varlen_tiles": (c_int, [c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int)]),
    "gvl_attention_varlen_bf16": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, ctypes.c_double, c_int, c_int, c_float,
                                          c_void_p]),
    "gvl_probe_attention_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "gvl_siglip_workspace_bytes": (c_size_t, [POINTER(VitWeights), c_int]),
    "gvl_siglip_forward": (c_int, [POINTER(VitWeights), c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                   c_void_p]),
    "gvl_siglip_ragged_workspace_bytes": (c_size_t, [POINTER(VitWeights), c_longlong, c_int]),
    "gvl_siglip_forward_ragged": (c_int, [POINTER(VitWeights), c_void_p, c_int, POINTER(RaggedGroup), c_void_p, c_size_t,
                                          c_void_p, c_void_p, c_void_p]),
    "gvl_project": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                            c_int, c_void_p]),
    "gvl_topk_cosine": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "gvl_topk_scratch_floats": (c_size_t, [c_int, c_int]),
    "gvl_topk_cosine_ex": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gvl_topk_cosine_f32": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "gvl_row_inv_norm": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "gvl_pil_bicubic_taps": (c_int, [c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32),
                                     POINTER(c_int)]),
    "gvl_region_scratch_bytes": (c_size_t, [c_int, POINTER(c_int32)]),
    "gvl_region_patches_pil_u8": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), c_void_p, c_longlong, c_void_p,
                                          c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gvl_resize_two_pass_u8": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_int32), c_void_p, c_longlong, c_int, c_int,
                                       c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_size_t, c_void_p]),
    "gvl_pos_interp_bicubic_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gvl_max_tokens_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "gvl_nvdec_available": (c_int, []),
    "gvl_nvdec_caps": (c_int, [c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "gvl_nvdec_open": (c_int, [c_int, c_int, POINTER(c_void_p)]),
    "gvl_nvdec_sampling": (c_int, [c_void_p, c_longlong, c_longlong, c_int, c_int]),
    "gvl_nvdec_feed": (c_int, [c_void_p, c_char_p, c_size_t, c_int, c_void_p, c_int, c_int, c_int, POINTER(c_int),
                               POINTER(c_longlong), c_void_p]),
    "gvl_nvdec_info": (c_int, [c_void_p, POINTER(c_int32)]),
    "gvl_nvdec_close": (c_int, [c_void_p]),
}

_LIB = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C gameplay_vision_llm_b200/csrc`). There is no CPU/PyTorch fallback for this path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.gvl_abi_version() != ABI_VERSION:
            raise RuntimeError("libgvl_sm100a.so ABI version mismatch; rebuild the extension")
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "gvl") -> None:
    if rc != 0:
        msg = lib().gvl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def launch_count() -> int:
    return int(lib().gvl_launch_count())


KERNEL_IDS = {"preprocess": 0, "gemm": 1, "layernorm": 2, "attention": 3, "probe_attention": 4, "topk_scores": 5,
              "topk_select": 6, "patchify": 7}


def prof_enable(on: bool) -> None:
    check(lib().gvl_prof_enable(1 if on else 0), "gvl_prof_enable")


def prof_summary() -> dict:
    """{kernel family: {"ms": total device ms, "launches": n, "work": algorithmic FLOPs or bytes}} of the session."""
    out = {}
    for name, kid in KERNEL_IDS.items():
        ms, n, work = ctypes.c_double(0), c_ulonglong(0), ctypes.c_double(0)
        check(lib().gvl_prof_summary(kid, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(work)), "gvl_prof_summary")
        if n.value:
            out[name] = {"ms": ms.value, "launches": int(n.value), "work": work.value}
    return out
