"""Tensor-level wrappers over the C ABI (include/gvl.h).  PyTorch is used for device memory and
streams only; every function below launches hand-written kernels from libgvl_sm100a.so on the current
CUDA stream and raises RuntimeError on any failure (there is no eager/PyTorch fallback).
"""
from __future__ import annotations

import contextlib
import ctypes
import functools

import numpy as np
import torch

from . import _lib
from .weights import ProjectorPack, SiglipGridView, SiglipPack, VideoMAEPack, resolve_device  # noqa: F401

LAYOUT_U8_CHW, LAYOUT_F32_CHW, LAYOUT_BF16_CHW, LAYOUT_BF16_PATCH = 0, 1, 2, 3
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2
BILINEAR, BICUBIC = 2, 3


def _stream() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _on_tensor_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument made current: the library launches on the current
    device and `_stream()` reads that device's current stream, so a tensor on cuda:1 must never be launched from a
    process whose current device is cuda:0.  All tensor arguments must live on that one device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise RuntimeError(f"{fn.__name__}: tensors on different devices ({dev} and {a.device})")
            elif isinstance(a, (SiglipPack, SiglipGridView, ProjectorPack, VideoMAEPack)) and dev is not None and \
                    torch.device(a.device) != dev:
                raise RuntimeError(f"{fn.__name__}: weights on {a.device}, tensors on {dev}")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gvl ops take CUDA tensors only (no CPU fallback exists for this path)")


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def resize_taps(in_size: int, out_size: int, resample: int = BILINEAR, max_taps: int = 64):
    """Host-only: the integer tap tables the preprocessing kernel uses (for parity checks on CPU)."""
    xmin = (ctypes.c_int32 * out_size)()
    xsize = (ctypes.c_int32 * out_size)()
    w = (ctypes.c_int16 * (out_size * max_taps))()
    prec, used = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().gvl_resize_taps(in_size, out_size, resample, max_taps, xmin, xsize, w, ctypes.byref(prec),
                                          ctypes.byref(used)), "gvl_resize_taps")
    wn = np.ctypeslib.as_array(w).reshape(out_size, max_taps)[:, : used.value].copy()
    return np.ctypeslib.as_array(xmin).copy(), np.ctypeslib.as_array(xsize).copy(), wn, prec.value


def fused_sub_div(image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), rescale_factor: float = 1.0 / 255.0):
    """The fp32 constants HF's fused rescale+normalize uses: fp32(mean) * fp32(1/rescale_factor)
    (HF: image_processing_backends.py:292-310)."""
    inv = 1.0 / rescale_factor
    sub = (torch.tensor(list(image_mean), dtype=torch.float32) * inv).numpy().astype(np.float32)
    div = (torch.tensor(list(image_std), dtype=torch.float32) * inv).numpy().astype(np.float32)
    return sub, div


@_on_tensor_device
def preprocess(frames: torch.Tensor, out_h: int = 384, out_w: int = 384, resample: int = BILINEAR,
               image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), layout: int = LAYOUT_BF16_PATCH, patch: int = 14,
               ld: int | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 [B,H,W,3] frames -> resized / normalized / patchified tensor (see GVL_LAYOUT_*)."""
    _need_cuda(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_contiguous():
        raise RuntimeError("frames must be a contiguous uint8 [B,H,W,3] tensor")
    B, H, W, _ = frames.shape
    dev = frames.device
    if layout == LAYOUT_BF16_PATCH:
        gh, gw = out_h // patch, out_w // patch
        ld = ld or (3 * patch * patch + 7) // 8 * 8
        shape, dtype = (B * gh * gw, ld), torch.bfloat16
    else:
        shape = (B, 3, out_h, out_w)
        dtype = {LAYOUT_U8_CHW: torch.uint8, LAYOUT_F32_CHW: torch.float32, LAYOUT_BF16_CHW: torch.bfloat16}[layout]
        ld = 0
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=dev)
    elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
        raise RuntimeError(f"preprocess: `out` must be contiguous {dtype} {shape}")
    sub, div = fused_sub_div(image_mean, image_std)
    if H < out_h or W < out_w or _aten_taps_table(W, out_w, resample)[1] > 24:
        # an axis is up-scaled (crops and small images through `encode_image`), or the horizontal window is wider than
        # the tuned K1 kernels hold in registers (24 taps: beyond a ~11x bilinear / ~5.5x bicubic down-scale): the
        # general two-pass kernel runs ATen's own tables for any geometry (still bit-exact)
        return _preprocess_any_geometry(frames, out_h, out_w, resample, sub, div, layout, patch, ld, out)
    _lib.check(_lib.lib().gvl_preprocess_u8(
        frames.data_ptr(), B, H, W, out_h, out_w, resample, sub.ctypes.data_as(_lib.c_float_p),
        div.ctypes.data_as(_lib.c_float_p), out.data_ptr(), layout, patch, ld, _stream()), "gvl_preprocess_u8")
    return out


@functools.lru_cache(maxsize=1024)
def _aten_taps_table(in_size: int, out_size: int, resample: int):
    """ATen's int16 tap table of one axis in the layout gvl_resize_two_pass_u8 reads: int32 [xmin | count | k[out, taps]]."""
    xmin, xsize, w, prec = resize_taps(in_size, out_size, resample, max_taps=1024)
    tab = np.concatenate([xmin.astype(np.int32), xsize.astype(np.int32), w.astype(np.int32).reshape(-1)])
    tab.setflags(write=False)
    return tab, int(w.shape[1]), int(prec)


def _preprocess_any_geometry(frames, out_h, out_w, resample, sub, div, layout, patch, ld, out):
    B, H, W, _ = frames.shape
    dev = frames.device
    th, kh, ph = _aten_taps_table(W, out_w, resample)
    tv, kv, pv = _aten_taps_table(H, out_h, resample)
    # the batch is one tall image: frame b is the crop (0, b*H, W, H)
    desc = np.zeros((B, 10), np.int32)
    for b in range(B):
        desc[b] = (0, b * H, W, H, out_w, out_h, kh, kv, 0, th.size)
    tabs = torch.from_numpy(np.concatenate([th, tv])).to(dev)
    dptr = desc.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    scratch = torch.empty(int(_lib.lib().gvl_region_scratch_bytes(B, dptr)) + 256, dtype=torch.uint8, device=dev)
    # HF's fused rescale + normalise `(float(u8) - sub[c]) / div[c]`, evaluated once per byte value in fp32
    lut32 = (np.arange(256, dtype=np.float32)[None, :] - sub[:, None]) / div[:, None]
    lut_f32 = torch.from_numpy(np.ascontiguousarray(lut32, dtype=np.float32)).to(dev)
    lut_bf16 = lut_f32.to(torch.bfloat16)
    patches = resized = f32 = None
    if layout == LAYOUT_BF16_PATCH:
        patches = out
    elif layout == LAYOUT_F32_CHW:
        f32 = out
    elif layout == LAYOUT_BF16_CHW:
        f32 = torch.empty(tuple(out.shape), dtype=torch.float32, device=dev)
    else:
        resized = torch.empty((B, out_h, out_w, 3), dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().gvl_resize_two_pass_u8(
        frames.data_ptr(), B * H, W, B, dptr, tabs.data_ptr(), tabs.numel(), ph, pv, lut_bf16.data_ptr(), lut_f32.data_ptr(),
        out_h, out_w, patch, ld if layout == LAYOUT_BF16_PATCH else 8 * ((3 * patch * patch + 7) // 8), _ptr(patches),
        _ptr(resized), _ptr(f32), scratch.data_ptr(), scratch.numel(), _stream()), "gvl_resize_two_pass_u8")
    if layout == LAYOUT_BF16_CHW:
        out.copy_(f32)
    elif layout == LAYOUT_U8_CHW:
        out.copy_(resized.permute(0, 3, 1, 2))
    return out


PRE_PATH_AUTO, PRE_PATH_PLANAR, PRE_PATH_V1 = 0, 1, 2


@contextlib.contextmanager
def preprocess_path(path: int):
    """Pin `preprocess` to one of its (bit-identical) kernels for tests and A/B timing; AUTO on exit."""
    _lib.check(_lib.lib().gvl_preprocess_path(path), "gvl_preprocess_path")
    try:
        yield
    finally:
        _lib.lib().gvl_preprocess_path(PRE_PATH_AUTO)


@_on_tensor_device
def preprocess_crop(frames: torch.Tensor, out_h: int, out_w: int, crop_y0: int, crop_x0: int, crop_h: int, crop_w: int,
                    resample: int = BILINEAR, image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5),
                    layout: int = LAYOUT_BF16_CHW, out: torch.Tensor | None = None,
                    band: tuple[int, int] | None = None) -> torch.Tensor:
    """uint8 [B,H,W,3] frames -> the crop window of the resized (out_h x out_w) image, normalized, CHW.

    band = (x0, full_width): `frames` holds only the source columns [x0, x0 + frames.shape[2]) of full_width-wide
    frames (see `crop_source_band` / `copy_band_h2d`); the result is bit-identical to the whole-frame call."""
    _need_cuda(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_contiguous():
        raise RuntimeError("frames must be a contiguous uint8 [B,H,W,3] tensor")
    B, H, W, _ = frames.shape
    dtype = {LAYOUT_U8_CHW: torch.uint8, LAYOUT_F32_CHW: torch.float32, LAYOUT_BF16_CHW: torch.bfloat16}[layout]
    shape = (B, 3, crop_h, crop_w)
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=frames.device)
    elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
        raise RuntimeError(f"preprocess_crop: `out` must be contiguous {dtype} {shape}")
    sub, div = fused_sub_div(image_mean, image_std)
    if band is None:
        _lib.check(_lib.lib().gvl_preprocess_u8_crop(
            frames.data_ptr(), B, H, W, out_h, out_w, crop_y0, crop_x0, crop_h, crop_w, resample,
            sub.ctypes.data_as(_lib.c_float_p), div.ctypes.data_as(_lib.c_float_p), out.data_ptr(), layout, _stream()),
            "gvl_preprocess_u8_crop")
    else:
        x0, full_w = int(band[0]), int(band[1])
        _lib.check(_lib.lib().gvl_preprocess_u8_crop_band(
            frames.data_ptr(), B, H, full_w, x0, W, out_h, out_w, crop_y0, crop_x0, crop_h, crop_w, resample,
            sub.ctypes.data_as(_lib.c_float_p), div.ctypes.data_as(_lib.c_float_p), out.data_ptr(), layout, _stream()),
            "gvl_preprocess_u8_crop_band")
    return out


def crop_source_band(W: int, out_w: int, crop_x0: int, crop_w: int, resample: int = BILINEAR, align: int = 16) -> tuple[int, int]:
    """(x0, width) of the source-column band the crop window [crop_x0, crop_x0 + crop_w) of a W -> out_w resize reads,
    widened to multiples of `align` pixels (host only: ATen's tap table through gvl_resize_taps)."""
    xmin, xsize, _, _ = resize_taps(W, out_w, resample)
    lo = int(xmin[crop_x0])
    hi = int(xmin[crop_x0 + crop_w - 1] + xsize[crop_x0 + crop_w - 1])
    x0 = lo // align * align
    x1 = min(W, -(-hi // align) * align)
    return x0, x1 - x0


def copy_band_h2d(dst: torch.Tensor, src_host: torch.Tensor, x0: int) -> torch.Tensor:
    """dst (device uint8 [n,H,bw,3], contiguous) <- src_host[:, :, x0:x0+bw, :] of a contiguous host uint8 [n,H,W,3]
    tensor (pinned for an asynchronous copy), as one strided cudaMemcpy2DAsync on dst's current stream."""
    _need_cuda(dst)
    if src_host.is_cuda or not src_host.is_contiguous() or not dst.is_contiguous() or src_host.dtype != torch.uint8 \
            or dst.dtype != torch.uint8:
        raise RuntimeError("copy_band_h2d: contiguous uint8 host source and device destination required")
    n, H, W, C = src_host.shape
    bw = dst.shape[2]
    if tuple(dst.shape) != (n, H, bw, C) or x0 < 0 or x0 + bw > W:
        raise RuntimeError(f"copy_band_h2d: destination {tuple(dst.shape)} is not a band of {tuple(src_host.shape)} at {x0}")
    with torch.cuda.device(dst.device):
        _lib.check(_lib.lib().gvl_copy_band_h2d(dst.data_ptr(), src_host.data_ptr(), n * H, W * C, x0 * C, bw * C,
                                                torch.cuda.current_stream(dst.device).cuda_stream), "gvl_copy_band_h2d")
    return dst


@_on_tensor_device
def patchify_tubelet(pixel_values: torch.Tensor, frames: int, patch: int = 16, tubelet: int = 2,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """bf16 pixel_values [clips*frames,3,H,W] -> bf16 tubelet rows [clips*(frames/tubelet)*gh*gw, 3*tubelet*p*p]."""
    _need_cuda(pixel_values)
    if pixel_values.dtype != torch.bfloat16 or pixel_values.dim() != 4 or pixel_values.shape[1] != 3 or \
            not pixel_values.is_contiguous() or pixel_values.shape[0] % frames:
        raise RuntimeError("patchify_tubelet: pixel_values must be contiguous bf16 [clips*frames,3,H,W]")
    n, _, H, W = pixel_values.shape
    clips = n // frames
    shape = (clips * (frames // tubelet) * (H // patch) * (W // patch), 3 * tubelet * patch * patch)
    if out is None:
        out = torch.empty(shape, dtype=torch.bfloat16, device=pixel_values.device)
    _lib.check(_lib.lib().gvl_patchify_tubelet_bf16(pixel_values.data_ptr(), clips, frames, H, W, patch, tubelet,
                                                    out.data_ptr(), _stream()), "gvl_patchify_tubelet_bf16")
    return out


@_on_tensor_device
def mean_tokens(x: torch.Tensor, B: int, T: int, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """bf16 [B*T, D] -> [B, D] mean over the T tokens of each item."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.shape[0] != B * T:
        raise RuntimeError("mean_tokens: x must be contiguous bf16 [B*T, D]")
    D = x.shape[1]
    out = torch.empty((B, D), dtype=out_dtype, device=x.device)
    _lib.check(_lib.lib().gvl_mean_tokens_bf16(x.data_ptr(), B, T, D, out.data_ptr(),
                                               1 if out_dtype == torch.float32 else 0, _stream()), "gvl_mean_tokens_bf16")
    return out


@_on_tensor_device
def videomae_forward(pack: VideoMAEPack, patches: torch.Tensor, workspace: torch.Tensor | None = None,
                     out_dtype: torch.dtype = torch.float32, return_tokens: bool = False):
    """bf16 tubelet patches [B*T, patch_k] -> mean-pooled clip embeddings [B, D] (fp32 like the reference)."""
    _need_cuda(patches)
    spec = pack.spec
    if patches.dtype != torch.bfloat16 or patches.shape[1] != spec.patch_ld or patches.shape[0] % spec.tokens:
        raise RuntimeError("videomae_forward: patches must be bf16 [B*T, patch_k]")
    B = patches.shape[0] // spec.tokens
    need = pack.workspace_bytes(B)
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=patches.device)
    pooled = torch.empty((B, spec.hidden), dtype=out_dtype, device=patches.device)
    tokens = torch.empty((B * spec.tokens, spec.hidden), dtype=torch.bfloat16, device=patches.device) if return_tokens else None
    _lib.check(_lib.lib().gvl_videomae_forward(ctypes.byref(pack.struct), patches.data_ptr(), B, workspace.data_ptr(),
                                               workspace.numel(), pooled.data_ptr(),
                                               1 if out_dtype == torch.float32 else 0, _ptr(tokens), _stream()),
               "gvl_videomae_forward")
    return (pooled, tokens) if return_tokens else pooled


@_on_tensor_device
def patchify(pixel_values: torch.Tensor, patch: int = 14, ld: int | None = None) -> torch.Tensor:
    """fp32 pixel_values [B,3,H,W] -> bf16 patch rows [B*gh*gw, ld] (the get_image_features seam)."""
    _need_cuda(pixel_values)
    if pixel_values.dtype != torch.float32 or pixel_values.dim() != 4 or pixel_values.shape[1] != 3:
        raise RuntimeError("patchify: pixel_values must be fp32 [B,3,H,W]")
    pixel_values = pixel_values.contiguous()
    B, _, H, W = pixel_values.shape
    ld = ld or (3 * patch * patch + 7) // 8 * 8
    out = torch.empty((B * (H // patch) * (W // patch), ld), dtype=torch.bfloat16, device=pixel_values.device)
    _lib.check(_lib.lib().gvl_patchify_f32(pixel_values.data_ptr(), B, H, W, patch, ld, out.data_ptr(), _stream()),
               "gvl_patchify_f32")
    return out


def gemm_stats_slots(N: int) -> int:
    """Slabs per row a GEMM with N output columns writes into `stats_out` (host only)."""
    return int(_lib.lib().gvl_gemm_stats_slots(int(N)))


@_on_tensor_device
def gemm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, residual: torch.Tensor | None = None,
         res_row_mod: int = 0, act: int = ACT_NONE, out: torch.Tensor | None = None,
         out_dtype: torch.dtype = torch.bfloat16, stats_out: torch.Tensor | None = None,
         ln_stats: torch.Tensor | None = None, ln_c1: torch.Tensor | None = None, ln_dim: int = 0,
         ln_eps: float = 0.0) -> torch.Tensor:
    """act(a @ w.T + bias) + residual; a [M,K] bf16, w [N,K] bf16, bias fp32 [N], residual bf16.
    stats_out fp32 [M, gemm_stats_slots(N), 2]: also write the partial (sum, sum of squares) of every stored row.
    ln_stats (+ ln_c1 fp32 [N], ln_dim, ln_eps): normalise the rows of `a` in the epilogue (see gvl_gemm_fusion)."""
    _need_cuda(a, w, bias, residual, out)
    M, K = a.shape
    N, K2 = w.shape
    if K != K2 or a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise RuntimeError("gemm: a [M,K] and w [N,K] must be bf16 with matching K")
    if a.stride(1) != 1 or w.stride(1) != 1:
        raise RuntimeError("gemm: operands must be K-contiguous")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N):
        raise RuntimeError("gemm: bias must be fp32 [N]")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    if out.dtype not in (torch.bfloat16, torch.float32) or out.stride(1) != 1:
        raise RuntimeError("gemm: out must be bf16 or fp32, row-major")
    ldr = 0
    if residual is not None:
        if residual.dtype != torch.bfloat16 or residual.stride(1) != 1:
            raise RuntimeError("gemm: residual must be bf16, row-major")
        ldr = residual.stride(0)
    fusion = None
    if stats_out is not None or ln_stats is not None:
        _need_cuda(stats_out, ln_stats, ln_c1)
        fusion = _lib.GemmFusion()
        fusion.stats_out = _ptr(stats_out)
        fusion.ln_stats = _ptr(ln_stats)
        fusion.ln_slots = 0 if ln_stats is None else int(ln_stats.shape[1])
        fusion.ln_dim = int(ln_dim)
        fusion.ln_c1 = _ptr(ln_c1)
        fusion.ln_eps = float(ln_eps)
    _lib.check(_lib.lib().gvl_gemm_bf16_fused(
        a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), _ptr(residual), ldr, res_row_mod,
        out.data_ptr(), out.stride(0), 1 if out.dtype == torch.float32 else 0, M, N, K, act,
        ctypes.byref(fusion) if fusion is not None else None, _stream()), "gvl_gemm_bf16_fused")
    return out


@_on_tensor_device
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out: torch.Tensor | None = None) -> torch.Tensor:
    _need_cuda(x, gamma, beta, out)
    if x.dtype != torch.bfloat16 or x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError("layernorm: x must be bf16 [rows, D], row-major")
    rows, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().gvl_layernorm_bf16(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                             out.data_ptr(), out.stride(0), rows, D, float(eps), _stream()),
               "gvl_layernorm_bf16")
    return out


@_on_tensor_device
def attention(qkv: torch.Tensor, B: int, T: int, H: int, hd: int, scale: float | None = None,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """qkv bf16 [B*T, 3*H*hd] -> bf16 [B*T, H*hd]."""
    _need_cuda(qkv, out)
    if qkv.dtype != torch.bfloat16 or tuple(qkv.shape) != (B * T, 3 * H * hd) or not qkv.is_contiguous():
        raise RuntimeError("attention: qkv must be contiguous bf16 [B*T, 3*H*hd]")
    if out is None:
        out = torch.empty((B * T, H * hd), dtype=torch.bfloat16, device=qkv.device)
    scale = float(hd) ** -0.5 if scale is None else float(scale)
    _lib.check(_lib.lib().gvl_attention_bf16(qkv.data_ptr(), out.data_ptr(), B, T, H, hd, scale, _stream()),
               "gvl_attention_bf16")
    return out


@_on_tensor_device
def attention_varlen(qkv: torch.Tensor, item_tokens, H: int, hd: int, scale: float | None = None) -> torch.Tensor:
    """Self-attention over a ragged batch in ONE launch: qkv bf16 [sum T_i, 3*H*hd], items of `item_tokens[i]` tokens
    back to back -> bf16 [sum T_i, H*hd]; every item attends to itself only."""
    _need_cuda(qkv)
    toks = np.ascontiguousarray(np.asarray(item_tokens, np.int32))
    M = int(toks.sum())
    if qkv.dtype != torch.bfloat16 or tuple(qkv.shape) != (M, 3 * H * hd) or not qkv.is_contiguous():
        raise RuntimeError("attention_varlen: qkv must be contiguous bf16 [sum T_i, 3*H*hd]")
    tp = toks.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    n = ctypes.c_int(0)
    _lib.check(_lib.lib().gvl_attention_varlen_tiles(len(toks), tp, None, ctypes.byref(n)), "gvl_attention_varlen_tiles")
    table = np.zeros((n.value, 4), np.int32)
    _lib.check(_lib.lib().gvl_attention_varlen_tiles(len(toks), tp, table.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                                     ctypes.byref(n)), "gvl_attention_varlen_tiles")
    tiles = torch.from_numpy(table).to(qkv.device)
    out = torch.empty((M, H * hd), dtype=torch.bfloat16, device=qkv.device)
    scale = float(hd) ** -0.5 if scale is None else float(scale)
    _lib.check(_lib.lib().gvl_attention_varlen_bf16(qkv.data_ptr(), out.data_ptr(), M, tiles.data_ptr(), n.value,
                                                    float((toks.astype(np.float64) ** 2).sum()), H, hd, scale, _stream()),
               "gvl_attention_varlen_bf16")
    return out


@_on_tensor_device
def probe_attention(q: torch.Tensor, kv: torch.Tensor, B: int, T: int, H: int, hd: int) -> torch.Tensor:
    """q fp32 [H*hd] (pre-scaled), kv bf16 [B*T, 2*H*hd] -> bf16 [B, H*hd]."""
    _need_cuda(q, kv)
    if q.dtype != torch.float32 or kv.dtype != torch.bfloat16 or not kv.is_contiguous():
        raise RuntimeError("probe_attention: q fp32, kv contiguous bf16")
    out = torch.empty((B, H * hd), dtype=torch.bfloat16, device=kv.device)
    _lib.check(_lib.lib().gvl_probe_attention_bf16(q.data_ptr(), kv.data_ptr(), out.data_ptr(), B, T, H, hd, _stream()),
               "gvl_probe_attention_bf16")
    return out


@_on_tensor_device
def siglip_forward(pack: "SiglipPack | SiglipGridView", patches: torch.Tensor, workspace: torch.Tensor | None = None,
                   return_tokens: bool = False):
    """bf16 patches [B*T, patch_ld] -> pooled bf16 [B, D] (and post-LN tokens [B*T, D] if asked).  `pack` may be a
    `siglip_grid_view` (T = gh*gw tokens per item, re-sampled position table)."""
    _need_cuda(patches)
    spec, T = pack.spec, pack.tokens
    if patches.dtype != torch.bfloat16 or patches.shape[1] != spec.patch_ld or patches.shape[0] % T:
        raise RuntimeError("siglip_forward: patches must be bf16 [B*T, patch_ld]")
    B = patches.shape[0] // T
    need = pack.workspace_bytes(B)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=patches.device)
    pooled = torch.empty((B, spec.hidden), dtype=torch.bfloat16, device=patches.device)
    tokens = torch.empty((B * T, spec.hidden), dtype=torch.bfloat16, device=patches.device) if return_tokens else None
    _lib.check(_lib.lib().gvl_siglip_forward(ctypes.byref(pack.struct), patches.data_ptr(), B, workspace.data_ptr(),
                                             workspace.numel(), pooled.data_ptr(), _ptr(tokens), _stream()),
               "gvl_siglip_forward")
    return (pooled, tokens) if return_tokens else pooled


@_on_tensor_device
def project(pp: ProjectorPack, x: torch.Tensor, out_dtype: torch.dtype = torch.float32,
            hidden: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """MultiModalProjector: bf16 [M, enc] -> [M, llm] (fp32 like the reference, or bf16 for the index)."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or x.dim() != 2 or not x.is_contiguous() or x.shape[1] != pp.encoder_dim:
        raise RuntimeError("project: x must be contiguous bf16 [M, encoder_dim]")
    M = x.shape[0]
    if hidden is None:
        hidden = torch.empty((M, pp.llm_dim), dtype=torch.bfloat16, device=x.device)
    if out is None:
        out = torch.empty((M, pp.llm_dim), dtype=out_dtype, device=x.device)
    elif tuple(out.shape) != (M, pp.llm_dim) or not out.is_contiguous() or out.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("project: `out` must be contiguous fp32/bf16 [M, llm_dim]")
    out_dtype = out.dtype
    _lib.check(_lib.lib().gvl_project(x.data_ptr(), M, pp.encoder_dim, pp.llm_dim, pp.w1.data_ptr(), pp.b1.data_ptr(),
                                      pp.w2.data_ptr(), pp.b2.data_ptr(), hidden.data_ptr(), out.data_ptr(),
                                      1 if out_dtype == torch.float32 else 0, _stream()), "gvl_project")
    return out


TOPK_AUTO, TOPK_SCAN, TOPK_TENSOR = 0, 1, 2


@_on_tensor_device
def row_inv_norm(rows: torch.Tensor, eps: float = 1e-12, out: torch.Tensor | None = None) -> torch.Tensor:
    """1 / max(|row|, eps) of bf16 rows [N,D] -> fp32 [N] (cache it next to an index that does not change)."""
    _need_cuda(rows, out)
    if rows.dtype != torch.bfloat16 or not rows.is_contiguous():
        raise RuntimeError("row_inv_norm: rows must be contiguous bf16")
    N, D = rows.shape
    if out is None:
        out = torch.empty((N,), dtype=torch.float32, device=rows.device)
    _lib.check(_lib.lib().gvl_row_inv_norm(rows.data_ptr(), N, D, float(eps), out.data_ptr(), _stream()),
               "gvl_row_inv_norm")
    return out


@_on_tensor_device
def topk_cosine(index: torch.Tensor, queries: torch.Tensor, k: int, eps: float = 1e-12,
                row_lo: torch.Tensor | None = None, row_hi: torch.Tensor | None = None,
                span: tuple[int, int] | None = None, mode: int = TOPK_AUTO, inv_norm: torch.Tensor | None = None):
    """index bf16 [N,D], queries bf16 [Q,D] (or both fp32) -> (scores fp32 [Q,k], idx int32 [Q,k]), score desc / idx asc.
    row_lo / row_hi (int32 [Q], device): query q only ranks rows [row_lo[q], row_hi[q]); `span` = (min lo, max hi),
    the only rows that are scored (computed here from the tensors if omitted — one host sync).  Slots beyond the
    number of eligible rows hold idx -1 / score -inf.  inv_norm: cached `row_inv_norm(index)` (TENSOR path)."""
    _need_cuda(index, queries, row_lo, row_hi, inv_norm)
    if index.dtype == torch.float32 and queries.dtype == torch.float32:
        # fp32 rows stay fp32 (reference: `.float()` before F.cosine_similarity); scan path, whole index only
        if row_lo is not None or not index.is_contiguous() or not queries.is_contiguous():
            raise RuntimeError("topk_cosine: the fp32 path takes contiguous tensors and no row ranges")
        N, D = index.shape
        Q = queries.shape[0]
        scratch = torch.empty((int(_lib.lib().gvl_topk_scratch_floats(N, Q)),), dtype=torch.float32, device=index.device)
        scores = torch.empty((Q, k), dtype=torch.float32, device=index.device)
        idx = torch.empty((Q, k), dtype=torch.int32, device=index.device)
        _lib.check(_lib.lib().gvl_topk_cosine_f32(index.data_ptr(), N, D, queries.data_ptr(), Q, k, float(eps),
                                                  scratch.data_ptr(), scores.data_ptr(), idx.data_ptr(), _stream()),
                   "gvl_topk_cosine_f32")
        return scores, idx
    if index.dtype != torch.bfloat16 or queries.dtype != torch.bfloat16 or not index.is_contiguous() or not queries.is_contiguous():
        raise RuntimeError("topk_cosine: index and queries must be contiguous bf16")
    N, D = index.shape
    Q = queries.shape[0]
    if (row_lo is None) != (row_hi is None):
        raise RuntimeError("topk_cosine: row_lo and row_hi come together")
    lo_hi = (0, N)
    if row_lo is not None:
        if row_lo.dtype != torch.int32 or row_hi.dtype != torch.int32 or row_lo.numel() != Q or row_hi.numel() != Q:
            raise RuntimeError("topk_cosine: row_lo / row_hi must be int32 [Q]")
        lo_hi = span if span is not None else (max(0, int(row_lo.min())), min(N, int(row_hi.max())))
        lo_hi = (lo_hi[0], max(lo_hi[0], lo_hi[1]))
    if inv_norm is not None and (inv_norm.dtype != torch.float32 or inv_norm.numel() < N):
        raise RuntimeError("topk_cosine: inv_norm must be fp32 [N]")
    scratch = torch.empty((int(_lib.lib().gvl_topk_scratch_floats(N, Q)),), dtype=torch.float32, device=index.device)
    scores = torch.empty((Q, k), dtype=torch.float32, device=index.device)
    idx = torch.empty((Q, k), dtype=torch.int32, device=index.device)
    _lib.check(_lib.lib().gvl_topk_cosine_ex(
        index.data_ptr(), N, D, queries.data_ptr(), Q, k, float(eps), _ptr(row_lo), _ptr(row_hi), int(lo_hi[0]),
        int(lo_hi[1]), int(mode), _ptr(inv_norm), scratch.data_ptr(), scores.data_ptr(), idx.data_ptr(), _stream()),
        "gvl_topk_cosine_ex")
    return scores, idx


# ------------------------------------------------------------------------------------------------ masked regions (K9)
@functools.lru_cache(maxsize=4096)
def pil_bicubic_taps(in_size: int, out_size: int):
    """Host-only: Pillow's 8-bit bicubic coefficient table of one axis as int32 [xmin[out] | count[out] | k[out, ksize]]
    (the layout gvl_region_patches_pil_u8 reads) and ksize."""
    ks = ctypes.c_int(0)
    _lib.check(_lib.lib().gvl_pil_bicubic_taps(in_size, out_size, 0, None, None, None, ctypes.byref(ks)), "gvl_pil_bicubic_taps")
    k = ks.value
    tab = np.zeros(out_size * (2 + k), np.int32)
    p = tab.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    i32 = ctypes.sizeof(ctypes.c_int32)
    at = lambda off: ctypes.cast(ctypes.addressof(p.contents) + off * i32, ctypes.POINTER(ctypes.c_int32))  # noqa: E731
    _lib.check(_lib.lib().gvl_pil_bicubic_taps(in_size, out_size, k, at(0), at(out_size), at(2 * out_size), None),
               "gvl_pil_bicubic_taps")
    tab.setflags(write=False)
    return tab, k


def region_lut(image_mean, image_std, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """[3, 256] table of `RegionExtractor.prepare_region_tensor`'s normalisation (reference :357-365) evaluated with the
    reference's own three fp32 CPU operations on every possible uint8 value, then cast like `.to(device, dtype)` (:539)."""
    t = torch.arange(256, dtype=torch.float32).view(1, 256).expand(3, 256).contiguous()
    t = t / 255.0
    t = (t - torch.tensor(list(image_mean)).view(3, 1)) / torch.tensor(list(image_std)).view(3, 1)
    return t.to(dtype)


@_on_tensor_device
def region_patches(frame: torch.Tensor, boxes, sizes, canvas_hw, lut: torch.Tensor | None, patch: int = 14,
                   ld: int | None = None, want_patches: bool = True, want_u8: bool = False, out: torch.Tensor | None = None):
    """Crop + Pillow-bicubic resize + normalise + zero-pad + im2col of R regions of ONE frame.

    frame: uint8 [H,W,3] on the device; boxes: R x (x1, y1, x2, y2) (`frame[y1:y2, x1:x2]`); sizes: R x (out_h, out_w);
    canvas_hw: (canvas_h, canvas_w) multiples of `patch`, >= every size; None = the RAGGED form: every region on its own
    canvas (its size, a multiple of `patch`), patch rows back to back.  Returns (patches bf16 [R*gh*gw, ld] — ragged:
    [sum gh_r*gw_r, ld] — or None, resized uint8 [R, canvas_h, canvas_w, 3] or None).  `out`: a contiguous bf16 [R*gh*gw, ld] slice to write the patch
    rows into (a ragged batch is assembled from several calls)."""
    _need_cuda(frame, lut)
    if frame.dtype != torch.uint8 or frame.dim() != 3 or frame.shape[2] != 3 or not frame.is_contiguous():
        raise RuntimeError("region_patches: frame must be contiguous uint8 [H,W,3]")
    R = len(boxes)
    if R == 0 or len(sizes) != R:
        raise RuntimeError("region_patches: need one (out_h, out_w) per box")
    H, W = int(frame.shape[0]), int(frame.shape[1])
    ragged = canvas_hw is None
    canvas_h, canvas_w = (0, 0) if ragged else (int(canvas_hw[0]), int(canvas_hw[1]))
    if ragged and want_u8:
        raise RuntimeError("region_patches: the ragged form writes patch rows only")
    ld = ld or (3 * patch * patch + 7) // 8 * 8
    desc = np.zeros((R, 10), np.int32)
    tabs, off = [], 0
    for r, ((x1, y1, x2, y2), (oh, ow)) in enumerate(zip(boxes, sizes)):
        cw, ch = int(x2) - int(x1), int(y2) - int(y1)
        if cw <= 0 or ch <= 0:
            raise ValueError(f"region {r}: empty box {(x1, y1, x2, y2)}")  # PIL raises on a zero-size crop too
        th, kh = pil_bicubic_taps(cw, int(ow))
        tv, kv = pil_bicubic_taps(ch, int(oh))
        desc[r] = (x1, y1, cw, ch, ow, oh, kh, kv, off, off + th.size)
        tabs += [th, tv]
        off += th.size + tv.size
    tabs_dev = torch.from_numpy(np.concatenate(tabs)).to(frame.device)
    dptr = desc.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    scratch = torch.empty(int(_lib.lib().gvl_region_scratch_bytes(R, dptr)) + 256, dtype=torch.uint8, device=frame.device)
    gh, gw = canvas_h // patch, canvas_w // patch
    rows = sum((int(oh) // patch) * (int(ow) // patch) for oh, ow in sizes) if ragged else R * gh * gw
    patches = None
    if want_patches:
        patches = out if out is not None else torch.empty((rows, ld), dtype=torch.bfloat16, device=frame.device)
        if patches.dtype != torch.bfloat16 or tuple(patches.shape) != (rows, ld) or not patches.is_contiguous():
            raise RuntimeError(f"region_patches: `out` must be contiguous bf16 [{rows}, {ld}]")
    resized = torch.empty((R, canvas_h, canvas_w, 3), dtype=torch.uint8, device=frame.device) if want_u8 else None
    if want_patches and (lut is None or lut.dtype != torch.bfloat16 or tuple(lut.shape) != (3, 256) or not lut.is_contiguous()):
        raise RuntimeError("region_patches: lut must be contiguous bf16 [3, 256]")
    _lib.check(_lib.lib().gvl_region_patches_pil_u8(frame.data_ptr(), H, W, R, dptr, tabs_dev.data_ptr(), tabs_dev.numel(),
                                                    _ptr(lut), canvas_h, canvas_w, patch, ld, _ptr(patches), _ptr(resized),
                                                    scratch.data_ptr(), scratch.numel(), _stream()),
               "gvl_region_patches_pil_u8")
    return patches, resized


@_on_tensor_device
def interpolate_pos(pos: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """Learned position table bf16 [g*g, D] -> bf16 [gh*gw, D] (HF `interpolate_pos_encoding`: bicubic, align_corners
    False).  The g x g grid returns the table itself, as HF does (HF:models/siglip/modeling_siglip.py:152-153)."""
    _need_cuda(pos)
    n, D = pos.shape
    g = int(round(n ** 0.5))
    if pos.dtype != torch.bfloat16 or not pos.is_contiguous() or g * g != n:
        raise RuntimeError("interpolate_pos: pos must be contiguous bf16 [g*g, D]")
    if gh == g and gw == g:
        return pos
    out = torch.empty((gh * gw, D), dtype=torch.bfloat16, device=pos.device)
    _lib.check(_lib.lib().gvl_pos_interp_bicubic_bf16(pos.data_ptr(), g, D, gh, gw, out.data_ptr(), _stream()),
               "gvl_pos_interp_bicubic_bf16")
    return out


@_on_tensor_device
def max_tokens(x: torch.Tensor, B: int, T: int, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """bf16 [B*T, D] -> [B, D] max over the T tokens of each item."""
    _need_cuda(x)
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.shape[0] != B * T:
        raise RuntimeError("max_tokens: x must be contiguous bf16 [B*T, D]")
    D = x.shape[1]
    out = torch.empty((B, D), dtype=out_dtype, device=x.device)
    _lib.check(_lib.lib().gvl_max_tokens_bf16(x.data_ptr(), B, T, D, out.data_ptr(),
                                              1 if out_dtype == torch.float32 else 0, _stream()), "gvl_max_tokens_bf16")
    return out


def siglip_ragged_workspace_bytes(pack: SiglipPack, M_total: int, B_total: int) -> int:
    return int(_lib.lib().gvl_siglip_ragged_workspace_bytes(ctypes.byref(pack.struct), int(M_total), int(B_total)))


@_on_tensor_device
def siglip_forward_ragged(pack: SiglipPack, patches: torch.Tensor, groups, workspace: torch.Tensor | None = None,
                          return_tokens: bool = False):
    """One tower pass over a ragged batch.  groups: [(B_g, gh_g, gw_g)] — B_g items on a gh_g x gw_g patch grid each,
    rows of `patches` (bf16 [sum B*gh*gw, patch_ld]) concatenated in that order.  Returns pooled bf16 [sum B, D] (MAP
    head) and, if asked, the post-layernorm tokens bf16 [sum B*gh*gw, D].  Row-wise kernels run once over all rows;
    each item's result is bit-identical to `siglip_forward` on its own grid view."""
    _need_cuda(patches)
    spec = pack.spec
    views = [siglip_grid_view(pack, gh, gw) for _, gh, gw in groups]
    arr = (_lib.RaggedGroup * len(groups))()
    M = Bt = 0
    for g, ((B, gh, gw), v) in enumerate(zip(groups, views)):
        arr[g].B, arr[g].T, arr[g].pos = int(B), gh * gw, v.pos_table.data_ptr()
        M += int(B) * gh * gw
        Bt += int(B)
    if patches.dtype != torch.bfloat16 or tuple(patches.shape) != (M, spec.patch_ld) or not patches.is_contiguous():
        raise RuntimeError(f"siglip_forward_ragged: patches must be contiguous bf16 [{M}, {spec.patch_ld}]")
    need = siglip_ragged_workspace_bytes(pack, M, Bt)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=patches.device)
    pooled = torch.empty((Bt, spec.hidden), dtype=torch.bfloat16, device=patches.device)
    tokens = torch.empty((M, spec.hidden), dtype=torch.bfloat16, device=patches.device) if return_tokens else None
    _lib.check(_lib.lib().gvl_siglip_forward_ragged(ctypes.byref(pack.struct), patches.data_ptr(), len(groups), arr,
                                                    workspace.data_ptr(), workspace.numel(), pooled.data_ptr(), _ptr(tokens),
                                                    _stream()), "gvl_siglip_forward_ragged")
    return (pooled, tokens) if return_tokens else pooled


def siglip_grid_view(pack: SiglipPack, gh: int, gw: int) -> "SiglipPack | SiglipGridView":
    """`pack` seen with a gh x gw patch grid: same weights, T = gh*gw, position table re-sampled once per grid and cached
    on the pack.  The checkpoint's own square grid is the pack itself."""
    g = pack.spec.grid
    if (gh, gw) == (g, g):
        return pack
    views = pack.__dict__.setdefault("_grid_views", {})
    if (gh, gw) not in views:
        if len(views) >= 64:  # bounded: variable-size inputs must not grow device memory without limit
            views.pop(next(iter(views)))
        views[(gh, gw)] = SiglipGridView(pack, gh, gw, interpolate_pos(pack.pos_table, gh, gw))
    return views[(gh, gw)]
