"""Per-kernel parity: every CUDA kernel, called through the C ABI (ops -> ctypes -> libgvl_sm100a.so),
against the oracle (integer work: bit-exact; floating point: stated tolerances vs fp32 math on the same
bf16 inputs)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gameplay_vision_llm_b200 import ops, synth  # noqa: E402
from oracle import preprocess_ref, siglip_ref  # noqa: E402

DEV = "cuda:0"


def _rel_err(out: torch.Tensor, ref: torch.Tensor):
    out, ref = out.double().cpu(), ref.double().cpu()
    err = (out - ref).abs()
    return err.max().item(), (err.max() / ref.abs().max().clamp_min(1e-30)).item(), err.mean().item()


# ------------------------------------------------------------------------------------------------ GEMM
GEMM_CASES = [
    # M, N, K, act, residual(0 none, 1 full, T = row modulo), out_fp32
    (128, 128, 64, 0, 0, False),
    (128, 256, 128, 0, 0, True),
    (300, 1152, 1152, 0, 1, False),      # out-proj: bias + residual, ragged M
    (1458, 3456, 1152, 0, 0, False),     # QKV, BN=192
    (520, 4304, 1152, 1, 0, False),      # fc1 + GELU-tanh, ragged N (4304 = 16*269)
    (260, 1152, 4304, 0, 1, False),      # fc2: ragged K (4304 = 67*64 + 16)
    (1458, 1152, 592, 0, 729, False),    # patch embedding + position embedding (row % 729)
    (64, 4096, 1152, 2, 0, False),       # projector fc1 + GELU-erf
    (64, 4096, 4096, 0, 0, True),        # projector fc2, fp32 out
    (5000, 2304, 1152, 0, 0, False),     # MAP-head K/V projection, many tiles per CTA? (40 x 12)
    (46656, 1152, 1152, 0, 1, False),    # full batch-64 out-proj: > 148 tiles, persistent loop + phases
]


@pytest.mark.parametrize("M,N,K,act,res,f32", GEMM_CASES)
def test_gemm(M, N, K, act, res, f32):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g)).to(torch.bfloat16).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.bfloat16).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    residual = None
    mod = 0
    if res == 1:
        residual = torch.randn(M, N, generator=g).to(torch.bfloat16).to(DEV)
    elif res > 1:
        residual = torch.randn(res, N, generator=g).to(torch.bfloat16).to(DEV)
        mod = res
    out = ops.gemm(a, w, bias, residual, res_row_mod=mod, act=act,
                   out_dtype=torch.float32 if f32 else torch.bfloat16)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().T + bias
    if act == 1:
        ref = siglip_ref.gelu_tanh(ref)
    elif act == 2:
        ref = siglip_ref.gelu_erf(ref)
    if residual is not None:
        rr = residual.float()
        ref = ref + (rr[torch.arange(M, device=DEV) % mod] if mod else rr)
    mx, rel, mean = _rel_err(out, ref)
    print(f"gemm M={M} N={N} K={K} act={act} res={res} f32={f32}: max_abs={mx:.3e} rel={rel:.3e} mean={mean:.3e}")
    assert torch.isfinite(out.float()).all()
    tol = 2e-5 if f32 else 6e-3  # fp32 out: accumulation-order noise only; bf16 out: one rounding (2^-8 relative)
    assert rel < tol, f"relative error {rel}"


def test_gemm_strided_views():
    """A and out as column slices of wider buffers (lda/ldo > logical width), in-place residual."""
    g = torch.Generator().manual_seed(5)
    big = torch.randn(400, 2304, generator=g).to(torch.bfloat16).to(DEV)
    a = big[:, 1152:]                                # lda = 2304
    w = (torch.randn(1152, 1152, generator=g) / 34).to(torch.bfloat16).to(DEV)
    x = torch.randn(400, 1152, generator=g).to(torch.bfloat16).to(DEV)
    ref = a.float() @ w.float().T + x.float()
    ops.gemm(a, w, None, residual=x, out=x)         # x += a @ w.T, in place
    torch.cuda.synchronize()
    _, rel, _ = _rel_err(x, ref)
    assert rel < 6e-3


@pytest.mark.parametrize("case", ["plain", "outlier_channels", "common_offset"])
def test_gemm_layernorm_fusion(case):
    """Producer GEMM writes partial row sums of what it stored; consumer GEMM on raw x with gamma-folded weights
    reproduces Linear(LayerNorm(x)) (include/gvl.h: gvl_gemm_fusion).  The folded statistics are E[x^2] - mean^2 from
    fp32 partial sums, so they lose precision only when a COMMON offset of all channels dwarfs their spread:
    `outlier_channels` (massive activations of +-200 in a few channels, as trained ViT residual streams have: they
    raise the variance itself, nothing cancels) and `common_offset` (every channel shifted by 20 standard deviations:
    mean^2 / var = 400, relative variance error ~400 x 2^-24) must both stay inside the same tolerance; an offset of
    hundreds of standard deviations would not, which is why EmbeddingPipeline(fold_ln=False) exists."""
    g = torch.Generator().manual_seed(11)
    M, D, N2 = 1000, 1152, 4304
    a = torch.randn(M, 592, generator=g).to(torch.bfloat16).to(DEV)
    w0 = (torch.randn(D, 592, generator=g) / 24).to(torch.bfloat16).to(DEV)
    res = torch.randn(M, D, generator=g) * 3 + 0.7
    if case == "outlier_channels":
        res[:, [5, 300, 777, 1100]] += torch.tensor([200.0, -150.0, 90.0, -220.0])
    elif case == "common_offset":
        res += 60.0  # 20 x the spread of 3
    res = res.to(torch.bfloat16).to(DEV)
    slots = ops.gemm_stats_slots(D)
    stats = torch.full((M, slots, 2), float("nan"), device=DEV)
    x = ops.gemm(a, w0, None, residual=res, stats_out=stats)  # x = a @ w0.T + res, bf16, + statistics
    torch.cuda.synchronize()
    xs = x.float()
    assert torch.allclose(stats[:, :, 0].sum(1), xs.sum(1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(stats[:, :, 1].sum(1), (xs * xs).sum(1), rtol=1e-5, atol=1e-2)
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
    beta = (0.05 * torch.randn(D, generator=g)).to(DEV)
    w1 = (torch.randn(N2, D, generator=g) / 34).to(torch.bfloat16).to(DEV)
    b1 = (0.02 * torch.randn(N2, generator=g)).to(DEV)
    wf = (w1.double() * gamma.double()[None]).to(torch.bfloat16)
    c1 = wf.double().sum(1).float()
    c2 = (b1.double() + w1.double() @ beta.double()).float()
    got = ops.gemm(x, wf, c2, act=1, ln_stats=stats, ln_c1=c1, ln_dim=D, ln_eps=1e-6)
    torch.cuda.synchronize()
    ref = siglip_ref.gelu_tanh(torch.nn.functional.layer_norm(xs, (D,), gamma, beta, 1e-6) @ w1.float().T + b1)
    mx, rel, mean = _rel_err(got, ref)
    print(f"gemm LN fusion [{case}]: max_abs={mx:.3e} rel={rel:.3e} mean={mean:.3e}")
    assert rel < 8e-3


# ------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,D", [(7, 1152), (1458, 1152), (33, 768), (5, 144), (9, 4096)])
def test_layernorm(rows, D):
    g = torch.Generator().manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 3 + 0.5).to(torch.bfloat16).to(DEV)
    gamma = (1 + 0.1 * torch.randn(D, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(D, generator=g)).to(DEV)
    out = ops.layernorm(x, gamma, beta, 1e-6)
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-6)
    mx, rel, _ = _rel_err(out, ref)
    print(f"layernorm {rows}x{D}: max_abs={mx:.3e} rel={rel:.3e}")
    assert rel < 5e-3


# ------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("B,T,H,hd", [(1, 16, 2, 72), (2, 729, 16, 72), (1, 200, 3, 64), (3, 129, 2, 72), (1, 1568, 12, 64)])
def test_attention(B, T, H, hd):
    g = torch.Generator().manual_seed(B * 100 + T)
    D = H * hd
    qkv = (torch.randn(B * T, 3 * D, generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    out = ops.attention(qkv, B, T, H, hd)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    mx, rel, mean = _rel_err(out, ref)
    print(f"attention B={B} T={T} H={H} hd={hd}: max_abs={mx:.3e} rel={rel:.3e} mean={mean:.3e}")
    assert torch.isfinite(out.float()).all()
    assert rel < 1.5e-2 and mean < 2e-3  # P is rounded to bf16 before PV (as HF eager/SDPA do)


@pytest.mark.parametrize("B,T,H,hd", [(8, 729, 16, 72), (40, 129, 8, 72), (24, 200, 8, 64), (330, 40, 1, 72)])
def test_attention_persistent_ctas_walk_several_items(B, T, H, hd):
    """The attention CTAs are persistent (two per SM): with more work items (query tile x head x image) than resident
    CTAs every CTA walks several items with its mbarrier phases, K/V ring and S-buffer parity running on — even block
    counts (12), odd ones (3: the buffer parity flips from item to item), a single block per item.  (a) against fp32
    softmax; (b) bit-identical to the same kernel launched with one item per CTA (the tuning hook of attention_sdb.cu)."""
    import ctypes

    from gameplay_vision_llm_b200 import _lib
    n_items = -(-T // 128) * H * B
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert n_items > 2 * sms, "case does not make the CTAs loop on this device"
    g = torch.Generator().manual_seed(B * 1000 + T)
    D = H * hd
    qkv = (torch.randn(B * T, 3 * D, generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    out = ops.attention(qkv, B, T, H, hd)
    torch.cuda.synchronize()
    ref = _attn_ref(qkv, B, T, H, hd)
    mx, rel, mean = _rel_err(out, ref)
    print(f"attention persistent B={B} T={T} H={H} hd={hd} ({n_items} items): max_abs={mx:.3e} rel={rel:.3e}")
    assert torch.isfinite(out.float()).all()
    assert rel < 1.5e-2 and mean < 2e-3
    lib = _lib.lib()
    lib.gvl_debug_set_attn_ctas_per_sm.argtypes = [ctypes.c_int]
    lib.gvl_debug_set_attn_ctas_per_sm.restype = None
    try:
        lib.gvl_debug_set_attn_ctas_per_sm(0)
        one = ops.attention(qkv, B, T, H, hd)
        torch.cuda.synchronize()
    finally:
        lib.gvl_debug_set_attn_ctas_per_sm(2)
    assert torch.equal(out, one)


@pytest.mark.parametrize("B,T,H,hd", [(1, 200, 2, 72), (3, 129, 2, 64), (1, 729, 16, 72)])
def test_attention_never_writes_past_the_sequence(B, T, H, hd):
    """The last query tile of an image holds rows beyond T (729 = 5 x 128 + 89): the staged read-out must not store them —
    in a batch they are the next image's first rows, after the last image they lie outside the buffer.  Guard rows in
    front of and behind the output keep their sentinel, and a batch equals its images attended one by one."""
    g = torch.Generator().manual_seed(7 * T + hd)
    D = H * hd
    qkv = (torch.randn(B * T, 3 * D, generator=g) * 1.5).to(torch.bfloat16).to(DEV)
    guard = 160  # more rows than a 128-row tile can overshoot
    big = torch.full((guard + B * T + guard, D), 123.0, dtype=torch.bfloat16, device=DEV)
    out = ops.attention(qkv, B, T, H, hd, out=big[guard:guard + B * T])
    torch.cuda.synchronize()
    assert torch.all(big[:guard] == 123.0) and torch.all(big[guard + B * T:] == 123.0)
    assert torch.isfinite(out.float()).all()
    for b in range(B):
        one = ops.attention(qkv[b * T:(b + 1) * T].contiguous(), 1, T, H, hd)
        assert torch.equal(out[b * T:(b + 1) * T], one), b


def _attn_ref(qkv, B, T, H, hd):
    D = H * hd
    q, k, v = qkv.float().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)


@pytest.mark.parametrize("hd", [72, 64])
def test_attention_wide_score_range_and_exact_path(hd):
    """The attention kernel keeps no running max: every row is exponentiated against one fixed reference (its score
    for key 0 + 95 in log2 units) and a row whose scores leave the representable window is detected by its row sum and
    recomputed exactly (attention_p2.cu).  (a) peaky rows — logits spread over +-60 nats, inside the window: fast path;
    (b) key 0 scoring ~136 nats below the row maximum and (c) ~136 nats above everything else: outside the window on
    the overflow side / far inside on the other — all three must match fp32 softmax."""
    B, T, H = 2, 300, 2
    D = H * hd
    g = torch.Generator().manual_seed(hd)
    base = (torch.randn(B * T, 3 * D, generator=g)).to(torch.bfloat16)
    # (a) scale q up: logits ~ N(0, 18^2) nats
    wide = base.clone().float()
    wide[:, :D] *= 18.0
    # (b) key 0 of every (image, head) far below: k_0 = -4 * ones, q = +4 * ones + noise  -> dot ~ -16 hd
    low = base.clone().float()
    low.view(B, T, 3, H, hd)[:, :, 0] = 4.0 + 0.1 * low.view(B, T, 3, H, hd)[:, :, 0]
    low.view(B, T, 3, H, hd)[:, 0, 1] = -4.0
    # (c) key 0 far above: k_0 = +4 * ones with the same positive queries
    high = low.clone()
    high.view(B, T, 3, H, hd)[:, 0, 1] = 4.0
    for name, x in (("wide", wide), ("key0 low (exact path)", low), ("key0 high", high)):
        qkv = x.to(torch.bfloat16).to(DEV)
        out = ops.attention(qkv, B, T, H, hd)
        torch.cuda.synchronize()
        ref = _attn_ref(qkv.cpu(), B, T, H, hd)
        mx, rel, mean = _rel_err(out, ref)
        print(f"attention hd={hd} {name}: max_abs={mx:.3e} rel={rel:.3e} mean={mean:.3e}")
        assert torch.isfinite(out.float()).all(), name
        assert rel < 1.5e-2 and mean < 2e-3, name


def test_probe_attention():
    B, T, H, hd = 3, 729, 16, 72
    D = H * hd
    g = torch.Generator().manual_seed(11)
    kv = torch.randn(B * T, 2 * D, generator=g).to(torch.bfloat16).to(DEV)
    q = (torch.randn(D, generator=g) * hd ** -0.5).to(DEV)
    out = ops.probe_attention(q, kv, B, T, H, hd)
    torch.cuda.synchronize()
    k, v = kv.float().view(B, T, 2, H, hd).permute(2, 0, 3, 1, 4)  # [B,H,T,hd]
    s = torch.einsum("hd,bhtd->bht", q.view(H, hd), k)
    ref = torch.einsum("bht,bhtd->bhd", torch.softmax(s, -1), v).reshape(B, D)
    _, rel, _ = _rel_err(out, ref)
    assert rel < 5e-3


# --------------------------------------------------------------------------------------- preprocessing
PRE_CASES = [(1080, 1920, 384, 384, 2), (1080, 1920, 384, 384, 3), (123, 211, 56, 56, 2), (270, 480, 96, 112, 3),
             (720, 1280, 384, 384, 2), (1080, 1920, 224, 398, 2),
             (2160, 3840, 384, 384, 2),   # 4K: 20 horizontal taps -> widest planar-kernel instantiation
             (1440, 2560, 384, 384, 3),   # bicubic 1440p: 27 horizontal taps
             (384, 384, 384, 384, 2)]     # identity scale


@pytest.mark.parametrize("H,W,oh,ow,rs", PRE_CASES)
def test_preprocess_resize_bit_exact(H, W, oh, ow, rs):
    frames = synth.noise_frames(3, H, W, seed=H + rs)
    want = preprocess_ref.resize_u8(frames.numpy(), oh, ow, rs)
    got = ops.preprocess(frames.to(DEV), oh, ow, rs, layout=ops.LAYOUT_U8_CHW).cpu().numpy()
    nd = int((got != want).sum())
    print(f"resize {H}x{W}->{oh}x{ow} rs={rs}: mismatching bytes {nd}/{want.size}")
    assert nd == 0


@pytest.mark.parametrize("H,W,size,rs", [(1080, 1920, 384, 2), (123, 211, 56, 2), (1080, 1920, 384, 3)])
def test_preprocess_layouts_bit_exact(H, W, size, rs):
    frames = synth.scene_frames_np(3, 2, H, W)
    pv = preprocess_ref.pixel_values(frames, size, size, rs)
    dev_frames = torch.from_numpy(frames).to(DEV)
    got_f32 = ops.preprocess(dev_frames, size, size, rs, layout=ops.LAYOUT_F32_CHW).cpu().numpy()
    assert np.array_equal(got_f32.view(np.uint32), pv.view(np.uint32)), "fp32 pixel_values differ"
    want_bf16 = torch.from_numpy(pv).to(torch.bfloat16)
    got_bf16 = ops.preprocess(dev_frames, size, size, rs, layout=ops.LAYOUT_BF16_CHW).cpu()
    assert torch.equal(got_bf16.view(torch.int16), want_bf16.view(torch.int16)), "bf16 CHW differs"
    ld = (3 * 14 * 14 + 7) // 8 * 8
    want_patch = torch.from_numpy(preprocess_ref.patchify(pv, 14, ld)).to(torch.bfloat16)
    got_patch = ops.preprocess(dev_frames, size, size, rs, layout=ops.LAYOUT_BF16_PATCH, patch=14).cpu()
    assert got_patch.shape == want_patch.shape
    assert torch.equal(got_patch.view(torch.int16), want_patch.view(torch.int16)), "bf16 patches differ"


ANY_GEOMETRY = [(60, 90, 140, 140, 2), (60, 90, 140, 140, 3),      # both axes up-scaled (a small crop through encode_image)
                (200, 100, 140, 140, 2), (100, 500, 384, 384, 3),  # one axis up, one down
                (30, 31, 56, 56, 2), (1, 1, 28, 28, 3),            # tiny sources
                (400, 1000, 37, 61, 2),                            # 16x down: 35 horizontal taps, beyond the tuned kernels
                (2160, 3840, 126, 224, 3)]                         # 17x bicubic down: 71 taps


@pytest.mark.parametrize("H,W,oh,ow,rs", ANY_GEOMETRY)
def test_preprocess_any_geometry_bit_exact(H, W, oh, ow, rs):
    """Up-scaling and very wide windows go through the general two-pass kernel (gvl_resize_two_pass_u8 with ATen's own
    tables): every layout must still equal the oracle bit for bit."""
    frames = synth.noise_frames(2, H, W, seed=H + W + rs)
    f = frames.to(DEV)
    want_u8 = preprocess_ref.resize_u8(frames.numpy(), oh, ow, rs)
    assert np.array_equal(ops.preprocess(f, oh, ow, rs, layout=ops.LAYOUT_U8_CHW).cpu().numpy(), want_u8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    pv = preprocess_ref.pixel_values(frames.numpy(), oh, ow, rs, mean, std)
    got = ops.preprocess(f, oh, ow, rs, mean, std, layout=ops.LAYOUT_F32_CHW).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), pv.view(np.uint32))
    got16 = ops.preprocess(f, oh, ow, rs, mean, std, layout=ops.LAYOUT_BF16_CHW).cpu()
    assert torch.equal(got16.view(torch.int16), torch.from_numpy(pv).to(torch.bfloat16).view(torch.int16))
    if oh >= 14 and ow >= 14:
        want_p = torch.from_numpy(preprocess_ref.patchify(pv, 14, 592)).to(torch.bfloat16)
        got_p = ops.preprocess(f, oh, ow, rs, mean, std, layout=ops.LAYOUT_BF16_PATCH, patch=14, ld=592).cpu()
        assert torch.equal(got_p.view(torch.int16), want_p.view(torch.int16))


def test_preprocess_legacy_kernel_matches():
    """The v1 kernel stays bit-exact: it is the fallback for out-of-range geometries."""
    frames = synth.noise_frames(2, 1080, 1920, seed=11)
    want = preprocess_ref.resize_u8(frames.numpy(), 384, 384, 2)
    with ops.preprocess_path(ops.PRE_PATH_V1):
        got = ops.preprocess(frames.to(DEV), 384, 384, 2, layout=ops.LAYOUT_U8_CHW).cpu().numpy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("B,H,W,size,patch,mean,std", [
    (5, 1080, 1920, 384, 14, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),          # the headline geometry (streaming 5:1 kernel)
    (3, 1080, 1920, 384, 16, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),          # 24 x 24 patches of 16: strips without slack
    (2, 1080, 1920, 384, 14, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),  # ImageNet constants
    (3, 385, 1920, 384, 14, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),           # odd height, near-identity vertical scale
    (2, 2160, 1920, 384, 14, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),          # vertical windows of 12 rows: falls back
    (2, 720, 1280, 256, 16, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5)),           # another 5:1 width
])
def test_preprocess_streaming_kernel_bit_exact(B, H, W, size, patch, mean, std):
    """Patch layout on 5:1 widths (auto path = preprocess_stream.cu where it applies) against the oracle AND against
    the planar kernel, on noise frames (every tap matters) plus a structured scene."""
    frames = torch.cat([synth.noise_frames(B - 1, H, W, seed=H + patch), torch.from_numpy(synth.scene_frames_np(5, 1, H, W))])
    pv = preprocess_ref.pixel_values(frames.numpy(), size, size, 2, image_mean=mean, image_std=std)
    ld = (3 * patch * patch + 7) // 8 * 8
    want = torch.from_numpy(preprocess_ref.patchify(pv, patch, ld)).to(torch.bfloat16)
    dev = frames.to(DEV)
    got = ops.preprocess(dev, size, size, 2, image_mean=mean, image_std=std, patch=patch).cpu()
    with ops.preprocess_path(ops.PRE_PATH_PLANAR):
        planar = ops.preprocess(dev, size, size, 2, image_mean=mean, image_std=std, patch=patch).cpu()
    assert torch.equal(planar.view(torch.int16), want.view(torch.int16)), "planar kernel differs from the oracle"
    nd = int((got.view(torch.int16) != want.view(torch.int16)).sum())
    assert nd == 0, f"{nd} of {want.numel()} elements differ from the oracle"


def test_preprocess_streaming_kernel_batch64_equals_planar():
    """BASELINE's batch: 64 x 1080p, streaming kernel == planar kernel on every byte (run length 3, 1728 CTAs)."""
    frames = synth.noise_frames(64, 1080, 1920, seed=64).to(DEV)
    got = ops.preprocess(frames, 384, 384, 2)
    with ops.preprocess_path(ops.PRE_PATH_PLANAR):
        planar = ops.preprocess(frames, 384, 384, 2)
    assert torch.equal(got.view(torch.int16), planar.view(torch.int16))


def test_preprocess_table_cache_is_bounded():
    """More distinct geometries than the table caches hold (64): the caches are emptied and rebuilt, results stay exact
    (variable-size crops / mixed-resolution videos must not grow device memory without bound)."""
    dev_frames = {}
    for i in range(70):
        H = 40 + i
        frames = synth.noise_frames(1, H, 64, seed=i)
        got = ops.preprocess(frames.to(DEV), 16, 24, 2, layout=ops.LAYOUT_U8_CHW).cpu().numpy()
        if i % 23 == 0 or i == 69:
            assert np.array_equal(got, preprocess_ref.resize_u8(frames.numpy(), 16, 24, 2)), f"geometry {i}"
    # the headline geometry after the trim (streaming kernel tables rebuilt)
    frames = synth.noise_frames(1, 1080, 1920, seed=3)
    pv = preprocess_ref.pixel_values(frames.numpy(), 384, 384, 2)
    want = torch.from_numpy(preprocess_ref.patchify(pv, 14, 592)).to(torch.bfloat16)
    assert torch.equal(ops.preprocess(frames.to(DEV), 384, 384, 2).cpu().view(torch.int16), want.view(torch.int16))


def test_preprocess_unaligned_base_pointer():
    """Frames whose base address is not 16-byte aligned take the byte-copy staging path."""
    H, W = 60, 101
    frames = synth.noise_frames(2, H, W, seed=3)
    buf = torch.empty(frames.numel() + 1, dtype=torch.uint8, device=DEV)
    view = buf[1:].view(2, H, W, 3)
    view.copy_(frames)
    want = preprocess_ref.resize_u8(frames.numpy(), 28, 28, 2)
    got = ops.preprocess(view, 28, 28, 2, layout=ops.LAYOUT_U8_CHW).cpu().numpy()
    assert np.array_equal(got, want)


# ----------------------------------------------------------------------------------------------- top-k
def test_topk_matches_oracle_with_ties():
    g = torch.Generator().manual_seed(21)
    N, D, Q, k = 5000, 4096, 11, 16
    index = torch.randn(N, D, generator=g).to(torch.bfloat16)
    index[4000] = index[17]          # exact duplicates -> exact ties, lower index must come first
    index[4500] = index[17]
    queries = torch.randn(Q, D, generator=g).to(torch.bfloat16)
    queries[0] = index[17]           # query 0 hits the tie group
    queries[1] = 0                   # zero query: all scores 0 -> indices 0..k-1
    want_s, want_i, margins = siglip_ref.cosine_topk(index.float().numpy(), queries.float().numpy(), k)
    got_s, got_i = ops.topk_cosine(index.to(DEV), queries.to(DEV), k)
    torch.cuda.synchronize()
    got_i, got_s = got_i.cpu().numpy(), got_s.cpu().numpy()
    print("top-k min margin", float(np.min(margins[2:])))
    assert np.array_equal(got_i, want_i), f"indices differ\n{got_i}\n{want_i}"
    assert np.allclose(got_s, want_s, atol=2e-6)
    assert list(got_i[0][:3]) == [17, 4000, 4500]
    assert list(got_i[1]) == list(range(k))


@pytest.mark.timeout(900)
def test_topk_config5_72k_index():
    """BASELINE.json configs[4]: cosine top-16 over a 72 000-row (10 h @ 2 fps) 4096-d bf16 timeline index for 128
    queries, indices against the float64 oracle.  Rows are clustered like scene embeddings (pairwise cosine ~0.9);
    every one of the 2048 (query, rank) positions must hold the oracle's row (the north star's "bit-exact")."""
    N, D, Q, k = 72000, 4096, 128, 16
    g = torch.Generator(device=DEV).manual_seed(5)
    centers = torch.randn(N // 60, D, device=DEV, generator=g)
    index = (centers.repeat_interleave(60, 0) + 0.35 * torch.randn(N, D, device=DEV, generator=g)).to(torch.bfloat16)
    qs = torch.randint(0, N // 60, (Q,), device=DEV, generator=g)
    queries = (centers[qs] + 0.35 * torch.randn(Q, D, device=DEV, generator=g)).to(torch.bfloat16)
    del centers
    got_s, got_i = ops.topk_cosine(index, queries, k)
    torch.cuda.synchronize()
    want_s, want_i, _ = siglip_ref.cosine_topk(index.float().cpu().numpy(), queries.float().cpu().numpy(), k)
    got_i = got_i.cpu().numpy().astype(np.int64)
    mism = np.argwhere(got_i != want_i)
    print(f"config 5 top-{k}: {mism.shape[0]} of {Q * k} positions differ from the float64 oracle")
    assert np.abs(got_s.cpu().numpy() - want_s).max() < 2e-6
    # the kernel's final ranking is done on float64 re-scored candidates (topk_refine_kernel): bit-exact indices
    assert mism.shape[0] == 0, mism[:8]


@pytest.mark.parametrize("mode", [ops.TOPK_SCAN, ops.TOPK_TENSOR])
@pytest.mark.parametrize("N", [5000, 4999])
def test_topk_windowed_matches_oracle(mode, N):
    """Per-query row ranges (the +-window time filter of a timestamp-ordered index) fused with the ranking: both scoring
    paths against the float64 oracle, including empty ranges, ranges shorter than k, ranges that are not multiples of
    the tensor path's 8-row granularity and an index whose row count is not one either."""
    D, Q, k = 512, 24, 8
    g = torch.Generator().manual_seed(11)
    index = torch.randn(N, D, generator=g).to(torch.bfloat16)
    index[4000] = index[17]          # exact ties inside / across windows
    index[4500] = index[17]
    queries = torch.randn(Q, D, generator=g).to(torch.bfloat16)
    queries[0] = index[17]
    lo = torch.randint(0, N - 600, (Q,), generator=g)
    hi = lo + torch.randint(1, 600, (Q,), generator=g)
    lo[0], hi[0] = 0, N              # whole index
    lo[1], hi[1] = 100, 100          # empty
    lo[2], hi[2] = 203, 206          # fewer rows than k
    lo[3], hi[3] = 3990, 4600        # contains two of the tied rows
    queries[3] = index[17]
    want_s, want_i = siglip_ref.cosine_topk_windowed(index.float().numpy(), queries.float().numpy(), k, lo.tolist(), hi.tolist())
    got_s, got_i = ops.topk_cosine(index.to(DEV), queries.to(DEV), k, row_lo=lo.to(torch.int32).to(DEV),
                                   row_hi=hi.to(torch.int32).to(DEV), mode=mode)
    got_s, got_i = got_s.cpu().numpy(), got_i.cpu().numpy().astype(np.int64)
    assert np.array_equal(got_i, want_i), np.argwhere(got_i != want_i)[:5]
    fin = np.isfinite(want_s)
    assert np.array_equal(np.isfinite(got_s), fin) and np.abs(got_s[fin] - want_s[fin]).max() < 5e-6
    assert list(got_i[0][:3]) == [17, 4000, 4500] and list(got_i[3][:2]) == [4000, 4500]


def test_topk_tensor_path_equals_scan_path_72k():
    """configs[4] size: the tensor-core scoring path (one skinny GEMM for all 128 queries + exact re-scoring of the
    near-top candidates) and the CUDA-core scan give bit-identical top-16 rows and scores; prints the time of both."""
    N, D, Q, k = 72000, 4096, 128, 16
    g = torch.Generator(device=DEV).manual_seed(5)
    centers = torch.randn(N // 60, D, device=DEV, generator=g)
    index = (centers.repeat_interleave(60, 0) + 0.35 * torch.randn(N, D, device=DEV, generator=g)).to(torch.bfloat16)
    qs = torch.randint(0, N // 60, (Q,), device=DEV, generator=g)
    queries = (centers[qs] + 0.35 * torch.randn(Q, D, device=DEV, generator=g)).to(torch.bfloat16)
    del centers
    inv = ops.row_inv_norm(index)
    res, times = {}, {}
    for name, mode, kw in (("scan", ops.TOPK_SCAN, {}), ("tensor", ops.TOPK_TENSOR, {}),
                           ("tensor+cached norms", ops.TOPK_TENSOR, {"inv_norm": inv})):
        ops.topk_cosine(index, queries, k, mode=mode, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res[name] = ops.topk_cosine(index, queries, k, mode=mode, **kw)
        e1.record()
        torch.cuda.synchronize()
        times[name] = e0.elapsed_time(e1)
    print("top-16 of 128 queries over 72k x 4096: " + ", ".join(f"{n} {t:.2f} ms" for n, t in times.items()))
    s0, i0 = res["scan"]
    for name in ("tensor", "tensor+cached norms"):
        s1, i1 = res[name]
        # the tensor path re-scores its near-top candidates with the scan path's arithmetic: identical results
        assert torch.equal(i0, i1) and torch.equal(s0, s1)
