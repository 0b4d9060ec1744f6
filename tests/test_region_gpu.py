"""Masked-region variant through the C ABI: Pillow-exact region bytes and patch rows, the position
table interpolation, the pooling kernels, and `encode_masked_regions` / `encode_with_context` end to end against the
goldens the REFERENCE'S OWN `encode_masked_regions` produced (tests/golden/make_golden_regions.py) and the fp32 oracle.
Tolerance: resized bytes and bf16 patch rows bit-exact; embeddings cosine >= 0.999 + stated max-abs."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gameplay_vision_llm_b200 import ops, synth  # noqa: E402
from gameplay_vision_llm_b200.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder  # noqa: E402
from gameplay_vision_llm_b200.weights import (SiglipPack, SiglipVisionSpec, synth_ren_projection_state_dict,  # noqa: E402
                                                synth_siglip_state_dict)
from oracle import preprocess_ref, region_ref, siglip_ref  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_regions import MID_CFG, MID_RECTS, MID_SPEC, SO_RECTS, rect_mask  # noqa: E402

DEV = "cuda:0"


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu(), dim=-1)


# (frame H, W, boxes, sizes (out_h, out_w)): down- and up-scaling, identity axes, 1-pixel crops, whole-frame crops
KERNEL_CASES = [
    (270, 480, [(14, 21, 326, 129), (194, 0, 266, 270), (90, 90, 210, 210), (4, 4, 13, 10), (0, 197, 480, 233)],
     [(42, 140), (140, 42), (140, 140), (84, 140), (42, 140)]),
    (1080, 1920, [(200, 156, 1400, 684), (852, 0, 1428, 1050), (0, 0, 1920, 1080), (340, 248, 1060, 872)],
     [(154, 378), (378, 196), (210, 378), (322, 378)]),
    (64, 96, [(10, 10, 11, 11), (0, 0, 96, 64), (5, 7, 47, 49)], [(14, 14), (42, 70), (42, 42)]),
]


@pytest.mark.parametrize("H,W,boxes,sizes", KERNEL_CASES)
def test_region_kernel_is_bit_exact_against_the_pillow_oracle(H, W, boxes, sizes):
    frame = synth.noise_frames(1, H, W, seed=H + W)[0]
    canvas = (max(s[0] for s in sizes), max(s[1] for s in sizes))
    lut = ops.region_lut(region_ref.IMAGENET_MEAN, region_ref.IMAGENET_STD).to(DEV)
    patches, u8 = ops.region_patches(frame.to(DEV), boxes, sizes, canvas, lut, patch=14, ld=592, want_u8=True)
    torch.cuda.synchronize()
    gh, gw = canvas[0] // 14, canvas[1] // 14
    assert patches.shape == (len(boxes) * gh * gw, 592)
    f = frame.numpy()
    for r, ((x1, y1, x2, y2), (oh, ow)) in enumerate(zip(boxes, sizes)):
        want = region_ref.pil_resize_bicubic_u8(f[y1:y2, x1:x2], ow, oh)
        got = u8[r].cpu().numpy()
        assert np.array_equal(got[:oh, :ow], want), f"region {r}: resized bytes differ from Pillow's"
        assert not got[oh:].any() and not got[:, ow:].any(), f"region {r}: padding is not zero"
        # patch rows: the reference's normalisation in fp32, zero padding, cast to bf16, im2col
        t = torch.from_numpy(want.copy()).float().permute(2, 0, 1) / 255.0
        t = (t - torch.tensor(region_ref.IMAGENET_MEAN).view(3, 1, 1)) / torch.tensor(region_ref.IMAGENET_STD).view(3, 1, 1)
        t = torch.nn.functional.pad(t, (0, canvas[1] - ow, 0, canvas[0] - oh))
        want_p = torch.from_numpy(preprocess_ref.patchify(t[None].numpy(), 14, 592)).to(torch.bfloat16)
        got_p = patches[r * gh * gw:(r + 1) * gh * gw].cpu()
        assert torch.equal(got_p.view(torch.int16), want_p.view(torch.int16)), f"region {r}: patch rows differ"


def test_region_kernel_rejects_bad_geometry():
    frame = synth.noise_frames(1, 64, 96, seed=1)[0].to(DEV)
    lut = ops.region_lut(region_ref.IMAGENET_MEAN, region_ref.IMAGENET_STD).to(DEV)
    with pytest.raises(RuntimeError, match="leaves the"):
        ops.region_patches(frame, [(90, 0, 100, 10)], [(14, 14)], (14, 14), lut)
    with pytest.raises(RuntimeError, match="exceeds"):
        ops.region_patches(frame, [(0, 0, 10, 10)], [(28, 14)], (14, 14), lut)
    with pytest.raises(ValueError, match="empty box"):
        ops.region_patches(frame, [(5, 5, 5, 9)], [(14, 14)], (14, 14), lut)


@pytest.mark.parametrize("g,D,gh,gw", [(27, 1152, 13, 27), (27, 1152, 27, 16), (10, 216, 3, 10), (10, 216, 10, 7), (4, 144, 9, 5)])
def test_position_interpolation_kernel(g, D, gh, gw):
    pos = torch.randn(g * g, D, generator=torch.Generator().manual_seed(g + gh)).to(torch.bfloat16)
    got = ops.interpolate_pos(pos.to(DEV), gh, gw).float().cpu()
    want = region_ref.interpolate_pos(pos.float(), gh, gw)
    # fp32 evaluation rounded to bf16: within one bf16 step of the float64 restatement
    assert ((got - want).abs() <= want.abs() * 2.0 ** -7 + 1e-5).all()  # + fp32 evaluation noise where taps cancel
    p = pos.to(DEV)
    assert ops.interpolate_pos(p, g, g) is p  # the checkpoint's own grid: the table itself, as in HF


def test_max_and_mean_tokens():
    B, T, D = 3, 37, 216
    x = torch.randn(B * T, D, generator=torch.Generator().manual_seed(0)).to(torch.bfloat16).to(DEV)
    assert torch.equal(ops.max_tokens(x, B, T).cpu(), x.float().view(B, T, D).max(dim=1)[0].cpu())
    assert torch.allclose(ops.mean_tokens(x, B, T).cpu(), x.float().view(B, T, D).mean(dim=1).cpu(), atol=1e-5)


def _encoder(spec, pool, bs, cfg_kw, fold=False):
    cfg = NaFlexConfig(device=DEV, embedding_dim=spec.hidden, pool_strategy=pool, batch_size=bs,
                       state_dict=synth_siglip_state_dict(spec, seed=0), num_attention_heads=spec.heads, **cfg_kw)
    enc = SigLIPSemanticEncoder(cfg)
    enc.projection.load_state_dict(synth_ren_projection_state_dict(spec.hidden, seed=3))
    return enc


@pytest.mark.parametrize("pool", ["mean", "cls", "max"])
@pytest.mark.parametrize("bs", [16, 1])
def test_encode_masked_regions_mid_tower_vs_reference_golden(golden_dir, pool, bs):
    gold = np.load(f"{golden_dir}/golden_regions.npz")
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    masks = [(f"e{i}", rect_mask(frame.shape, r)) for i, r in enumerate(MID_RECTS)]
    enc = _encoder(MID_SPEC, pool, bs, MID_CFG)
    res = enc.encode_masked_regions(frame, masks)
    assert [r.entity_id for r in res] == [m[0] for m in masks]
    assert np.array_equal(np.array([r.original_bbox for r in res]), gold["mid_bbox"])
    assert np.array_equal(np.array([r.aspect_ratio for r in res]), gold["mid_aspect"])
    got = torch.stack([r.embedding for r in res])
    assert got.dtype == torch.float32 and got.device.type == "cpu"  # `.cpu()` fp32 rows like the reference (:553)
    want = torch.from_numpy(gold[f"mid_{pool}_bs{bs}"])
    cos, err = _cos(got, want), (got - want).abs().max().item()
    print(f"regions mid pool={pool} bs={bs}: cos min {cos.min():.6f} max_abs {err:.4f} (|want| max {want.abs().max():.3f})")
    assert cos.min() > 0.999 and err < 0.015 * want.abs().max().item() + 0.005  # measured 0.005 x max


def test_encode_masked_regions_accepts_a_resident_frame_and_empty_input():
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    masks = [("a", rect_mask(frame.shape, MID_RECTS[0])), ("b", rect_mask(frame.shape, MID_RECTS[1]))]
    enc = _encoder(MID_SPEC, "mean", 16, MID_CFG)
    assert enc.encode_masked_regions(frame, []) == []
    host = enc.encode_masked_regions(frame, masks)
    dev = enc.encode_masked_regions(torch.from_numpy(frame).to(DEV), masks)
    assert all(torch.equal(a.embedding, b.embedding) for a, b in zip(host, dev))


def test_encode_with_context_vs_reference_golden(golden_dir):
    gold = np.load(f"{golden_dir}/golden_regions.npz")
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    enc = _encoder(MID_SPEC, "mean", 16, MID_CFG)
    a, b = enc.encode_with_context(frame, rect_mask(frame.shape, MID_RECTS[2]), context_radius=30)
    assert np.array_equal(np.array([a.original_bbox, b.original_bbox]), gold["mid_context_bbox"])
    cos = _cos(torch.stack([a.embedding, b.embedding]), torch.from_numpy(gold["mid_context"]))
    assert cos.min() > 0.999


def test_non_square_region_raises_like_the_reference_without_interpolation():
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    enc = _encoder(MID_SPEC, "mean", 16, dict(MID_CFG, interpolate_pos_encoding=False))
    with pytest.raises(RuntimeError, match="positions"):
        enc.encode_masked_regions(frame, [("wide", rect_mask(frame.shape, MID_RECTS[0]))])
    # a region that resizes to the checkpoint's own square grid works either way (the reference's working case)
    sq = enc.encode_masked_regions(frame, [("sq", rect_mask(frame.shape, MID_RECTS[2]))])
    enc2 = _encoder(MID_SPEC, "mean", 16, MID_CFG)
    assert torch.equal(sq[0].embedding, enc2.encode_masked_regions(frame, [("sq", rect_mask(frame.shape, MID_RECTS[2]))])[0].embedding)


def test_encoder_forward_on_a_non_square_input_vs_oracle():
    """`SigLIPEncoder.forward` (reference :246-289) returns (sequence, pooled) for an h x w input."""
    spec = MID_SPEC
    sd = synth_siglip_state_dict(spec, seed=0)
    pv = torch.randn(2, 3, 42, 98, generator=torch.Generator().manual_seed(5)).clamp(-2, 2)
    pv = pv.to(torch.bfloat16).float()
    pos = region_ref.interpolate_pos(sd["vision_model.embeddings.position_embedding.weight"], 3, 7)
    seams = {}
    want = siglip_ref.vision_forward(sd, pv, spec.heads, spec.patch, spec.eps, seams=seams, pos=pos)
    enc = _encoder(spec, "mean", 16, MID_CFG)
    seq, pooled = enc.encoder(pv.to(DEV))
    assert seq.shape == (2, 21, spec.hidden)
    assert _cos(seq.float(), seams["last_hidden_state"]).min() > 0.999 and _cos(pooled.float(), want).min() > 0.999


@pytest.mark.parametrize("fold_ln", [False, True])
@pytest.mark.timeout(900)
def test_encode_masked_regions_so400m_vs_reference_golden(golden_dir, fold_ln):
    """Full-size tower, three regions of a 1080p frame: one call per detection (scripts/extract_features.py:568 — grids
    12 x 27, 27 x 15 and the untouched 27 x 27) and one padded batch of three, against the reference's own outputs."""
    gold = np.load(f"{golden_dir}/golden_regions.npz")
    spec = SiglipVisionSpec.so400m()
    enc = _encoder(spec, "mean", 16, {"fold_layernorm": fold_ln})
    frame = synth.scene_frames_np(40, 1)[0]
    masks = [(f"s{i}", rect_mask(frame.shape, r)) for i, r in enumerate(SO_RECTS)]
    single = [enc.encode_masked_regions(frame, [mk])[0] for mk in masks]
    batched = enc.encode_masked_regions(frame, masks)
    assert np.array_equal(np.array([r.original_bbox for r in single]), gold["so_bbox"])
    assert np.array_equal(np.array([r.aspect_ratio for r in single]), gold["so_aspect"])
    for name, res in (("single", single), ("batched", batched)):
        got, want = torch.stack([r.embedding for r in res]), torch.from_numpy(gold[f"so_mean_{name}"])
        cos, err = _cos(got, want), (got - want).abs().max().item()
        print(f"regions so400m fold_ln={fold_ln} {name}: cos {cos.tolist()} max_abs {err:.4f} (|want| max {want.abs().max():.3f})")
        assert cos.min() > 0.999 and err < 0.015 * want.abs().max().item() + 0.005  # measured 0.005 x max


def test_run_siglip_encoder_equals_the_reference_loop_call_by_call():
    """`pipeline.run_siglip_encoder` (scripts/extract_features.py:502-610) batches full frames and groups detections by
    target size; every record must equal what the reference's loop order produces with one encoder call each."""
    from PIL import Image

    from gameplay_vision_llm_b200.pipeline import run_siglip_encoder
    enc = _encoder(MID_SPEC, "mean", 4, MID_CFG)
    raw = synth.scene_frames_np(3, 6, 270, 480)
    frames = [(i * 0.5, Image.fromarray(f)) for i, f in enumerate(raw)]
    sam = [
        {"timestamp": 0.5, "bbox": [40, 30, 300, 120], "entity_type": "player", "entity_id": "p1", "description": "the player"},
        {"timestamp": 0.5, "bbox": [60, 40, 320, 130], "entity_type": "enemy"},           # same target size as p1
        {"timestamp": 0.5, "bbox": None, "entity_type": "ghost"},                           # skipped (:558)
        {"timestamp": 0.5, "bbox": [200.7, 10.2, 260.9, 250.0], "entity_type": "tower", "entity_id": "t1"},
        {"timestamp": 2.0, "bbox": [100, 100, 200, 200], "entity_type": "item", "entity_id": "i1"},
    ]
    got = run_siglip_encoder(frames, DEV, sam_results=sam, entity_tracker=object(), encoder=enc)
    # the reference's loop, literally, on the same encoder
    want = []
    for ts, frame in frames:
        dets = [d for d in sam if d["timestamp"] == ts]
        frame_np = np.array(frame.convert("RGB"))
        if not dets:
            e = enc.encode_image(frame)
            want.append({"timestamp": ts, "embedding": e.cpu(), "embedding_shape": list(e.shape), "entity_type": "full_frame",
                         "description": "Full frame encoding (no SAM detection)"})
            continue
        for det in dets:
            if det.get("bbox") is None:
                continue
            entity_type = det.get("entity_type", "unknown")
            entity_id = det.get("entity_id", f"{entity_type}_{ts}")
            x1, y1, x2, y2 = [int(c) for c in det["bbox"]]
            mask = np.zeros(frame_np.shape[:2], dtype=np.bool_)
            mask[y1:y2, x1:x2] = True
            emb = enc.encode_masked_regions(frame_np, [(entity_id, mask)])[0]
            want.append({"timestamp": ts, "embedding": emb.embedding.cpu(), "embedding_shape": list(emb.embedding.shape),
                         "entity_type": entity_type, "entity_id": entity_id,
                         "description": det.get("description", f"Detected {entity_type}"), "bbox": det["bbox"]})
    assert len(got) == len(want) == 8
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for k in w:
            if k == "embedding":
                assert g[k].dtype == w[k].dtype and torch.equal(g[k], w[k]), (w["timestamp"], w["entity_type"])
            else:
                assert g[k] == w[k], k
    # fallback branch (:590-607): no detections -> every frame whole, batched
    plain = run_siglip_encoder(frames, DEV, encoder=enc)
    assert [d["description"] for d in plain] == ["Full frame encoding"] * 6
    assert all(torch.equal(d["embedding"], enc.encode_image(f).cpu()) for d, (_, f) in zip(plain, frames))


@pytest.mark.parametrize("pool", ["mean", "cls", "max"])
@pytest.mark.parametrize("fold_ln", [False, True])
def test_ragged_pass_equals_one_call_per_region_bit_for_bit(pool, fold_ln):
    """`encode_regions_individually` = ONE ragged tower pass over all regions (`gvl_siglip_forward_ragged`); every row
    must equal `encode_masked_regions(frame, [m])[0]`, the reference's one-call-per-detection loop, bit for bit —
    with the LayerNorms as kernels and folded into the GEMMs, and when max_tokens splits the pass."""
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    rects = MID_RECTS + [(60, 40, 320, 130), (210, 5, 270, 245), (120, 90, 220, 190), (0, 0, 480, 270)]
    masks = [(f"e{i}", rect_mask(frame.shape, r)) for i, r in enumerate(rects)]
    enc = _encoder(MID_SPEC, pool, 16, dict(MID_CFG, fold_layernorm=fold_ln))
    want = [enc.encode_masked_regions(frame, [mk])[0] for mk in masks]
    for max_tokens in (32768, 150):
        got = enc.encode_regions_individually(frame, masks, max_tokens=max_tokens)
        assert [g.entity_id for g in got] == [w.entity_id for w in want]
        for g, w in zip(got, want):
            assert g.original_bbox == w.original_bbox and g.aspect_ratio == w.aspect_ratio
            assert torch.equal(g.embedding, w.embedding), (g.entity_id, max_tokens)


@pytest.mark.timeout(900)
def test_ragged_pass_so400m_equals_one_call_per_region():
    spec = SiglipVisionSpec.so400m()
    enc = _encoder(spec, "mean", 16, {})
    frame = synth.scene_frames_np(40, 1)[0]
    masks = [(f"s{i}", rect_mask(frame.shape, r)) for i, r in enumerate(SO_RECTS + [(100, 100, 1100, 540), (0, 0, 1920, 1080)])]
    want = [enc.encode_masked_regions(frame, [mk])[0] for mk in masks]
    got = enc.encode_regions_individually(frame, masks)
    assert all(torch.equal(g.embedding, w.embedding) for g, w in zip(got, want))


@pytest.mark.parametrize("hd,H,items", [(72, 3, [21, 100, 30, 351, 729, 1, 130]), (64, 2, [200, 64, 65, 128, 127, 1000]),
                                         (72, 16, [324, 405, 729, 621])])
def test_varlen_attention_equals_one_launch_per_item(hd, H, items):
    """One launch over a ragged batch (query-tile table) against one gvl_attention_bf16 launch per item: bit-identical
    (a tile that runs past its item reads the neighbour's rows; those keys are masked, those query rows never stored)."""
    M = sum(items)
    qkv = (torch.randn(M, 3 * H * hd, generator=torch.Generator().manual_seed(M)) * 1.5).to(torch.bfloat16).to(DEV)
    got = ops.attention_varlen(qkv, items, H, hd)
    r0 = 0
    for T in items:
        want = ops.attention(qkv[r0:r0 + T].contiguous(), 1, T, H, hd)
        assert torch.equal(got[r0:r0 + T], want), (T, r0)
        r0 += T
    assert torch.isfinite(got.float()).all()


def test_region_kernel_random_geometries_and_ragged_form():
    """40 random crops of one frame at random target sizes (up- and down-scaling, 1-pixel crops, whole-frame crops):
    bytes equal to the Pillow oracle, and the ragged form (every region on its own canvas, rows back to back) holds
    exactly the rows the padded-canvas form holds inside each region's rectangle."""
    rng = np.random.default_rng(11)
    H, W = 300, 500
    frame = synth.noise_frames(1, H, W, seed=3)[0]
    boxes, sizes = [], []
    for _ in range(40):
        x1, y1 = int(rng.integers(0, W - 1)), int(rng.integers(0, H - 1))
        x2, y2 = int(rng.integers(x1 + 1, W + 1)), int(rng.integers(y1 + 1, H + 1))
        boxes.append((x1, y1, x2, y2))
        sizes.append((14 * int(rng.integers(1, 15)), 14 * int(rng.integers(1, 15))))
    boxes[0], boxes[1] = (0, 0, W, H), (7, 9, 8, 10)
    canvas = (max(s[0] for s in sizes), max(s[1] for s in sizes))
    lut = ops.region_lut(region_ref.IMAGENET_MEAN, region_ref.IMAGENET_STD).to(DEV)
    fd = frame.to(DEV)
    padded, u8 = ops.region_patches(fd, boxes, sizes, canvas, lut, patch=14, ld=592, want_u8=True)
    ragged, _ = ops.region_patches(fd, boxes, sizes, None, lut, patch=14, ld=592)
    torch.cuda.synchronize()
    f = frame.numpy()
    gwc, tok = canvas[1] // 14, (canvas[0] // 14) * (canvas[1] // 14)
    r0 = 0
    for r, ((x1, y1, x2, y2), (oh, ow)) in enumerate(zip(boxes, sizes)):
        want = region_ref.pil_resize_bicubic_u8(f[y1:y2, x1:x2], ow, oh)
        assert np.array_equal(u8[r, :oh, :ow].cpu().numpy(), want), (r, boxes[r], sizes[r])
        gh, gw = oh // 14, ow // 14
        mine = ragged[r0:r0 + gh * gw].view(gh, gw, 592)
        theirs = padded[r * tok:(r + 1) * tok].view(canvas[0] // 14, gwc, 592)[:gh, :gw]
        assert torch.equal(mine, theirs), r
        r0 += gh * gw
    assert r0 == ragged.shape[0]


def test_resizer_and_extractor_helpers_match_the_reference_classes():
    """`AspectPreservingResizer.resize_with_aspect_ratio`, `RegionExtractor.extract_masked_region` /
    `prepare_region_tensor` (reference :137-163, :301-367): same return types and values as the oracle's restatement
    (PIL bytes bit-exact, fp32 tensor bit-exact)."""
    from PIL import Image
    enc = _encoder(MID_SPEC, "mean", 16, MID_CFG)
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    mask = rect_mask(frame.shape, MID_RECTS[0])
    region, bbox = enc.region_extractor.extract_masked_region(frame, mask)
    assert bbox == region_ref.extract_bbox(frame.shape, mask) and isinstance(region, Image.Image)
    resized, ar = enc.region_extractor.resizer.resize_with_aspect_ratio(region)
    x1, y1, x2, y2 = bbox
    th, tw = region_ref.compute_optimal_size(y2 - y1, x2 - x1, MID_CFG["base_resolution"], MID_CFG["min_resolution"],
                                             MID_CFG["max_resolution"])
    assert resized.size == (tw, th) and ar == (x2 - x1) / (y2 - y1)
    assert np.array_equal(np.asarray(resized), np.asarray(region.resize((tw, th), Image.Resampling.BICUBIC)))
    tensor, ar2 = enc.region_extractor.prepare_region_tensor(region)
    want, _ = region_ref.prepare_region_tensor(frame[y1:y2, x1:x2], True, MID_CFG["base_resolution"],
                                               MID_CFG["min_resolution"], MID_CFG["max_resolution"])
    assert ar2 == ar and torch.equal(tensor, want)
