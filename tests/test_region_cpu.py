"""Masked-region variant, CPU side: the oracle pinned against Pillow / torch live and against the
goldens produced by the reference's own `encode_masked_regions` (tests/golden/make_golden_regions.py); the host-side
geometry and coefficient tables of the product pinned against the oracle.  No GPU compute."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from gameplay_vision_llm_b200 import ops, synth
from gameplay_vision_llm_b200.siglip_semantic_encoder import AspectPreservingResizer, NaFlexConfig, RegionExtractor
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_ren_projection_state_dict, synth_siglip_state_dict
from oracle import region_ref

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
from make_golden_regions import MID_CFG, MID_RECTS, MID_SPEC, SO_RECTS, rect_mask  # noqa: E402

GEOMETRIES = [(200, 333, 126, 378), (50, 40, 378, 294), (540, 960, 210, 378), (378, 378, 378, 378), (100, 378, 98, 378),
              (378, 100, 378, 98), (17, 900, 14, 378), (5, 3, 378, 224), (1, 1, 14, 14)]


@pytest.mark.parametrize("h,w,oh,ow", GEOMETRIES)
def test_pil_bicubic_restatement_is_bit_exact_against_pillow(h, w, oh, ow):
    Image = pytest.importorskip("PIL.Image")
    img = np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
    assert np.array_equal(region_ref.pil_resize_bicubic_u8(img, ow, oh), want)


@pytest.mark.parametrize("a,b", [(333, 378), (40, 294), (1920, 378), (378, 378), (17, 14), (3, 224), (1080, 210), (1, 14)])
def test_library_coefficient_tables_equal_the_oracle(a, b):
    tab, k = ops.pil_bicubic_taps(a, b)
    xmin, cnt, kk = region_ref.pil_coeffs(a, b)
    assert k == kk.shape[1]
    assert np.array_equal(tab[:b], xmin) and np.array_equal(tab[b:2 * b], cnt)
    assert np.array_equal(tab[2 * b:].reshape(b, k), kk)


@pytest.mark.parametrize("gh,gw", [(27, 13), (16, 27), (27, 27), (9, 27), (30, 40), (1, 1)])
def test_position_interpolation_restatement_against_torch(gh, gw):
    g, D = 27, 24
    pos = torch.randn(g * g, D, generator=torch.Generator().manual_seed(gh * 100 + gw))
    got = region_ref.interpolate_pos(pos, gh, gw)
    if (gh, gw) == (g, g):
        assert got is pos
        return
    want = torch.nn.functional.interpolate(pos.reshape(1, g, g, D).permute(0, 3, 1, 2), size=(gh, gw), mode="bicubic",
                                           align_corners=False).permute(0, 2, 3, 1).reshape(gh * gw, D)
    assert torch.allclose(got, want, atol=2e-5, rtol=0)  # torch evaluates source index and weights in fp32


def test_host_geometry_equals_the_reference_rules():
    cfg = NaFlexConfig()
    rz = AspectPreservingResizer(cfg)
    rng = np.random.default_rng(0)
    for _ in range(300):
        h, w = int(rng.integers(1, 1200)), int(rng.integers(1, 2000))
        assert rz.compute_optimal_size(h, w) == region_ref.compute_optimal_size(h, w)
    shape = (270, 480, 3)
    for rect in MID_RECTS + [(0, 0, 480, 270), (479, 269, 480, 270), (10, 10, 11, 200)]:
        m = rect_mask(shape, rect)
        assert RegionExtractor.region_bbox(shape, m) == region_ref.extract_bbox(shape, m)
    ragged = np.zeros(shape[:2], np.bool_)
    ragged[50, 100] = ragged[200, 30] = ragged[120, 400] = True
    assert RegionExtractor.region_bbox(shape, ragged) == region_ref.extract_bbox(shape, ragged)


def test_region_lut_is_the_reference_normalisation():
    """[3, 256] table == the reference's three fp32 operations applied to an image holding every byte value."""
    img = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, axis=2)
    want, _ = region_ref.prepare_region_tensor(img, preserve_aspect_ratio=False, base=16)  # 16 x 16 -> 16 x 16: no resize
    lut = ops.region_lut(region_ref.IMAGENET_MEAN, region_ref.IMAGENET_STD, torch.float32)
    assert torch.equal(lut, want.reshape(3, 256))


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def _oracle_rows(spec, frame, rects, pool, bs, cfg):
    sd = synth_siglip_state_dict(spec, seed=0)
    psd = synth_ren_projection_state_dict(spec.hidden, seed=3)
    masks = [(f"e{i}", rect_mask(frame.shape, r)) for i, r in enumerate(rects)]
    res = region_ref.encode_masked_regions(sd, psd, frame, masks, spec.heads, spec.patch, spec.eps, pool, bs,
                                           cfg.get("base_resolution", 384), cfg.get("min_resolution", 128),
                                           cfg.get("max_resolution", 768))
    return torch.stack([r[1] for r in res]).numpy(), res


@pytest.mark.parametrize("pool", ["mean", "cls", "max"])
@pytest.mark.parametrize("bs", [16, 1])
def test_oracle_equals_the_reference_encode_masked_regions_mid(golden_dir, pool, bs):
    gold = np.load(f"{golden_dir}/golden_regions.npz")
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    got, res = _oracle_rows(MID_SPEC, frame, MID_RECTS, pool, bs, MID_CFG)
    want = gold[f"mid_{pool}_bs{bs}"]
    assert np.abs(got - want).max() < 2e-4, np.abs(got - want).max()
    assert np.array_equal(np.array([r[2] for r in res]), gold["mid_bbox"])
    assert np.array_equal(np.array([r[3] for r in res]), gold["mid_aspect"])


def test_oracle_resized_bytes_equal_the_reference_pillow_bytes(golden_dir):
    gold = np.load(f"{golden_dir}/golden_regions.npz")
    frame = synth.scene_frames_np(7, 1, 270, 480)[0]
    for i, rect in enumerate(MID_RECTS):
        x1, y1, x2, y2 = region_ref.extract_bbox(frame.shape, rect_mask(frame.shape, rect))
        th, tw = region_ref.compute_optimal_size(y2 - y1, x2 - x1, MID_CFG["base_resolution"], MID_CFG["min_resolution"],
                                                 MID_CFG["max_resolution"])
        got = region_ref.pil_resize_bicubic_u8(frame[y1:y2, x1:x2], tw, th)
        assert np.array_equal(_sha(got), gold["mid_resized_sha"][i]), i


def test_oracle_equals_the_reference_encode_masked_regions_so400m(golden_dir):
    """One full-size region (the landscape 12 x 27 grid) through the fp32 oracle: 27 layers on the CPU."""
    gold = np.load(f"{golden_dir}/golden_regions.npz")
    spec = SiglipVisionSpec.so400m()
    frame = synth.scene_frames_np(40, 1)[0]
    got, res = _oracle_rows(spec, frame, SO_RECTS[:1], "mean", 16, {})
    want = gold["so_mean_single"][:1]
    cos = torch.nn.functional.cosine_similarity(torch.from_numpy(got), torch.from_numpy(want)).item()
    assert cos > 0.999999 and np.abs(got - want).max() < 5e-4, (cos, np.abs(got - want).max())
    assert res[0][2] == tuple(gold["so_bbox"][0])


def test_box_mask_extent_equals_the_materialised_mask():
    """`BoxMask` (the caller's `mask[y1:y2, x1:x2] = True`, extract_features.py:563-565) must give the bounding box the
    reference's `np.where` finds on the materialised array — including negative / out-of-frame slice bounds."""
    from gameplay_vision_llm_b200.siglip_semantic_encoder import BoxMask
    shape = (54, 96, 3)
    rng = np.random.default_rng(1)
    cases = [(0, 54, 0, 96), (10, 10, 5, 50), (-5, 20, -10, 30), (40, 200, 90, 300), (-60, -50, 3, 9), (53, 54, 95, 96)]
    cases += [tuple(int(v) for v in rng.integers(-70, 130, 4)) for _ in range(300)]
    for y1, y2, x1, x2 in cases:
        bm = BoxMask(shape[:2], y1, y2, x1, x2)
        arr = np.zeros(shape[:2], np.bool_)
        arr[y1:y2, x1:x2] = True
        assert np.array_equal(np.asarray(bm), arr)
        assert RegionExtractor.region_bbox(shape, bm) == region_ref.extract_bbox(shape, arr), (y1, y2, x1, x2)
