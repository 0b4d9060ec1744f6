"""The oracle pinned against the golden vectors produced by the real third-party implementations the
reference calls (HF SiglipVisionModel / SiglipImageProcessor, torchvision, torch.nn) — see
tests/golden/make_golden.py.  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from gameplay_vision_llm_b200 import synth
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_projector_state_dict, synth_siglip_state_dict
from oracle import preprocess_ref, siglip_ref


def test_preprocess_oracle_matches_hf_golden(golden_dir):
    gold = np.load(f"{golden_dir}/golden_preprocess.npz")
    small = synth.noise_frames(3, 123, 211, seed=7).numpy()
    assert np.array_equal(preprocess_ref.resize_u8(small, 56, 56, 2), gold["small_123x211_to56_rs2"])
    for name, frames in (("noise", synth.noise_frames(2, seed=1000).numpy()), ("scene", synth.scene_frames_np(0, 2))):
        for rs in (2, 3):
            u8 = preprocess_ref.resize_u8(frames, 384, 384, rs)
            assert np.array_equal(u8[:, :, ::48, :], gold[f"{name}_rs{rs}_rows"])
            sha = hashlib.sha256(np.ascontiguousarray(u8).tobytes()).digest()
            assert sha == gold[f"{name}_rs{rs}_sha"].tobytes(), f"{name} rs={rs}: full-image hash differs from HF"


def test_preprocess_oracle_matches_torchvision_live():
    tvF = pytest.importorskip("torchvision.transforms.v2.functional")
    from torchvision.transforms import InterpolationMode as IM
    x = synth.noise_frames(2, 250, 333, seed=5)
    for rs, mode in ((2, IM.BILINEAR), (3, IM.BICUBIC)):
        want = tvF.resize(x.permute(0, 3, 1, 2), [97, 64], interpolation=mode, antialias=True).numpy()
        assert np.array_equal(preprocess_ref.resize_u8(x.numpy(), 97, 64, rs), want)


@pytest.mark.parametrize("H,W,oh,ow", [(60, 90, 140, 140), (200, 100, 140, 140), (100, 500, 384, 384), (30, 31, 56, 56),
                                       (400, 1000, 37, 61)])
def test_preprocess_oracle_matches_torchvision_live_any_geometry(H, W, oh, ow):
    """Up-scaling (ATen keeps the filter support at the interpolation radius) and very wide down-scaling windows."""
    tvF = pytest.importorskip("torchvision.transforms.v2.functional")
    from torchvision.transforms import InterpolationMode as IM
    x = synth.noise_frames(2, H, W, seed=H + W)
    for rs, mode in ((2, IM.BILINEAR), (3, IM.BICUBIC)):
        want = tvF.resize(x.permute(0, 3, 1, 2), [oh, ow], interpolation=mode, antialias=True).numpy()
        assert np.array_equal(preprocess_ref.resize_u8(x.numpy(), oh, ow, rs), want)


def test_pixel_values_and_patchify():
    frames = synth.noise_frames(2, 60, 80, seed=1).numpy()
    pv = preprocess_ref.pixel_values(frames, 28, 28, 2)
    assert pv.dtype == np.float32 and pv.shape == (2, 3, 28, 28) and np.abs(pv).max() <= 1.0
    p = preprocess_ref.patchify(pv, 14, 592)
    assert p.shape == (2 * 4, 592) and np.all(p[:, 588:] == 0)
    # row = b*gh*gw + py*gw + px ; col = c*196 + ky*14 + kx
    assert p[1 * 4 + 1 * 2 + 0, 2 * 196 + 3 * 14 + 5] == pv[1, 2, 14 + 3, 5]


def test_tiny_tower_oracle_matches_hf_golden(golden_dir):
    gold = np.load(f"{golden_dir}/golden_tiny.npz")
    spec = SiglipVisionSpec.tiny()
    sd = synth_siglip_state_dict(spec, seed=0)
    frames = synth.noise_frames(3, 123, 211, seed=7).numpy()
    pv = preprocess_ref.pixel_values(frames, spec.image, spec.image, 2)
    assert np.array_equal(pv, gold["pixel_values"])
    seams = {}
    pooled = siglip_ref.vision_forward(sd, torch.from_numpy(pv), spec.heads, spec.patch, spec.eps, seams=seams)
    for k in ("embeddings", "layer0", "layer1", "last_hidden_state"):
        assert np.abs(seams[k].numpy() - gold[k]).max() < 1e-4, k
    assert np.abs(pooled.numpy() - gold["pooled"]).max() < 1e-4
    proj = siglip_ref.projector_forward(synth_projector_state_dict(spec.hidden, 256, seed=1), pooled)
    assert np.abs(proj.numpy() - gold["projected"]).max() < 1e-4


@pytest.mark.timeout(600)
def test_so400m_oracle_matches_hf_golden_one_frame(golden_dir):
    """Full-size restatement vs the HF fp32 output for frame 0 of the G1 stream (≈10 s of CPU)."""
    gold = np.load(f"{golden_dir}/golden_so400m.npz")
    spec = SiglipVisionSpec.so400m()
    sd = synth_siglip_state_dict(spec, seed=0)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    assert abs(wsum - float(gold["weight_checksum"][0])) < 1e-6 * max(1.0, abs(wsum))
    pv = torch.from_numpy(preprocess_ref.pixel_values(synth.scene_frames_np(0, 1), 384, 384, 2))
    with torch.no_grad():
        pooled = siglip_ref.vision_forward(sd, pv, spec.heads, spec.patch, spec.eps)
        proj = siglip_ref.projector_forward(synth_projector_state_dict(1152, 4096, seed=1), pooled)
    assert np.abs(pooled.numpy()[0] - gold["pooled"][0]).max() < 2e-3
    assert np.abs(proj.numpy()[0] - gold["projected"][0]).max() < 2e-3
    cos = torch.nn.functional.cosine_similarity(pooled[0], torch.from_numpy(gold["pooled"][0]), dim=0)
    assert cos > 0.999999


def test_cosine_topk_oracle_order_and_ties():
    rng = np.random.default_rng(0)
    e = rng.standard_normal((50, 16)).astype(np.float32)
    e[30] = e[3]
    e[40] = 2.0 * e[3]  # same direction -> same cosine up to rounding; float64 keeps them tied or ordered stably
    q = e[3:4].copy()
    s, idx, margin = siglip_ref.cosine_topk(e, q, 4)
    assert idx[0, 0] == 3 and set(idx[0, :3]) == {3, 30, 40} and np.all(np.diff(s[0]) <= 1e-12)
    t = torch.nn.functional.normalize(torch.from_numpy(q), dim=1) @ torch.nn.functional.normalize(torch.from_numpy(e), dim=1).T
    assert set(torch.topk(t[0], 4).indices.tolist()) == set(idx[0].tolist())


def test_scene_generators_agree():
    a = synth.scene_frames_np(28, 3, 90, 160)
    b = synth.scene_frames(28, 3, 90, 160).numpy()
    assert np.array_equal(a, b) and a.std() > 20
