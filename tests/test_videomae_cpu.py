"""The VideoMAE oracle (oracle/videomae_ref.py) pinned against the golden vectors produced by HF
`VideoMAEImageProcessor` + `VideoMAEModel` (tests/golden/make_golden_videomae.py), plus the host-side clip
bookkeeping.  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from gameplay_vision_llm_b200 import synth
from gameplay_vision_llm_b200.videomae_encoder import resize_geometry
from gameplay_vision_llm_b200.weights import (VideoMAESpec, sinusoid_table, synth_projector_state_dict,
                                                synth_videomae_state_dict)
from oracle import siglip_ref, videomae_ref


def test_resize_geometry_1080p():
    assert videomae_ref.resize_geometry(1080, 1920) == (224, 398, 0, 87)
    assert resize_geometry(1080, 1920) == (224, 398, 0, 87)
    assert resize_geometry(1920, 1080) == (398, 224, 87, 0)
    assert resize_geometry(123, 211, 48, 48) == videomae_ref.resize_geometry(123, 211, 48, 48)


def test_preprocess_oracle_matches_hf_golden(golden_dir):
    gold = np.load(f"{golden_dir}/golden_videomae.npz")
    frames = np.concatenate([synth.scene_frames_np(0, 2), synth.noise_frames(2, seed=1001).numpy()], 0)
    pv = videomae_ref.pixel_values(frames)
    u8 = np.rint(pv * 127.5 + 127.5).astype(np.uint8)
    assert np.array_equal(u8[:, :, ::28, :], gold["pre1080_rows"])
    assert hashlib.sha256(u8.tobytes()).digest() == gold["pre1080_sha"].tobytes()


def test_sinusoid_tables_agree():
    a, b = sinusoid_table(1568, 768), videomae_ref.sinusoid_table(1568, 768)
    assert torch.equal(a, b)


def test_tiny_encoder_oracle_matches_hf_golden(golden_dir):
    gold = np.load(f"{golden_dir}/golden_videomae.npz")
    spec = VideoMAESpec.tiny()
    sd = synth_videomae_state_dict(spec, seed=2)
    small = synth.noise_frames(2 * spec.frames, 123, 211, seed=9).numpy()
    pv = videomae_ref.pixel_values(small, spec.image, spec.image)
    assert np.array_equal(pv.reshape(2, spec.frames, 3, spec.image, spec.image), gold["tiny_pixel_values"])
    seams = {}
    pooled = videomae_ref.encoder_forward(sd, torch.from_numpy(pv), spec.frames, spec.heads, spec.patch, spec.tubelet,
                                          spec.eps, seams=seams)
    for k in ("embeddings", "layer0", "last_hidden_state"):
        assert np.abs(seams[k].numpy() - gold["tiny_" + k]).max() < 1e-4, k
    assert np.abs(pooled.numpy() - gold["tiny_pooled"]).max() < 1e-4


@pytest.mark.timeout(600)
def test_base_encoder_oracle_matches_hf_golden(golden_dir):
    """Full VideoMAE-base geometry: one 16-frame 1080p clip against the HF fp32 output (a few seconds of CPU)."""
    gold = np.load(f"{golden_dir}/golden_videomae.npz")
    spec = VideoMAESpec.base()
    sd = synth_videomae_state_dict(spec, seed=2)
    wsum = float(sum(v.double().sum() for v in sd.values()))
    assert abs(wsum - float(gold["base_weight_checksum"][0])) < 1e-6 * max(1.0, abs(wsum))
    clip = np.concatenate([synth.scene_frames_np(0, 8), synth.scene_frames_np(30, 8)], 0)
    pv = torch.from_numpy(videomae_ref.pixel_values(clip))
    with torch.no_grad():
        pooled = videomae_ref.encoder_forward(sd, pv, spec.frames, spec.heads, spec.patch, spec.tubelet, spec.eps)
        proj = siglip_ref.projector_forward(synth_projector_state_dict(spec.hidden, 4096, seed=3), pooled)
    assert np.abs(pooled.numpy() - gold["base_pooled"]).max() < 2e-3
    assert np.abs(proj.numpy() - gold["base_projected"]).max() < 2e-3


def test_tubelet_patch_layout():
    pv = torch.arange(4 * 3 * 32 * 32, dtype=torch.float32).reshape(4, 3, 32, 32)  # 1 clip of 4 frames, 2x2 patches
    cols = videomae_ref.tubelet_patches(pv, 4, 16, 2)
    assert cols.shape == (1, 2 * 2 * 2, 3 * 2 * 256)
    # token (tt=1, py=0, px=1), column (c=2, kt=1, ky=3, kx=5) == frame 3, channel 2, pixel (3, 16+5)
    assert cols[0, 1 * 4 + 0 * 2 + 1, 2 * 512 + 1 * 256 + 3 * 16 + 5] == pv[3, 2, 3, 21]
