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
    pv = videomae_ref.pixel_values(frames, image_mean=(0.5,) * 3, image_std=(0.5,) * 3)  # processor class defaults
    u8 = np.rint(pv * 127.5 + 127.5).astype(np.uint8)
    assert np.array_equal(u8[:, :, ::28, :], gold["pre1080_rows"])
    assert hashlib.sha256(u8.tobytes()).digest() == gold["pre1080_sha"].tobytes()
    # the checkpoint's constants (ImageNet mean / std): fp32 pixel_values bit for bit
    pv = videomae_ref.pixel_values(frames)
    assert np.array_equal(pv[:, :, ::28, :].view(np.uint32), gold["pre1080_imagenet_rows"].view(np.uint32))
    assert hashlib.sha256(np.ascontiguousarray(pv).tobytes()).digest() == gold["pre1080_imagenet_sha"].tobytes()


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
    # use_mean_pooling=True checkpoints: no final LayerNorm in VideoMAEModel
    import dataclasses
    spec_nf = dataclasses.replace(spec, final_norm=False)
    seams = {}
    pooled = videomae_ref.encoder_forward(synth_videomae_state_dict(spec_nf, seed=2), torch.from_numpy(pv), spec.frames,
                                          spec.heads, spec.patch, spec.tubelet, spec.eps, seams=seams)
    assert np.abs(seams["last_hidden_state"].numpy() - gold["tiny_nofinal_last_hidden_state"]).max() < 1e-4
    assert np.abs(pooled.numpy() - gold["tiny_nofinal_pooled"]).max() < 1e-4


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


def test_clip_windows_match_both_reference_loops():
    from gameplay_vision_llm_b200.videomae_encoder import clip_windows
    for n in (0, 1, 15, 16, 17, 24, 31, 32, 40, 100):
        # scripts/extract_features.py:355-365
        want = [(a, min(a + 16, n)) for a in range(0, n, 16)]
        assert clip_windows(n) == want
        # scripts/realtime_inference.py:352-355 (the `len(chunk) < chunk_size: continue` guard never fires)
        want_rt = [(i, i + 16) for i in range(0, n - 16 + 1, 16 // 2)]
        assert clip_windows(n, 16, 8, drop_tail=True) == want_rt
    assert clip_windows(40, 16, 8, drop_tail=True) == [(0, 16), (8, 24), (16, 32), (24, 40)]


def test_from_checkpoint_reads_configs(tmp_path):
    """`from_checkpoint` takes geometry and preprocessing constants from the files beside the weights; construction
    needs a GPU, so only the parsing is checked here: no CUDA device -> RuntimeError after the files were read."""
    import json

    from safetensors.torch import save_file

    from gameplay_vision_llm_b200.videomae_encoder import IMAGENET_MEAN, VideoMAEClipEncoder
    spec = VideoMAESpec.tiny()
    save_file({k: v.contiguous() for k, v in synth_videomae_state_dict(spec, seed=2).items()}, str(tmp_path / "model.safetensors"))
    json.dump({"hidden_size": 128, "intermediate_size": 272, "num_hidden_layers": 2, "num_attention_heads": 2,
               "image_size": 48, "patch_size": 16, "num_frames": 4, "tubelet_size": 2, "use_mean_pooling": False},
              open(tmp_path / "config.json", "w"))
    json.dump({"image_mean": list(IMAGENET_MEAN), "image_std": [0.229, 0.224, 0.225], "resample": 2},
              open(tmp_path / "preprocessor_config.json", "w"))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            VideoMAEClipEncoder.from_checkpoint(str(tmp_path), device="cuda")
