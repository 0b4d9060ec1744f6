"""Host-side logic that needs no GPU: drop-in construction and error behaviour, cache layouts, sharding."""
import os

import numpy as np
import pytest
import torch

from gameplay_vision_llm_b200 import feature_cache as fc
from gameplay_vision_llm_b200.pipeline import shard_range
from gameplay_vision_llm_b200.projector import MultiModalProjector, ProjectorBank, ReasoningCoreConfig
from gameplay_vision_llm_b200.siglip_semantic_encoder import (BatchFeature, NaFlexConfig, SigLIPSemanticEncoder,
                                                               spec_from_state_dict)
from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_projector_state_dict, synth_siglip_state_dict


def test_flops_per_frame_matches_baseline_md():
    assert SiglipVisionSpec.so400m().flops_per_frame() == 670_389_441_536
    assert SiglipVisionSpec.so400m().flops_per_frame(None) == 670_389_441_536 - 42_991_616


def test_naflex_config_defaults_match_reference():
    c = NaFlexConfig()
    assert (c.model_name, c.device, c.dtype, c.base_resolution, c.embedding_dim, c.batch_size) == (
        "google/siglip2-so400m-patch14-384", "cuda", torch.bfloat16, 384, 1152, 16)


def test_encoder_raises_instead_of_placeholder():
    enc = SigLIPSemanticEncoder(NaFlexConfig(device="cpu", synthetic_seed=0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc.encode_image(np.zeros((8, 8, 3), np.uint8))
    enc2 = SigLIPSemanticEncoder(NaFlexConfig(model_name="/nonexistent/checkpoint", device="cuda"))
    with pytest.raises(RuntimeError):
        enc2.encoder._load_model()
    assert enc.encode_masked_regions(None, []) == []  # `if not masks: return []` (reference :503-504)
    mask = np.ones((8, 8), np.bool_)
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # the region route has no CPU path either
        enc.encode_masked_regions(np.zeros((8, 8, 3), np.uint8), [("e", mask)])


def test_spec_from_state_dict_roundtrip():
    spec = SiglipVisionSpec.tiny()
    assert spec_from_state_dict(synth_siglip_state_dict(spec, 0), 384) == spec


def test_batchfeature_seam():
    bf = BatchFeature({"pixel_values": torch.zeros(1, 3, 4, 4)})
    moved = bf.to("cpu")
    assert set(dict(**moved)) == {"pixel_values"} and moved.pixel_values.shape == (1, 3, 4, 4)


def test_projector_bank_construction_and_state_dict_layout(tmp_path):
    bank = ProjectorBank(ReasoningCoreConfig(device="cpu")).to("cpu")  # what the reference's tests construct
    assert bank.siglip_proj.net[0].weight.shape == (4096, 1152)
    assert bank.videomae_proj.net[0].weight.shape == (4096, 768)
    assert bank.video_proj.net[0].weight.shape == (4096, 1408) and bank.audio_proj.net[0].weight.shape == (4096, 1024)
    p = tmp_path / "projector_weights.pt"
    bank.save_weights(str(p))
    sd = torch.load(p, weights_only=False)
    assert set(sd) == {"siglip", "videomae", "audio", "video"}
    assert set(sd["siglip"]) == {"net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias"}
    bank2 = ProjectorBank(ReasoningCoreConfig(device="cpu"))
    bank2.load_weights(str(p))
    assert torch.equal(bank2.siglip_proj.net[2].weight, bank.siglip_proj.net[2].weight)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bank.project_region(torch.zeros(2, 1152))


def test_projector_accepts_reference_state_dict():
    m = MultiModalProjector(1152, 4096)
    m.load_state_dict(synth_projector_state_dict(1152, 4096, 1))
    assert list(m.state_dict()) == ["net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias"]


def test_shard_range_covers_timeline_in_order():
    for n, w in ((3600, 8), (3600, 1), (10, 4), (72000, 8), (5, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
    lo, hi = shard_range(100, 1, 3, align=16)
    assert lo % 16 == 0


def test_feature_cache_layout_and_eviction(tmp_path):
    video = tmp_path / "clip.mp4"
    video.write_bytes(os.urandom(3 * 1024 * 1024 + 17))
    cache = fc.FeatureCache(str(tmp_path / "cache"), max_cache_size_gb=1.0)
    h = cache._get_video_hash(str(video))
    assert len(h) == 16 and cache._get_cache_path(str(video)).name == f"clip_{h}.pt"
    # same recipe as the reference: sha256(str(size) + first MiB + last MiB)
    import hashlib
    raw = video.read_bytes()
    want = hashlib.sha256(str(len(raw)).encode() + raw[: 1 << 20] + raw[-(1 << 20):]).hexdigest()[:16]
    assert h == want
    emb = torch.randn(3, 1152).to(torch.bfloat16)
    payload = {"siglip": fc.siglip_cache_entries([0.0, 1.0, 2.0], emb), "videomae": []}
    assert not cache.has_features(str(video))
    cache.save_features(str(video), payload)
    assert cache.has_features(str(video))
    fresh = fc.FeatureCache(str(tmp_path / "cache"))
    got = fresh.load_features(str(video))
    assert got["siglip"][1]["timestamp"] == 1.0 and torch.equal(got["siglip"][2]["embedding"], emb[2])
    # corrupt file -> unlinked, None
    fresh2 = fc.FeatureCache(str(tmp_path / "cache"))
    fresh2._get_cache_path(str(video)).write_bytes(b"garbage")
    assert fresh2.load_features(str(video)) is None and not fresh2._get_cache_path(str(video)).exists()
    assert cache._get_video_hash(str(tmp_path / "missing.mp4")) == ""


def test_embeddings_pt_and_npz_layouts(tmp_path):
    emb = torch.randn(4, 1152).to(torch.bfloat16)
    recs = fc.siglip_embedding_records([0.0, 1.0, 2.0, 3.0], emb)
    assert recs[0]["entity_type"] == "full_frame" and recs[3]["embedding_shape"] == [1152]
    data = fc.write_embeddings_pt(str(tmp_path / "v_embeddings.pt"), recs)
    back = torch.load(tmp_path / "v_embeddings.pt", weights_only=False)
    assert set(back) >= {"siglip", "videomae", "wav2vec2", "hico", "visual_events", "audio_transcripts"}
    assert back["siglip"][2]["shape"] == [1152] and back["siglip"][2]["embedding"].dtype == torch.bfloat16
    stacked = torch.stack([e["embedding"] for e in back["siglip"]])  # what demo_projector_inference.py does
    assert stacked.shape == (4, 1152) and data["siglip"][0]["timestamp"] == 0.0
    video = tmp_path / "v.mp4"
    video.write_bytes(b"x" * 100)
    d = fc.write_perception_npz(str(tmp_path / "pc"), str(video), siglip=emb, timestamps=np.arange(4.0),
                                frame_indices=np.arange(4))
    z = np.load(os.path.join(d, "siglip.npz"))
    assert z["embeddings"].shape == (4, 1152) and z["embeddings"].dtype == np.float32


def test_timeline_window_range_matches_reference_filters():
    """window_range (binary search) == the reference's literal list filters (oracle.window_rows), including
    timestamps that sit exactly on t +- w and windows that fall off either end of the timeline."""
    from gameplay_vision_llm_b200.timeline import TimelineEmbeddingIndex
    from oracle import siglip_ref
    rng = np.random.default_rng(7)
    for fps in (1.0, 2.0, 29.97):
        ts = np.arange(400, dtype=np.float64) / fps
        idx = TimelineEmbeddingIndex(400, 8, device="cpu", timestamps=ts)
        cases = [(0.0, 30.0), (ts[-1], 30.0), (ts[100], 0.0), (ts[100], 1.0 / fps), (ts[57] + 0.25 / fps, 2.5 / fps),
                 (-100.0, 10.0), (1e6, 10.0), (ts[200], 1e9)]
        cases += [(float(rng.uniform(-5, ts[-1] + 5)), float(rng.uniform(0, 40))) for _ in range(40)]
        for t, w in cases:
            for inclusive in (False, True):
                assert idx.window_range(t, w, inclusive) == siglip_ref.window_rows(ts, t, w, inclusive), (fps, t, w, inclusive)
    # irregular (but sorted) timestamps with duplicates
    ts = np.sort(np.round(rng.uniform(0, 50, 300), 1))
    idx = TimelineEmbeddingIndex(300, 8, device="cpu", timestamps=ts)
    for t in (0.0, 10.0, 10.05, 25.3, 50.0):
        for w in (0.0, 0.1, 0.3, 5.0):
            for inclusive in (False, True):
                assert idx.window_range(t, w, inclusive) == siglip_ref.window_rows(ts, t, w, inclusive)


def test_oracle_windowed_topk_and_hybrid_merge():
    from oracle import siglip_ref
    rng = np.random.default_rng(3)
    e = rng.standard_normal((50, 16))
    q = rng.standard_normal((3, 16))
    s, i = siglip_ref.cosine_topk_windowed(e, q, 4, [0, 10, 20], [50, 12, 20])
    full = siglip_ref.cosine_topk(e, q[:1], 4)
    assert np.array_equal(i[0], full[1][0]) and np.allclose(s[0], full[0][0])
    assert set(i[1][:2]) == {10, 11} and list(i[1][2:]) == [-1, -1] and np.isinf(s[1][2:]).all()
    assert list(i[2]) == [-1] * 4
    assert siglip_ref.hybrid_merge([4, 5, 6], [9, 5, 2]) == [4, 5, 6, 9, 2]


def test_empty_timeline_and_empty_query_return_nothing_like_the_reference():
    """TimelineRetriever.retrieve_by_semantic returns [] without embeddings (qwen_reasoning_core.py:1508-1511); the
    index does the same without touching the device."""
    from gameplay_vision_llm_b200.timeline import TimelineEmbeddingIndex
    idx = TimelineEmbeddingIndex(0, 16, device="cpu")
    q = torch.randn(1, 16)
    assert idx.retrieve_by_semantic(q[0]) == [] and idx.retrieve_by_timestamp(3.0) == []
    assert idx.hybrid_retrieve(q) == [] and idx.hybrid_retrieve(q, timestamp=1.0) == []
    s, i = idx.search(q, top_k=4)
    assert s.shape == (1, 4) and i.tolist() == [[-1] * 4]
    sel, emb = idx.window(5.0)
    assert sel.size == 0 and emb.shape == (0, 16)
    idx2 = TimelineEmbeddingIndex(5, 16, device="cpu")
    s, i = idx2.search(torch.empty(0, 16), top_k=3)
    assert s.shape == (0, 3) and i.shape == (0, 3)


def test_checkpoint_loader_safetensors_bin_and_configs(tmp_path):
    """`_load_state_dict` / `_load_checkpoint_configs` / `spec_from_state_dict` on files laid out like an HF
    `SiglipModel` checkpoint: `vision_model.*` keys next to text-tower keys, config.json with a nested
    `vision_config`, preprocessor_config.json (reference: AutoModel / AutoProcessor.from_pretrained, :184-204)."""
    import json

    from safetensors.torch import save_file

    from gameplay_vision_llm_b200.siglip_semantic_encoder import (_load_checkpoint_configs, _load_state_dict,
                                                                  spec_from_state_dict)
    from gameplay_vision_llm_b200.weights import SiglipVisionSpec, synth_siglip_state_dict
    spec = SiglipVisionSpec.tiny()
    sd = synth_siglip_state_dict(spec, seed=0)
    assert all(k.startswith("vision_model.") for k in sd)
    full = dict(sd)
    full.update({"text_model.embeddings.token_embedding.weight": torch.zeros(8, 4), "logit_scale": torch.zeros(1),
                 "logit_bias": torch.zeros(1), "text_model.encoder.layers.0.mlp.fc1.weight": torch.zeros(4, 4)})
    d = tmp_path / "ckpt"
    d.mkdir()
    save_file({k: v.contiguous() for k, v in full.items()}, str(d / "model.safetensors"))
    json.dump({"model_type": "siglip", "vision_config": {"num_attention_heads": spec.heads, "layer_norm_eps": 1e-6,
                                                        "hidden_size": spec.hidden}, "text_config": {}},
              open(d / "config.json", "w"))
    json.dump({"resample": 3, "image_mean": [0.5, 0.5, 0.5], "image_std": [0.5, 0.5, 0.5]},
              open(d / "preprocessor_config.json", "w"))
    for target in (str(d), str(d / "model.safetensors")):
        got = _load_state_dict(target)
        assert set(got) == set(full) and all(torch.equal(got[k], full[k]) for k in sd)
        vision, pre = _load_checkpoint_configs(target)
        assert vision["num_attention_heads"] == spec.heads and pre["resample"] == 3
        got_spec = spec_from_state_dict(got, image=spec.image, heads=vision["num_attention_heads"], eps=vision["layer_norm_eps"])
        assert got_spec == spec
    # torch .bin with bare vision-tower keys (a `SiglipVisionModel.state_dict()` without the prefix)
    bare = {k[len("vision_model."):]: v for k, v in sd.items()}
    d2 = tmp_path / "ckpt_bin"
    d2.mkdir()
    torch.save(bare, d2 / "pytorch_model.bin")
    got = _load_state_dict(str(d2))
    assert set(got) == set(bare) and _load_checkpoint_configs(str(d2)) == (None, None)
    # head count: explicit > table of published widths > error (never a silent guess)
    assert spec_from_state_dict(got, image=spec.image, heads=spec.heads) == spec
    so = SiglipVisionSpec.so400m()
    shapes = {"embeddings.patch_embedding.weight": torch.empty(1152, 3, 14, 14, device="meta"),
              "embeddings.position_embedding.weight": torch.empty(729, 1152, device="meta"),
              "encoder.layers.26.mlp.fc1.weight": torch.empty(4304, 1152, device="meta"),
              "encoder.layers.0.mlp.fc1.weight": torch.empty(4304, 1152, device="meta")}
    assert spec_from_state_dict(shapes) == so  # 1152 -> 16 heads of 72 from the table
    assert spec_from_state_dict(got, image=spec.image) == spec  # 144 = 2 x 72, and 64 does not divide it
    shapes576 = dict(shapes)
    shapes576["embeddings.patch_embedding.weight"] = torch.empty(576, 3, 14, 14, device="meta")
    with pytest.raises(RuntimeError, match="head count"):  # 576 % 72 == 0 used to be mis-read as 8 heads of 72
        spec_from_state_dict(shapes576)
    with pytest.raises(RuntimeError, match="not supported"):
        spec_from_state_dict(shapes576, heads=6)  # head dim 96
    with pytest.raises(RuntimeError, match="not a local checkpoint"):
        _load_state_dict(str(tmp_path / "missing"))
