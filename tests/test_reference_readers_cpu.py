"""the cache files this build writes are read back by the REFERENCE's own readers,
imported from /root/reference when it is mounted (the build container); skipped on the GPU box where it is not.

* `benchmarks/perception_cache.py` `PerceptionCache.load` (:285-372) on `write_perception_npz` output,
* `scripts/train_projectors.py` `MultimodalProjectorDataset` reader loop (:195-279) and the stacking step of
  `scripts/demo_projector_inference.py:104-117` on `write_embeddings_pt` output.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

from gameplay_vision_llm_b200 import feature_cache as fc

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod  # dataclasses resolves annotations through sys.modules
    spec.loader.exec_module(mod)
    return mod


def test_perception_npz_has_reference_layout(tmp_path):
    video = tmp_path / "v.mp4"
    video.write_bytes(b"x" * 1000)
    emb = torch.randn(5, 1152).to(torch.bfloat16)
    vm = torch.randn(2, 768)
    d = fc.write_perception_npz(str(tmp_path / "pc"), str(video), siglip=emb, videomae=vm, timestamps=np.arange(5) / 4.0,
                                video_duration_sec=1.25, perception_config={"fps": 4.0})
    meta = json.load(open(os.path.join(d, "metadata.json")))
    # the keys of CachedFeatures.to_metadata() (benchmarks/perception_cache.py:99-116)
    assert set(meta) == {"video_hash", "video_path", "video_duration_sec", "cache_version", "cached_at",
                         "perception_config", "has_siglip", "has_videomae", "has_sam", "has_ocr", "has_audio",
                         "has_timeline", "has_kb", "num_frames"}
    assert meta["has_siglip"] and meta["has_videomae"] and not meta["has_sam"] and meta["num_frames"] == 5
    index = json.load(open(tmp_path / "pc" / "index.json"))
    assert index[meta["video_hash"]] == meta and os.path.basename(d) == meta["video_hash"]
    # a second video is added to the same index, the first entry stays
    video2 = tmp_path / "w.mp4"
    video2.write_bytes(b"y" * 10)
    fc.write_perception_npz(str(tmp_path / "pc"), str(video2), siglip=emb[:2], timestamps=np.arange(2.0))
    assert len(json.load(open(tmp_path / "pc" / "index.json"))) == 2
    back = fc.load_perception_npz(str(tmp_path / "pc"), str(video))
    assert np.array_equal(back["siglip_embeddings"], emb.float().numpy()) and back["metadata"] == meta
    assert fc.load_perception_npz(str(tmp_path / "pc"), str(tmp_path / "other.mp4")) is None
    # a path that does not exist hashes the bare string, like the reference (:173-175)
    import hashlib
    assert fc.perception_video_hash("/no/such.mp4") == hashlib.md5(b"/no/such.mp4").hexdigest()[:16]


@needs_ref
def test_reference_perception_cache_loads_our_files(tmp_path):
    pc = _load_module(f"{REF}/benchmarks/perception_cache.py", "ref_perception_cache")
    video = tmp_path / "clip.mp4"
    video.write_bytes(os.urandom(4096))
    emb = torch.randn(7, 1152).to(torch.bfloat16)
    vm = torch.randn(3, 768)
    ts = np.arange(7, dtype=np.float64) / 2.0
    d = fc.write_perception_npz(str(tmp_path / "pc"), str(video), siglip=emb, videomae=vm, frame_indices=np.arange(7) * 30,
                                timestamps=ts, video_duration_sec=3.5)
    cache = pc.PerceptionCache(pc.CacheConfig(cache_dir=str(tmp_path / "pc")))
    assert cache.has_cache(str(video)) and str(cache._get_cache_path(cache.compute_video_hash(str(video)))) == d
    got = cache.load(str(video))
    assert got is not None and got.video_duration_sec == 3.5 and got.cache_version == pc.CACHE_VERSION
    assert np.array_equal(got.siglip_embeddings, emb.float().numpy())  # bit for bit (bf16 -> fp32 is exact)
    assert np.array_equal(got.videomae_embeddings, vm.numpy())
    assert np.array_equal(got.frame_indices, np.arange(7) * 30) and np.array_equal(got.frame_timestamps, ts)
    assert cache.get_cache_info(str(video))["has_siglip"] is True
    # and the other direction: what the reference saves, our reader loads
    feats = pc.CachedFeatures(video_hash="", video_path="", video_duration_sec=1.0, frame_indices=np.arange(2),
                              frame_timestamps=np.arange(2.0), siglip_embeddings=emb[:2].float().numpy())
    video2 = tmp_path / "b.mp4"
    video2.write_bytes(b"z" * 99)
    cache.save(str(video2), feats)
    back = fc.load_perception_npz(str(tmp_path / "pc"), str(video2))
    assert np.array_equal(back["siglip_embeddings"], emb[:2].float().numpy())
    assert set(back["metadata"]) == set(json.load(open(os.path.join(d, "metadata.json"))))


@needs_ref
def test_reference_training_reader_consumes_embeddings_pt(tmp_path):
    tp = _load_module(f"{REF}/scripts/train_projectors.py", "ref_train_projectors")
    emb = torch.randn(6, 1152).to(torch.bfloat16)
    vm = torch.randn(2, 768)
    recs = fc.siglip_embedding_records(np.arange(6.0), emb)
    vrecs = [{"start_time": 0.0, "end_time": 15.0, "embedding": vm[0], "source_frame_count": 16},
             {"start_time": 16.0, "end_time": 31.0, "embedding": vm[1], "source_frame_count": 16}]
    fc.write_embeddings_pt(str(tmp_path / "video1_embeddings.pt"), recs, vrecs)
    ds = tp.MultimodalProjectorDataset(str(tmp_path))  # scripts/train_projectors.py:190-279
    assert len(ds) == 8
    sig = [s for s in ds.samples if s["modality"] == "siglip"]
    vmae = [s for s in ds.samples if s["modality"] == "videomae"]
    assert len(sig) == 6 and len(vmae) == 2 and sig[3]["video"] == "video1" and sig[3]["timestamp"] == 3.0
    assert torch.equal(torch.stack([s["embedding"] for s in sig]), emb)
    assert torch.equal(torch.stack([s["embedding"] for s in vmae]), vm)
    batch = tp.collate_fn([ds[0], ds[1]]) if hasattr(tp, "collate_fn") else None
    assert batch is not None
    # scripts/demo_projector_inference.py:104-117
    data = torch.load(tmp_path / "video1_embeddings.pt", weights_only=False)
    assert torch.stack([e["embedding"] for e in data["siglip"][:5] if e.get("embedding") is not None]).shape == (5, 1152)
