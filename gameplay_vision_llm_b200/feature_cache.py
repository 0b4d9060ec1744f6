"""On-disk contracts around the hot path, host-side Python like the reference:

* `FeatureCache`          — `src/agent_core/qwen_reasoning_core.py:355-528`: `cache_dir/{stem}_{sha256(str(size) +
                            first 1 MiB + last 1 MiB)[:16]}.pt`, in-memory layer, oldest-first eviction down to
                            80 % of the size limit, corrupt files unlinked.
* `write_embeddings_pt`   — `scripts/extract_features.py:1431-1468`: `{video}_embeddings.pt`, read back by
                            `scripts/train_projectors.py:206-255` and `scripts/demo_projector_inference.py:91-124`.
* `siglip_cache_entries`  — payload `realtime_inference.py:868-878` stores under "siglip".
* `write_perception_npz`  — `benchmarks/perception_cache.py:203-283`: `siglip.npz` key `embeddings` (N,1152) etc.
"""
from __future__ import annotations

import hashlib
import json
import logging
import os
from pathlib import Path
from typing import Optional, Sequence

import numpy as np
import torch

logger = logging.getLogger(__name__)
_MIB = 1024 * 1024


class FeatureCache:
    def __init__(self, cache_dir: str = "data/cache", max_cache_size_gb: float = 10.0):
        self.cache_dir = Path(cache_dir)
        self.cache_dir.mkdir(parents=True, exist_ok=True)
        self.max_cache_size_bytes = int(max_cache_size_gb * 1024 * 1024 * 1024)
        self._memory_cache: dict[str, dict] = {}

    def _get_video_hash(self, video_path: str) -> str:
        path = Path(video_path)
        if not path.exists():
            return ""
        size = path.stat().st_size
        chunk = min(_MIB, size)
        h = hashlib.sha256()
        h.update(str(size).encode())
        with open(path, "rb") as f:
            h.update(f.read(chunk))
            if size > chunk * 2:
                f.seek(-chunk, 2)
                h.update(f.read(chunk))
        return h.hexdigest()[:16]

    def _get_cache_path(self, video_path: str) -> Path:
        return self.cache_dir / f"{Path(video_path).stem}_{self._get_video_hash(video_path)}.pt"

    def has_features(self, video_path: str) -> bool:
        return self._get_video_hash(video_path) in self._memory_cache or self._get_cache_path(video_path).exists()

    def load_features(self, video_path: str) -> Optional[dict]:
        key = self._get_video_hash(video_path)
        if key in self._memory_cache:
            return self._memory_cache[key]
        path = self._get_cache_path(video_path)
        if path.exists():
            try:
                features = torch.load(path, map_location="cpu", weights_only=False)
            except Exception as exc:  # corrupt file: drop it, like the reference (:462-464)
                logger.warning("Failed to load cache: %s", exc)
                path.unlink(missing_ok=True)
                return None
            self._memory_cache[key] = features
            return features
        return None

    def save_features(self, video_path: str, features: dict) -> None:
        self._memory_cache[self._get_video_hash(video_path)] = features
        try:
            self._cleanup_old_cache()
            torch.save(features, self._get_cache_path(video_path))
        except Exception as exc:
            logger.warning("Failed to save cache: %s", exc)

    def _cleanup_old_cache(self) -> None:
        files = list(self.cache_dir.glob("*.pt"))
        total = sum(f.stat().st_size for f in files)
        if total <= self.max_cache_size_bytes:
            return
        files.sort(key=lambda f: f.stat().st_mtime)
        while total > self.max_cache_size_bytes * 0.8 and files:
            oldest = files.pop(0)
            total -= oldest.stat().st_size
            oldest.unlink()

    def clear(self) -> None:
        self._memory_cache.clear()
        for f in self.cache_dir.glob("*.pt"):
            f.unlink()

    def get_cache_stats(self) -> dict:
        files = list(self.cache_dir.glob("*.pt"))
        return {"num_cached_videos": len(files), "total_size_mb": sum(f.stat().st_size for f in files) / _MIB,
                "memory_cache_entries": len(self._memory_cache)}


def siglip_embedding_records(timestamps: Sequence[float], embeddings: torch.Tensor) -> list[dict]:
    """The list `run_siglip_encoder` builds (scripts/extract_features.py:597-603): one dict per frame with a CPU
    tensor in the dtype the encoder produced (bf16)."""
    emb = embeddings.detach().cpu()
    return [{"timestamp": float(ts), "embedding": emb[i], "embedding_shape": list(emb[i].shape),
             "entity_type": "full_frame", "description": f"Full frame embedding at {float(ts):.1f}s"}
            for i, ts in enumerate(timestamps)]


def siglip_cache_entries(timestamps: Sequence[float], embeddings: torch.Tensor) -> list[dict]:
    """The "siglip" entry of the FeatureCache payload (scripts/realtime_inference.py:316-322, 868-878)."""
    emb = embeddings.detach().cpu()
    return [{"timestamp": float(ts), "label": "full_frame", "embedding": emb[i]} for i, ts in enumerate(timestamps)]


def write_embeddings_pt(path: str, siglip_records: Sequence[dict], videomae_records: Sequence[dict] = (),
                        extra: Optional[dict] = None) -> dict:
    """`{video}_embeddings.pt` (scripts/extract_features.py:1431-1468)."""
    data = {
        "siglip": [{"timestamp": r["timestamp"], "embedding": r["embedding"],
                    "shape": r.get("embedding_shape", list(r["embedding"].shape))} for r in siglip_records],
        "videomae": [{"start_time": r["start_time"], "end_time": r["end_time"], "embedding": r["embedding"],
                      "source_frame_count": r.get("source_frame_count", 16)} for r in videomae_records],
        "wav2vec2": [], "hico": [], "visual_events": [], "audio_transcripts": [],
    }
    if extra:
        data.update(extra)
    torch.save(data, path)
    return data


PERCEPTION_CACHE_VERSION = "1.0.0"  # benchmarks/perception_cache.py:42


def perception_video_hash(video_path: str) -> str:
    """`PerceptionCache.compute_video_hash` (benchmarks/perception_cache.py:166-181): md5 of `path:size:mtime`, or of
    the bare path string when the file does not exist; first 16 hex digits."""
    if not Path(video_path).exists():
        return hashlib.md5(video_path.encode()).hexdigest()[:16]
    st = Path(video_path).stat()
    return hashlib.md5(f"{video_path}:{st.st_size}:{st.st_mtime}".encode()).hexdigest()[:16]


def _np_embeddings(t) -> Optional[np.ndarray]:
    """numpy has no bf16: device / bf16 rows are stored as fp32 (every bf16 value is exact in fp32)."""
    if t is None:
        return None
    if torch.is_tensor(t):
        t = t.detach().cpu()
        return (t.float() if t.dtype in (torch.bfloat16, torch.float16) else t).numpy()
    return np.asarray(t)


def write_perception_npz(cache_dir: str, video_path: str, siglip=None, videomae=None,
                         frame_indices: Optional[np.ndarray] = None, timestamps: Optional[np.ndarray] = None,
                         video_duration_sec: float = 0.0, perception_config: Optional[dict] = None) -> str:
    """What `PerceptionCache.save` leaves on disk for the embedding part of `CachedFeatures`
    (benchmarks/perception_cache.py:203-283), so that the reference's own `PerceptionCache.load` (:285-372) reads it:

      cache_dir/{hash}/metadata.json   the keys of `CachedFeatures.to_metadata()` (:99-116)
      cache_dir/{hash}/frames.npz      `indices`, `timestamps`           (only with frame_indices, :232-237)
      cache_dir/{hash}/siglip.npz      `embeddings` (N, 1152)
      cache_dir/{hash}/videomae.npz    `embeddings` (N, 768)
      cache_dir/index.json             {hash: metadata} — `load` returns None for a hash that is not listed (:297)
    """
    import time
    vid = perception_video_hash(video_path)
    root = Path(cache_dir)
    d = root / vid
    d.mkdir(parents=True, exist_ok=True)
    sig, vmae = _np_embeddings(siglip), _np_embeddings(videomae)
    if timestamps is not None and frame_indices is None:
        frame_indices = np.arange(len(timestamps), dtype=np.int64)
    metadata = {
        "video_hash": vid,
        "video_path": video_path,
        "video_duration_sec": float(video_duration_sec),
        "cache_version": PERCEPTION_CACHE_VERSION,
        "cached_at": time.strftime("%Y-%m-%dT%H:%M:%SZ"),
        "perception_config": dict(perception_config or {}),
        "has_siglip": sig is not None,
        "has_videomae": vmae is not None,
        "has_sam": False, "has_ocr": False, "has_audio": False, "has_timeline": False, "has_kb": False,
        "num_frames": int(len(frame_indices)) if frame_indices is not None else 0,
    }
    with open(d / "metadata.json", "w") as f:
        json.dump(metadata, f, indent=2)
    if frame_indices is not None:
        np.savez_compressed(d / "frames.npz", indices=np.asarray(frame_indices),
                            timestamps=np.asarray(timestamps if timestamps is not None else [], dtype=np.float64))
    if sig is not None:
        np.savez_compressed(d / "siglip.npz", embeddings=sig)
    if vmae is not None:
        np.savez_compressed(d / "videomae.npz", embeddings=vmae)
    index_path = root / "index.json"
    index: dict = {}
    if index_path.exists():
        try:
            with open(index_path) as f:
                index = json.load(f)
        except Exception as exc:  # same tolerance as `_load_index` (:148-156)
            logger.warning("Failed to load cache index: %s", exc)
            index = {}
    index[vid] = metadata
    with open(index_path, "w") as f:
        json.dump(index, f, indent=2)
    return str(d)


def load_perception_npz(cache_dir: str, video_path: str) -> Optional[dict]:
    """Reader twin of `PerceptionCache.load` for the embedding part (benchmarks/perception_cache.py:285-346):
    None unless the hash is listed in index.json and its directory + metadata.json exist."""
    root = Path(cache_dir)
    vid = perception_video_hash(video_path)
    try:
        with open(root / "index.json") as f:
            index = json.load(f)
    except Exception:
        return None
    d = root / vid
    if vid not in index or not (d / "metadata.json").exists():
        return None
    with open(d / "metadata.json") as f:
        out = {"metadata": json.load(f), "frame_indices": None, "frame_timestamps": None, "siglip_embeddings": None,
               "videomae_embeddings": None}
    if (d / "frames.npz").exists():
        z = np.load(d / "frames.npz")
        out["frame_indices"], out["frame_timestamps"] = z["indices"], z["timestamps"]
    if (d / "siglip.npz").exists():
        out["siglip_embeddings"] = np.load(d / "siglip.npz")["embeddings"]
    if (d / "videomae.npz").exists():
        out["videomae_embeddings"] = np.load(d / "videomae.npz")["embeddings"]
    return out
