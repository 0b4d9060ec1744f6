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


def write_perception_npz(cache_dir: str, video_path: str, siglip: Optional[torch.Tensor] = None,
                         videomae: Optional[torch.Tensor] = None, frame_indices: Optional[np.ndarray] = None,
                         timestamps: Optional[np.ndarray] = None) -> str:
    """`PerceptionCache` directory (benchmarks/perception_cache.py:11-24, 203-283): `{md5(path:size:mtime)[:16]}/`
    with `siglip.npz` / `videomae.npz` (key `embeddings`, fp32 — numpy has no bf16), `frames.npz`, `metadata.json`."""
    st = os.stat(video_path)
    vid = hashlib.md5(f"{video_path}:{st.st_size}:{st.st_mtime}".encode()).hexdigest()[:16]
    d = Path(cache_dir) / vid
    d.mkdir(parents=True, exist_ok=True)
    if siglip is not None:
        np.savez_compressed(d / "siglip.npz", embeddings=siglip.detach().float().cpu().numpy())
    if videomae is not None:
        np.savez_compressed(d / "videomae.npz", embeddings=videomae.detach().float().cpu().numpy())
    if frame_indices is not None or timestamps is not None:
        np.savez_compressed(d / "frames.npz", indices=np.asarray(frame_indices if frame_indices is not None else []),
                            timestamps=np.asarray(timestamps if timestamps is not None else []))
    with open(d / "metadata.json", "w") as f:
        json.dump({"video_path": video_path, "video_id": vid, "cache_version": "1.0.0",
                   "num_frames": int(siglip.shape[0]) if siglip is not None else 0}, f)
    return str(d)
