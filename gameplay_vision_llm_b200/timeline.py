"""The timeline index the projected embeddings land in, and its retrieval.

The reference keeps embeddings in a Python list of `{"timestamp","label","embedding"}` dicts filtered by
±30 s (scripts/realtime_inference.py:988-1003) and ranks text events by cosine similarity with
`argsort(descending)[:top_k]` (src/agent_core/qwen_reasoning_core.py:1492-1528).  Here the index is one
(N, llm_dim) bf16 tensor in timestamp order plus float64 timestamps; retrieval is the cosine top-k kernel.
Across GPUs each rank fills the rows of its contiguous chunk and one NCCL all-gather (NVLink/NVSwitch)
gives every rank the whole index.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .pipeline import shard_range


class TimelineEmbeddingIndex:
    def __init__(self, n_items: int, dim: int = 4096, fps: float = 1.0, device: str | torch.device = "cuda:0",
                 rank: int = 0, world: int = 1, timestamps: Optional[np.ndarray] = None):
        self.n, self.dim, self.rank, self.world = int(n_items), int(dim), int(rank), int(world)
        self.device = torch.device(device)
        self.per_rank = -(-self.n // self.world)
        self.lo, self.hi = shard_range(self.n, self.rank, self.world)
        # padded so every rank contributes per_rank rows to the gather
        self.embeddings = torch.zeros((self.per_rank * self.world, self.dim), dtype=torch.bfloat16, device=self.device)
        self.timestamps = (np.arange(self.n, dtype=np.float64) / float(fps)) if timestamps is None else np.asarray(
            timestamps, np.float64)
        self.gathered = self.world == 1

    def local_rows(self) -> torch.Tensor:
        """This rank's slice of the index (rows [lo, hi)), to be filled by EmbeddingPipeline.embed(out_index=...)."""
        return self.embeddings[self.lo:self.hi]

    def all_gather(self) -> torch.Tensor:
        """One in-place all-gather of the (per_rank, dim) bf16 shards; afterwards every rank holds rows [0, n)."""
        if self.world > 1:
            import torch.distributed as dist
            shard = self.embeddings[self.rank * self.per_rank:(self.rank + 1) * self.per_rank]
            dist.all_gather_into_tensor(self.embeddings, shard)
        self.gathered = True
        return self.embeddings[: self.n]

    def index(self) -> torch.Tensor:
        return self.embeddings[: self.n]

    def search(self, queries: torch.Tensor, top_k: int = 16, eps: float = 1e-12):
        """queries [Q, dim] -> (scores fp32 [Q,k], row indices int32 [Q,k]), score desc / index asc."""
        if not self.gathered:
            raise RuntimeError("index is sharded: call all_gather() first")
        q = queries.to(self.device).to(torch.bfloat16).reshape(-1, self.dim).contiguous()
        return ops.topk_cosine(self.index(), q, min(top_k, self.n), eps=eps)

    def retrieve_by_semantic(self, query_embedding: torch.Tensor, top_k: int = 10) -> list[tuple[float, float]]:
        """[(timestamp, score)] best first — the embedding-index analogue of `TimelineRetriever.retrieve_by_semantic`."""
        scores, idx = self.search(query_embedding.reshape(1, -1), top_k)
        return [(float(self.timestamps[i]), float(s)) for s, i in zip(scores[0].tolist(), idx[0].tolist()) if i >= 0]

    def window(self, timestamp: float, window_sec: float = 30.0, fallback_first: int = 20):
        """Rows whose |ts - timestamp| < window (scripts/realtime_inference.py:988-998; first 20 if none)."""
        sel = np.nonzero(np.abs(self.timestamps - float(timestamp)) < float(window_sec))[0]
        if sel.size == 0:
            sel = np.arange(min(fallback_first, self.n))
        return sel, self.embeddings[torch.as_tensor(sel, device=self.device)]
