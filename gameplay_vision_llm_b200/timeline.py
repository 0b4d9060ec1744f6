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
from .weights import resolve_device


class TimelineEmbeddingIndex:
    def __init__(self, n_items: int, dim: int = 4096, fps: float = 1.0, device: str | torch.device = "cuda",
                 rank: int = 0, world: int = 1, timestamps: Optional[np.ndarray] = None):
        self.n, self.dim, self.rank, self.world = int(n_items), int(dim), int(rank), int(world)
        self.device = torch.device(device)
        if self.device.type == "cuda":  # a CPU index exists only to assemble shards in the gloo tests; search needs CUDA
            self.device = resolve_device(self.device)
        self.per_rank = -(-self.n // self.world)
        self.lo, self.hi = shard_range(self.n, self.rank, self.world)
        # padded so every rank contributes per_rank rows to the gather
        self.embeddings = torch.zeros((self.per_rank * self.world, self.dim), dtype=torch.bfloat16, device=self.device)
        self.timestamps = (np.arange(self.n, dtype=np.float64) / float(fps)) if timestamps is None else np.asarray(
            timestamps, np.float64)
        self.gathered = self.world == 1
        self._inv_norm = None  # 1/|e_n| of the gathered index (cosine top-k, tensor path)

    def local_rows(self) -> torch.Tensor:
        """This rank's slice of the index (rows [lo, hi)), to be filled by EmbeddingPipeline.embed(out_index=...).
        Handing out a writable view invalidates the cached 1/|e_n| (streaming use: rows appended between searches)."""
        self._inv_norm = None
        return self.embeddings[self.lo:self.hi]

    def mark_dirty(self) -> None:
        """Call after writing rows through a view obtained earlier: the next search recomputes the cached norms."""
        self._inv_norm = None

    def all_gather(self) -> torch.Tensor:
        """One in-place all-gather of the (per_rank, dim) bf16 shards; afterwards every rank holds rows [0, n)."""
        if self.world > 1:
            import torch.distributed as dist
            shard = self.embeddings[self.rank * self.per_rank:(self.rank + 1) * self.per_rank]
            dist.all_gather_into_tensor(self.embeddings, shard)
        self.gathered = True
        self._inv_norm = None
        return self.embeddings[: self.n]

    def index(self) -> torch.Tensor:
        """Rows [0, n) — a view; callers that WRITE through it must call mark_dirty() before the next search."""
        return self.embeddings[: self.n]

    # ---- time filters: the index is in timestamp order, so a window is a contiguous row range ----------------
    def window_range(self, timestamp: float, window_sec: float = 30.0, inclusive: bool = False) -> tuple[int, int]:
        """[lo, hi) rows with |ts - timestamp| < window (inclusive=False, scripts/realtime_inference.py:994-998) or
        timestamp - window <= ts <= timestamp + window (inclusive=True, TimelineRetriever.retrieve_by_timestamp,
        src/agent_core/qwen_reasoning_core.py:1482-1490).  Binary search on the float64 timestamps; the comparison is
        evaluated exactly as the reference writes it at the two boundaries, so rounding of t +- w cannot move them."""
        ts, t, w = self.timestamps, float(timestamp), float(window_sec)
        if inclusive:
            inside = lambda x: t - w <= x <= t + w  # noqa: E731
        else:
            inside = lambda x: abs(x - t) < w  # noqa: E731
        lo = int(np.searchsorted(ts, t - w, side="left"))
        hi = int(np.searchsorted(ts, t + w, side="right"))
        lo, hi = max(0, lo - 1), min(self.n, hi + 1)
        while lo < hi and not inside(ts[lo]):
            lo += 1
        while hi > lo and not inside(ts[hi - 1]):
            hi -= 1
        return (lo, hi) if hi > lo else (0, 0)

    def search(self, queries: torch.Tensor, top_k: int = 16, eps: float = 1e-12, windows=None,
               inclusive: bool = False, mode: int = ops.TOPK_AUTO):
        """queries [Q, dim] -> (scores fp32 [Q,k], row indices int32 [Q,k]), score desc / index asc.
        windows: optional list of Q (timestamp, window_sec) pairs (None entries = whole timeline): query q only ranks
        the rows inside its window — filter and ranking run in one fused device call; slots beyond the number of rows
        in the window hold index -1."""
        if not self.gathered:
            raise RuntimeError("index is sharded: call all_gather() first")
        if not 1 <= int(top_k) <= 64:
            raise RuntimeError(f"search: top_k={top_k} out of range [1, 64] (the selection kernel's limit)")
        if self.n == 0 or queries.numel() == 0:
            # an empty timeline (or no query) has no neighbours: the reference returns [] here
            # (src/agent_core/qwen_reasoning_core.py:1508-1511), no device call is made
            nq = 0 if queries.numel() == 0 else queries.reshape(-1, self.dim).shape[0]
            return (torch.full((nq, top_k), float("-inf"), dtype=torch.float32, device=self.device),
                    torch.full((nq, top_k), -1, dtype=torch.int32, device=self.device))
        top_k = min(int(top_k), self.n)
        q = queries.to(self.device).to(torch.bfloat16).reshape(-1, self.dim).contiguous()
        if self._inv_norm is None:  # cached: the index does not change after the gather
            self._inv_norm = ops.row_inv_norm(self.index(), eps)
        if windows is None:
            return ops.topk_cosine(self.index(), q, top_k, eps=eps, mode=mode, inv_norm=self._inv_norm)
        if len(windows) != q.shape[0]:
            raise RuntimeError("search: one window (or None) per query")
        rng = np.array([(0, self.n) if w is None else self.window_range(w[0], w[1], inclusive) for w in windows], np.int32)
        lo = torch.as_tensor(rng[:, 0].copy(), device=self.device)
        hi = torch.as_tensor(rng[:, 1].copy(), device=self.device)
        span = (int(rng[:, 0].min()), int(rng[:, 1].max()))
        return ops.topk_cosine(self.index(), q, top_k, eps=eps, row_lo=lo, row_hi=hi, span=span, mode=mode,
                               inv_norm=self._inv_norm)

    def retrieve_by_semantic(self, query_embedding: torch.Tensor, top_k: int = 10) -> list[tuple[float, float]]:
        """[(timestamp, score)] best first — the embedding-index analogue of `TimelineRetriever.retrieve_by_semantic`
        (src/agent_core/qwen_reasoning_core.py:1492-1528, default semantic_top_k = 10, :655)."""
        scores, idx = self.search(query_embedding.reshape(1, -1), top_k)
        return [(float(self.timestamps[i]), float(s)) for s, i in zip(scores[0].tolist(), idx[0].tolist()) if i >= 0]

    def retrieve_by_timestamp(self, timestamp: float, window: float = 30.0) -> list[int]:
        """Rows with timestamp - window <= ts <= timestamp + window, in timeline order
        (`TimelineRetriever.retrieve_by_timestamp`, src/agent_core/qwen_reasoning_core.py:1462-1490;
        default retrieval_window_sec = 30)."""
        lo, hi = self.window_range(timestamp, window, inclusive=True)
        return list(range(lo, hi))

    def hybrid_retrieve(self, query_embedding: torch.Tensor, timestamp: float | None = None, window: float = 30.0,
                        semantic_top_k: int = 10, context_top_k: int = 5) -> list[int]:
        """Row indices in the order `TimelineRetriever.hybrid_retrieve` returns events
        (src/agent_core/qwen_reasoning_core.py:1530-1565): with a timestamp, the rows of the time window in timeline
        order followed by the semantic top-5 over the WHOLE timeline that are not among them; without one, the semantic
        top-k (default 10)."""
        if timestamp is None:
            _, idx = self.search(query_embedding.reshape(1, -1), semantic_top_k)
            return [i for i in idx[0].tolist() if i >= 0]
        rows = self.retrieve_by_timestamp(timestamp, window)
        _, idx = self.search(query_embedding.reshape(1, -1), context_top_k)
        seen = set(rows)
        for i in idx[0].tolist():
            if i >= 0 and i not in seen:
                rows.append(i)
                seen.add(i)
        return rows

    def window(self, timestamp: float, window_sec: float = 30.0, fallback_first: int = 20):
        """Rows whose |ts - timestamp| < window and their embeddings — the `siglip_near` selection of
        scripts/realtime_inference.py:988-1003 (timestamp None there = the first 20 rows)."""
        if timestamp is None:
            lo, hi = 0, min(fallback_first, self.n)
        else:
            lo, hi = self.window_range(timestamp, window_sec, inclusive=False)
        return np.arange(lo, hi), self.embeddings[lo:hi]
