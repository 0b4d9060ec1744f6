"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of the arithmetic the reference executes for one frame after preprocessing:
`self.encoder._model.get_image_features(**inputs)` (reference: src/perception/
siglip_semantic_encoder.py:479-481), `ProjectorBank.project_region`
(src/agent_core/qwen_reasoning_core.py:1076 -> MultiModalProjector :1007-1027) and the cosine top-k of
`TimelineRetriever.retrieve_by_semantic` (:1492-1528) / `find_similar_regions`
(src/perception/siglip_semantic_encoder.py:616-638).

The vision-tower arithmetic lives in HuggingFace transformers (pinned inconsistently by the
reference: pyproject.toml:14 git ff13eb66…, uv.lock:4686 4.57.3; 5.5.0 is installed here and is
what this file restates): models/siglip/modeling_siglip.py
  :116-186 SiglipVisionEmbeddings   :229-249 eager attention   :252-312 SiglipAttention
  :315-327 SiglipMLP (gelu_pytorch_tanh)   :330-361 SiglipEncoderLayer
  :604-625 SiglipVisionTransformer.forward   :628-649 SiglipMultiheadAttentionPoolingHead.

Written with explicit tensor ops (matmul / softmax / mean / var) on a plain state-dict — no HF
modules — and pinned against the HF modules themselves by tests/test_oracle_cpu.py and the golden
vectors under tests/golden/ (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _ln(x, g, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _mha(x, wq, bq, wk, bk, wv, bv, heads):
    """softmax(q k^T / sqrt(hd)) v per head; x [B,T,D] -> [B,T,D] (before out-proj)."""
    B, T, D = x.shape
    hd = D // heads
    q = (x @ wq.T + bq).view(B, T, heads, hd).transpose(1, 2)
    k = (x @ wk.T + bk).view(B, T, heads, hd).transpose(1, 2)
    v = (x @ wv.T + bv).view(B, T, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, T, D)


def vision_forward(sd: dict, pixel_values: torch.Tensor, heads: int, patch: int, eps: float = 1e-6,
                   act=gelu_tanh, seams: dict | None = None, dtype=torch.float32, pos: torch.Tensor | None = None) -> torch.Tensor:
    """pixel_values [B,3,h,w] -> pooled [B,D].  `seams` (if given) receives the intermediate tensors.  `pos`
    replaces the learned position table (the `interpolate_pos_encoding=True` route for h x w inputs whose patch
    grid differs from the checkpoint's, HF :137-186; see oracle/region_ref.py)."""
    pre = "vision_model." if any(k.startswith("vision_model.") for k in sd) else ""
    W = lambda n: sd[pre + n].to(dtype)  # noqa: E731
    x = pixel_values.to(dtype)
    B = x.shape[0]
    wp = W("embeddings.patch_embedding.weight")  # [D,3,p,p]
    D = wp.shape[0]
    gh, gw = x.shape[-2] // patch, x.shape[-1] // patch
    # valid stride-p conv == im2col GEMM over the top-left gh*p x gw*p pixels
    cols = x[:, :, : gh * patch, : gw * patch].reshape(B, 3, gh, patch, gw, patch).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, -1)
    h = cols @ wp.reshape(D, -1).T + W("embeddings.patch_embedding.bias")
    h = h + (W("embeddings.position_embedding.weight") if pos is None else pos.to(dtype))[None]
    if seams is not None:
        seams["patches"] = cols
        seams["embeddings"] = h
    n_layers = 1 + max(int(k[len(pre):].split(".")[2]) for k in sd if k.startswith(pre + "encoder.layers."))
    for i in range(n_layers):
        p = f"encoder.layers.{i}."
        y = _ln(h, W(p + "layer_norm1.weight"), W(p + "layer_norm1.bias"), eps)
        a = _mha(y, W(p + "self_attn.q_proj.weight"), W(p + "self_attn.q_proj.bias"),
                 W(p + "self_attn.k_proj.weight"), W(p + "self_attn.k_proj.bias"),
                 W(p + "self_attn.v_proj.weight"), W(p + "self_attn.v_proj.bias"), heads)
        h = h + (a @ W(p + "self_attn.out_proj.weight").T + W(p + "self_attn.out_proj.bias"))
        y = _ln(h, W(p + "layer_norm2.weight"), W(p + "layer_norm2.bias"), eps)
        y = act(y @ W(p + "mlp.fc1.weight").T + W(p + "mlp.fc1.bias"))
        h = h + (y @ W(p + "mlp.fc2.weight").T + W(p + "mlp.fc2.bias"))
        if seams is not None:
            seams[f"layer{i}"] = h
    tokens = _ln(h, W("post_layernorm.weight"), W("post_layernorm.bias"), eps)
    if seams is not None:
        seams["last_hidden_state"] = tokens
    # MAP head: nn.MultiheadAttention(probe, tokens, tokens) with packed in_proj = [Wq; Wk; Wv]
    ipw, ipb = W("head.attention.in_proj_weight"), W("head.attention.in_proj_bias")
    hd = D // heads
    probe = W("head.probe").reshape(1, 1, D).expand(B, 1, D)
    q = (probe @ ipw[:D].T + ipb[:D]).view(B, 1, heads, hd).transpose(1, 2)
    k = (tokens @ ipw[D:2 * D].T + ipb[D:2 * D]).view(B, -1, heads, hd).transpose(1, 2)
    v = (tokens @ ipw[2 * D:].T + ipb[2 * D:]).view(B, -1, heads, hd).transpose(1, 2)
    p_att = torch.softmax((q * (hd ** -0.5)) @ k.transpose(-1, -2), dim=-1)
    a = (p_att @ v).transpose(1, 2).reshape(B, 1, D)
    a = a @ W("head.attention.out_proj.weight").T + W("head.attention.out_proj.bias")
    if seams is not None:
        seams["head_attn"] = a[:, 0]
    y = _ln(a, W("head.layernorm.weight"), W("head.layernorm.bias"), eps)
    y = act(y @ W("head.mlp.fc1.weight").T + W("head.mlp.fc1.bias"))
    out = a + (y @ W("head.mlp.fc2.weight").T + W("head.mlp.fc2.bias"))
    return out[:, 0]


def projector_forward(psd: dict, x: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """MultiModalProjector: Linear -> nn.GELU() (exact erf) -> Linear."""
    h = gelu_erf(x.to(dtype) @ psd["net.0.weight"].to(dtype).T + psd["net.0.bias"].to(dtype))
    return h @ psd["net.2.weight"].to(dtype).T + psd["net.2.bias"].to(dtype)


def cosine_topk(index: np.ndarray, queries: np.ndarray, k: int, eps: float = 1e-12):
    """float64 cosine scores + the (score desc, index asc) total order.

    sentence_transformers.util.cos_sim == normalize(a) @ normalize(b).T with eps 1e-12, then
    `argsort(descending=True)[:k]` (qwen_reasoning_core.py:1521-1528); ties resolved like Python's
    stable sort in find_similar_regions (siglip_semantic_encoder.py:633-638): lower index first.
    Returns (scores [Q,k] float64, idx [Q,k] int64, margins [Q] = score[k-1] - score[k] (inf if N<=k)).
    """
    e = np.asarray(index, np.float64)
    q = np.asarray(queries, np.float64)
    en = e / np.maximum(np.linalg.norm(e, axis=1, keepdims=True), eps)
    qn = q / np.maximum(np.linalg.norm(q, axis=1, keepdims=True), eps)
    s = qn @ en.T
    N = s.shape[1]
    kk = min(k, N)
    idx = np.empty((s.shape[0], kk), np.int64)
    top = np.empty((s.shape[0], kk), np.float64)
    margins = np.full(s.shape[0], np.inf)
    for i in range(s.shape[0]):
        order = np.lexsort((np.arange(N), -s[i]))  # primary: -score, secondary: index
        idx[i] = order[:kk]
        top[i] = s[i, order[:kk]]
        if N > kk:
            margins[i] = s[i, order[kk - 1]] - s[i, order[kk]]
    return top, idx, margins


def window_rows(timestamps: np.ndarray, center: float, window: float, inclusive: bool):
    """Rows of a timestamp-ordered timeline inside the reference's time filters, as a [lo, hi) range.

    inclusive=False: `abs(ts - center) < window` (scripts/realtime_inference.py:994-998);
    inclusive=True:  `center - window <= ts <= center + window` (src/agent_core/qwen_reasoning_core.py:1482-1490,
    and TimelineIndexer.query_range `start <= ts <= end`, src/fusion_indexing/timeline_indexer.py:588-614).
    Restated literally (a Python scan), not with searchsorted."""
    ts = np.asarray(timestamps, np.float64)
    if inclusive:
        sel = [i for i, t in enumerate(ts) if center - window <= t <= center + window]
    else:
        sel = [i for i, t in enumerate(ts) if abs(t - center) < window]
    if not sel:
        return 0, 0
    assert sel == list(range(sel[0], sel[-1] + 1)), "timestamps must be sorted"
    return sel[0], sel[-1] + 1


def cosine_topk_windowed(index: np.ndarray, queries: np.ndarray, k: int, row_lo, row_hi, eps: float = 1e-12):
    """cosine_topk restricted per query to rows [row_lo[q], row_hi[q]); missing slots are idx -1 / score -inf."""
    Q = np.asarray(queries).shape[0]
    top = np.full((Q, k), -np.inf)
    idx = np.full((Q, k), -1, np.int64)
    for q in range(Q):
        lo, hi = int(row_lo[q]), int(row_hi[q])
        if hi <= lo:
            continue
        s, i, _ = cosine_topk(np.asarray(index)[lo:hi], np.asarray(queries)[q:q + 1], k, eps)
        top[q, : s.shape[1]] = s[0]
        idx[q, : i.shape[1]] = i[0] + lo
    return top, idx


def hybrid_merge(time_rows: list, semantic_rows: list) -> list:
    """TimelineRetriever.hybrid_retrieve's merge (src/agent_core/qwen_reasoning_core.py:1548-1561): the time-window
    events in timeline order, then the semantic hits that are not already present, in rank order."""
    out = list(time_rows)
    seen = set(out)
    for r in semantic_rows:
        if r not in seen:
            out.append(r)
    return out
