"""Model specs, seeded synthetic weights in the HF state-dict layout, and the repack into the device
weight pack `gvl_vit_weights` of include/gvl.h.

The reference loads `google/siglip2-so400m-patch14-384` with `AutoModel.from_pretrained(...,
torch_dtype=bfloat16)` (src/perception/siglip_semantic_encoder.py:195-204).  There is no network here,
so benchmarks and parity tests use random weights *in the same state-dict layout* (SURVEY.md §8a row
M1); a real checkpoint's `state_dict()` goes through the same `pack_siglip`.

All synthetic values are rounded to bf16 (then stored as fp32), i.e. they are exactly the numbers a
bf16 checkpoint holds; the oracle computes with them in fp32, the device path stores matrices as bf16
and vectors (biases, LayerNorm, probe query) as fp32 holding the same bf16-representable values.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import torch

from . import _lib


def resolve_device(device) -> torch.device:
    """'cuda' -> the process's CURRENT device (what the reference's `device="cuda"` means under torchrun with
    `set_device(local_rank)`), 'cuda:N' -> itself; anything else raises (no CPU fallback)."""
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(f"this path runs on a CUDA (sm_100a) device only, got '{device}': there is no CPU fallback")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


@dataclass(frozen=True)
class SiglipVisionSpec:
    hidden: int = 1152
    intermediate: int = 4304
    layers: int = 27
    heads: int = 16
    image: int = 384
    patch: int = 14
    eps: float = 1e-6
    act: str = "gelu_pytorch_tanh"

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid

    @property
    def patch_k(self) -> int:
        return 3 * self.patch * self.patch

    @property
    def patch_ld(self) -> int:
        """patch row length padded to a multiple of 8 elements (16-byte TMA row stride): 588 -> 592."""
        return (self.patch_k + 7) // 8 * 8

    @staticmethod
    def so400m() -> "SiglipVisionSpec":
        """SigLIP2-so400m-patch14-384 vision tower (428 225 600 parameters)."""
        return SiglipVisionSpec()

    @staticmethod
    def tiny() -> "SiglipVisionSpec":
        """Small tower with the same awkward shapes (head dim 72, ragged tiles) for fast parity tests."""
        return SiglipVisionSpec(hidden=144, intermediate=304, layers=2, heads=2, image=56, patch=14)

    def flops_per_frame(self, with_projector_llm: int | None = 4096) -> int:
        """Algorithmic FLOPs (2*M*N*K, no padding) — BASELINE.md §3."""
        T, D, I = self.tokens, self.hidden, self.intermediate
        patch = 2 * T * D * self.patch_k
        qkv = 2 * T * 3 * D * D
        out = 2 * T * D * D
        attn = 2 * 2 * T * T * D
        mlp = 2 * 2 * T * D * I
        # MAP head: K/V projections, probe q projection, probe attention, out-proj, MLP
        head = 2 * T * 2 * D * D + 2 * D * D + 2 * 2 * T * D + 2 * D * D + 2 * 2 * D * I
        total = patch + self.layers * (qkv + out + attn + mlp) + head
        if with_projector_llm:
            total += 2 * D * with_projector_llm + 2 * with_projector_llm * with_projector_llm
        return total


def _bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def synth_siglip_state_dict(spec: SiglipVisionSpec, seed: int = 0) -> dict[str, torch.Tensor]:
    """Random vision-tower weights under the HF `SiglipVisionModel`/`SiglipModel` names
    (HF: models/siglip/modeling_siglip.py).  Deterministic on any machine (CPU generator).

    Scales follow HF's `_init_weights` in spirit (lecun-normal matrices, 1/sqrt(D) position
    embedding), except that biases and LayerNorm affine parameters are perturbed so a kernel that
    dropped one of them cannot pass, and q/k get gain 1.5 so the softmax is not near-uniform.
    """
    g = torch.Generator().manual_seed(int(seed))
    D, I, P = spec.hidden, spec.intermediate, spec.patch
    sd: dict[str, torch.Tensor] = {}

    def normal(*shape, std):
        return _bf16_round(torch.randn(*shape, generator=g) * std)

    def linear(prefix, out_f, in_f, gain=1.0):
        sd[prefix + ".weight"] = normal(out_f, in_f, std=gain / math.sqrt(in_f))
        sd[prefix + ".bias"] = normal(out_f, std=0.02)

    def layernorm(prefix):
        sd[prefix + ".weight"] = _bf16_round(1.0 + 0.1 * torch.randn(D, generator=g))
        sd[prefix + ".bias"] = normal(D, std=0.05)

    vm = "vision_model."
    sd[vm + "embeddings.patch_embedding.weight"] = normal(D, 3, P, P, std=1.0 / math.sqrt(3 * P * P))
    sd[vm + "embeddings.patch_embedding.bias"] = normal(D, std=0.02)
    sd[vm + "embeddings.position_embedding.weight"] = normal(spec.tokens, D, std=1.0 / math.sqrt(D))
    for i in range(spec.layers):
        p = f"{vm}encoder.layers.{i}."
        layernorm(p + "layer_norm1")
        linear(p + "self_attn.q_proj", D, D, gain=1.5)
        linear(p + "self_attn.k_proj", D, D, gain=1.5)
        linear(p + "self_attn.v_proj", D, D)
        linear(p + "self_attn.out_proj", D, D)
        layernorm(p + "layer_norm2")
        linear(p + "mlp.fc1", I, D)
        linear(p + "mlp.fc2", D, I)
    layernorm(vm + "post_layernorm")
    h = vm + "head."
    sd[h + "probe"] = normal(1, 1, D, std=1.0)
    sd[h + "attention.in_proj_weight"] = normal(3 * D, D, std=1.5 / math.sqrt(D))
    sd[h + "attention.in_proj_bias"] = normal(3 * D, std=0.02)
    linear(h + "attention.out_proj", D, D)
    layernorm(h + "layernorm")
    linear(h + "mlp.fc1", I, D)
    linear(h + "mlp.fc2", D, I)
    return sd


def synth_projector_state_dict(encoder_dim: int, llm_dim: int = 4096, seed: int = 1) -> dict[str, torch.Tensor]:
    """Random `MultiModalProjector` weights under the reference's names `net.0.*`, `net.2.*`
    (src/agent_core/qwen_reasoning_core.py:1009-1013)."""
    g = torch.Generator().manual_seed(int(seed))
    return {
        "net.0.weight": _bf16_round(torch.randn(llm_dim, encoder_dim, generator=g) / math.sqrt(encoder_dim)),
        "net.0.bias": _bf16_round(torch.randn(llm_dim, generator=g) * 0.02),
        "net.2.weight": _bf16_round(torch.randn(llm_dim, llm_dim, generator=g) / math.sqrt(llm_dim)),
        "net.2.bias": _bf16_round(torch.randn(llm_dim, generator=g) * 0.02),
    }


def synth_ren_projection_state_dict(dim: int, seed: int = 3) -> dict[str, torch.Tensor]:
    """Random weights for the REN-style `projection` head of `SigLIPSemanticEncoder` under the reference's
    nn.Sequential names "0.*" / "2.*" (src/perception/siglip_semantic_encoder.py:416-420); bf16-representable."""
    sd = synth_projector_state_dict(dim, dim, seed)
    return {k[len("net."):]: v for k, v in sd.items()}


_ACT = {"gelu_pytorch_tanh": 1, "gelu": 2, "gelu_erf": 2, "none": 0}


class SiglipPack:
    """Device-resident weights + the ctypes `gvl_vit_weights` struct pointing at them."""

    def __init__(self, sd: dict[str, torch.Tensor], spec: SiglipVisionSpec, device: torch.device | str,
                 fold_ln: bool = False):
        """fold_ln: fold every token-level LayerNorm into the Linear that consumes it (include/gvl.h,
        `gvl_gemm_fusion`): W' = bf16(W * gamma), c1[n] = sum_k W'[n,k], c2 = bias + W @ beta.  The LayerNorm
        kernels then disappear from the forward pass.  Measured on B200 (profiles/README.md): the folded epilogues cost
        +18 us (qkv), +20 us (fc1), +6 / +9 us (the two producers) per layer against 2 x 44 us of LayerNorm kernels.
        EmbeddingPipeline turns it on; the default here keeps the reference's op order."""
        self.spec = spec
        self.device = resolve_device(device)
        self._keep: list[torch.Tensor] = []
        D, I, T = spec.hidden, spec.intermediate, spec.tokens
        vm = "vision_model." if any(k.startswith("vision_model.") for k in sd) else ""

        def folded(weight: torch.Tensor, bias: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
            wf = (weight.double() * gamma.double()[None, :]).to(torch.bfloat16)
            c1 = wf.double().sum(dim=1).to(torch.float32)
            c2 = (bias.double() + weight.double() @ beta.double()).to(torch.float32)
            return wf, c1, c2

        def get(name):
            return sd[vm + name].detach().to(torch.float32)

        def mat(t: torch.Tensor) -> torch.Tensor:  # bf16 matrix on device
            d = t.to(torch.bfloat16).contiguous().to(self.device)
            self._keep.append(d)
            return d

        def vec(t: torch.Tensor) -> torch.Tensor:  # fp32 vector on device
            d = t.to(torch.float32).contiguous().to(self.device)
            self._keep.append(d)
            return d

        w = _lib.VitWeights()
        w.D, w.I, w.H, w.hd, w.L, w.T = D, I, spec.heads, spec.head_dim, spec.layers, T
        w.patch_k, w.patch_ld = spec.patch_k, spec.patch_ld
        w.eps = spec.eps
        w.act = _ACT[spec.act]

        wp = torch.zeros(D, spec.patch_ld)
        wp[:, : spec.patch_k] = get("embeddings.patch_embedding.weight").reshape(D, spec.patch_k)
        w.w_patch = mat(wp).data_ptr()
        w.b_patch = vec(get("embeddings.patch_embedding.bias")).data_ptr()
        self.pos_table = mat(get("embeddings.position_embedding.weight"))  # bf16 [T, D]
        w.pos = self.pos_table.data_ptr()

        self._layers = (_lib.VitLayer * spec.layers)()
        for i in range(spec.layers):
            p = f"encoder.layers.{i}."
            ly = self._layers[i]
            ly.ln1_g = vec(get(p + "layer_norm1.weight")).data_ptr()
            ly.ln1_b = vec(get(p + "layer_norm1.bias")).data_ptr()
            wqkv = torch.cat([get(p + f"self_attn.{n}_proj.weight") for n in "qkv"], 0)
            bqkv = torch.cat([get(p + f"self_attn.{n}_proj.bias") for n in "qkv"], 0)
            if fold_ln:
                wqkv, c1, bqkv = folded(wqkv, bqkv, get(p + "layer_norm1.weight"), get(p + "layer_norm1.bias"))
                ly.c1_qkv = vec(c1).data_ptr()
            ly.w_qkv = mat(wqkv).data_ptr()
            ly.b_qkv = vec(bqkv).data_ptr()
            ly.w_o = mat(get(p + "self_attn.out_proj.weight")).data_ptr()
            ly.b_o = vec(get(p + "self_attn.out_proj.bias")).data_ptr()
            ly.ln2_g = vec(get(p + "layer_norm2.weight")).data_ptr()
            ly.ln2_b = vec(get(p + "layer_norm2.bias")).data_ptr()
            wfc1, bfc1 = get(p + "mlp.fc1.weight"), get(p + "mlp.fc1.bias")
            if fold_ln:
                wfc1, c1, bfc1 = folded(wfc1, bfc1, get(p + "layer_norm2.weight"), get(p + "layer_norm2.bias"))
                ly.c1_fc1 = vec(c1).data_ptr()
            ly.w_fc1 = mat(wfc1).data_ptr()
            ly.b_fc1 = vec(bfc1).data_ptr()
            ly.w_fc2 = mat(get(p + "mlp.fc2.weight")).data_ptr()
            ly.b_fc2 = vec(get(p + "mlp.fc2.bias")).data_ptr()
        w.layers = ctypes.cast(self._layers, ctypes.POINTER(_lib.VitLayer))
        w.post_g = vec(get("post_layernorm.weight")).data_ptr()
        w.post_b = vec(get("post_layernorm.bias")).data_ptr()

        # MAP head: q = probe . Wq^T + bq is input independent -> precompute (fp32), fold in hd^-0.5
        ipw, ipb = get("head.attention.in_proj_weight"), get("head.attention.in_proj_bias")
        probe = get("head.probe").reshape(D)
        q = (ipw[:D] @ probe + ipb[:D]) * (spec.head_dim ** -0.5)
        w.probe_q = vec(q).data_ptr()
        wkv, bkv = ipw[D:], ipb[D:]
        w.fold_ln = 1 if fold_ln else 0
        if fold_ln:
            wkv, c1, bkv = folded(wkv, bkv, get("post_layernorm.weight"), get("post_layernorm.bias"))
            w.c1_kv = vec(c1).data_ptr()
        w.w_kv = mat(wkv).data_ptr()
        w.b_kv = vec(bkv).data_ptr()
        w.w_ho = mat(get("head.attention.out_proj.weight")).data_ptr()
        w.b_ho = vec(get("head.attention.out_proj.bias")).data_ptr()
        w.hln_g = vec(get("head.layernorm.weight")).data_ptr()
        w.hln_b = vec(get("head.layernorm.bias")).data_ptr()
        w.w_hfc1 = mat(get("head.mlp.fc1.weight")).data_ptr()
        w.b_hfc1 = vec(get("head.mlp.fc1.bias")).data_ptr()
        w.w_hfc2 = mat(get("head.mlp.fc2.weight")).data_ptr()
        w.b_hfc2 = vec(get("head.mlp.fc2.bias")).data_ptr()
        self.struct = w

    @property
    def tokens(self) -> int:
        return self.spec.tokens

    def workspace_bytes(self, batch: int) -> int:
        return int(_lib.lib().gvl_siglip_workspace_bytes(ctypes.byref(self.struct), int(batch)))

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._keep)


class SiglipGridView:
    """A `SiglipPack` seen with a gh x gw patch grid (the masked-region variant, include/gvl.h K9): the same device
    weights, `T = gh * gw` tokens per item and the position table re-sampled to that grid
    (`gvl_pos_interp_bicubic_bf16`).  Holds the base pack alive; `struct` is a copy of the base struct."""

    def __init__(self, base: SiglipPack, gh: int, gw: int, pos: torch.Tensor):
        if pos.dtype != torch.bfloat16 or tuple(pos.shape) != (gh * gw, base.spec.hidden) or not pos.is_contiguous():
            raise RuntimeError("SiglipGridView: pos must be contiguous bf16 [gh*gw, D]")
        self.base, self.spec, self.device = base, base.spec, base.device
        self.grid, self.tokens, self.pos_table = (gh, gw), gh * gw, pos
        w = _lib.VitWeights.from_buffer_copy(base.struct)
        w.T = gh * gw
        w.pos = pos.data_ptr()
        self.struct = w

    def workspace_bytes(self, batch: int) -> int:
        return int(_lib.lib().gvl_siglip_workspace_bytes(ctypes.byref(self.struct), int(batch)))


class ProjectorPack:
    """Device-resident bf16 projector (Linear - GELU(erf) - Linear) weights."""

    def __init__(self, sd: dict[str, torch.Tensor], device: torch.device | str):
        dev = self.device = resolve_device(device)
        self.w1 = sd["net.0.weight"].detach().to(torch.bfloat16).contiguous().to(dev)
        self.b1 = sd["net.0.bias"].detach().to(torch.float32).contiguous().to(dev)
        self.w2 = sd["net.2.weight"].detach().to(torch.bfloat16).contiguous().to(dev)
        self.b2 = sd["net.2.bias"].detach().to(torch.float32).contiguous().to(dev)
        self.llm_dim, self.encoder_dim = self.w1.shape


# ---------------------------------------------------------------------------------------------- VideoMAE
@dataclass(frozen=True)
class VideoMAESpec:
    """`VideoMAEModel` geometry (HF: models/videomae/configuration_videomae.py).  The reference loads
    `MCG-NJU/videomae-base` (scripts/extract_features.py:348-350)."""
    hidden: int = 768
    intermediate: int = 3072
    layers: int = 12
    heads: int = 12
    image: int = 224
    patch: int = 16
    frames: int = 16
    tubelet: int = 2
    eps: float = 1e-12
    act: str = "gelu"
    final_norm: bool = True  # checkpoints with use_mean_pooling=False (MCG-NJU/videomae-base) carry `layernorm`

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return (self.frames // self.tubelet) * self.grid * self.grid

    @property
    def patch_k(self) -> int:
        return 3 * self.tubelet * self.patch * self.patch

    patch_ld = patch_k

    @staticmethod
    def base() -> "VideoMAESpec":
        return VideoMAESpec()

    @staticmethod
    def tiny() -> "VideoMAESpec":
        """Small encoder (head dim 64, ragged token count 2*3*3 = 18 per clip) for fast parity tests."""
        return VideoMAESpec(hidden=128, intermediate=272, layers=2, heads=2, image=48, patch=16, frames=4)

    def flops_per_clip(self, with_projector_llm: int | None = 4096) -> int:
        T, D, I = self.tokens, self.hidden, self.intermediate
        total = 2 * T * D * self.patch_k + self.layers * (2 * T * 3 * D * D + 2 * T * D * D + 4 * T * T * D + 4 * T * D * I)
        if with_projector_llm:
            total += 2 * D * with_projector_llm + 2 * with_projector_llm * with_projector_llm
        return total


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """The fixed position table of `VideoMAEEmbeddings` (HF: models/videomae/modeling_videomae.py:80-91), float64
    math like the numpy original, returned as fp32 [n_position, d_hid]."""
    pos = torch.arange(n_position, dtype=torch.float64)[:, None]
    j = torch.arange(d_hid, dtype=torch.float64)[None, :]
    angle = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2.0 * torch.div(j, 2, rounding_mode="floor") / d_hid)
    table = angle.clone()
    table[:, 0::2] = torch.sin(angle[:, 0::2])
    table[:, 1::2] = torch.cos(angle[:, 1::2])
    return table.to(torch.float32)


def synth_videomae_state_dict(spec: VideoMAESpec, seed: int = 2) -> dict[str, torch.Tensor]:
    """Random encoder weights under the HF `VideoMAEModel` names; bf16-representable values (see module doc)."""
    g = torch.Generator().manual_seed(int(seed))
    D, I, P = spec.hidden, spec.intermediate, spec.patch
    sd: dict[str, torch.Tensor] = {}

    def normal(*shape, std):
        return _bf16_round(torch.randn(*shape, generator=g) * std)

    def layernorm(prefix):
        sd[prefix + ".weight"] = _bf16_round(1.0 + 0.1 * torch.randn(D, generator=g))
        sd[prefix + ".bias"] = normal(D, std=0.05)

    sd["embeddings.patch_embeddings.projection.weight"] = normal(D, 3, spec.tubelet, P, P, std=1.0 / math.sqrt(spec.patch_k))
    sd["embeddings.patch_embeddings.projection.bias"] = normal(D, std=0.02)
    for i in range(spec.layers):
        p = f"encoder.layer.{i}."
        layernorm(p + "layernorm_before")
        sd[p + "attention.attention.query.weight"] = normal(D, D, std=1.5 / math.sqrt(D))
        sd[p + "attention.attention.key.weight"] = normal(D, D, std=1.5 / math.sqrt(D))
        sd[p + "attention.attention.value.weight"] = normal(D, D, std=1.0 / math.sqrt(D))
        sd[p + "attention.attention.q_bias"] = normal(D, std=0.02)
        sd[p + "attention.attention.v_bias"] = normal(D, std=0.02)
        sd[p + "attention.output.dense.weight"] = normal(D, D, std=1.0 / math.sqrt(D))
        sd[p + "attention.output.dense.bias"] = normal(D, std=0.02)
        layernorm(p + "layernorm_after")
        sd[p + "intermediate.dense.weight"] = normal(I, D, std=1.0 / math.sqrt(D))
        sd[p + "intermediate.dense.bias"] = normal(I, std=0.02)
        sd[p + "output.dense.weight"] = normal(D, I, std=1.0 / math.sqrt(I))
        sd[p + "output.dense.bias"] = normal(D, std=0.02)
    if spec.final_norm:
        layernorm("layernorm")
    return sd


class VideoMAEPack:
    """Device-resident VideoMAE encoder weights in the `gvl_vit_weights` layout (include/gvl.h, K6b)."""

    def __init__(self, sd: dict[str, torch.Tensor], spec: VideoMAESpec, device: torch.device | str):
        self.spec = spec
        self.device = resolve_device(device)
        self._keep: list[torch.Tensor] = []
        D, I = spec.hidden, spec.intermediate
        pre = "videomae." if any(k.startswith("videomae.") for k in sd) else ""

        def get(name):
            return sd[pre + name].detach().to(torch.float32)

        def mat(t):
            d = t.to(torch.bfloat16).contiguous().to(self.device)
            self._keep.append(d)
            return d

        def vec(t):
            d = t.to(torch.float32).contiguous().to(self.device)
            self._keep.append(d)
            return d

        w = _lib.VitWeights()
        w.D, w.I, w.H, w.hd, w.L, w.T = D, I, spec.heads, spec.head_dim, spec.layers, spec.tokens
        w.patch_k, w.patch_ld = spec.patch_k, spec.patch_ld
        w.eps = spec.eps
        w.act = _ACT[spec.act]
        # Conv3d weight (D, 3, tubelet, P, P) flattens to exactly the im2col column order c, kt, ky, kx
        w.w_patch = mat(get("embeddings.patch_embeddings.projection.weight").reshape(D, spec.patch_k)).data_ptr()
        w.b_patch = vec(get("embeddings.patch_embeddings.projection.bias")).data_ptr()
        w.pos = mat(sinusoid_table(spec.tokens, D)).data_ptr()
        self._layers = (_lib.VitLayer * spec.layers)()
        for i in range(spec.layers):
            p = f"encoder.layer.{i}."
            a = p + "attention.attention."
            ly = self._layers[i]
            ly.ln1_g = vec(get(p + "layernorm_before.weight")).data_ptr()
            ly.ln1_b = vec(get(p + "layernorm_before.bias")).data_ptr()
            ly.w_qkv = mat(torch.cat([get(a + "query.weight"), get(a + "key.weight"), get(a + "value.weight")], 0)).data_ptr()
            has_bias = (pre + a + "q_bias") in sd
            qb = get(a + "q_bias") if has_bias else torch.zeros(D)
            vb = get(a + "v_bias") if has_bias else torch.zeros(D)
            ly.b_qkv = vec(torch.cat([qb, torch.zeros(D), vb], 0)).data_ptr()  # the key projection has no bias
            ly.w_o = mat(get(p + "attention.output.dense.weight")).data_ptr()
            ly.b_o = vec(get(p + "attention.output.dense.bias")).data_ptr()
            ly.ln2_g = vec(get(p + "layernorm_after.weight")).data_ptr()
            ly.ln2_b = vec(get(p + "layernorm_after.bias")).data_ptr()
            ly.w_fc1 = mat(get(p + "intermediate.dense.weight")).data_ptr()
            ly.b_fc1 = vec(get(p + "intermediate.dense.bias")).data_ptr()
            ly.w_fc2 = mat(get(p + "output.dense.weight")).data_ptr()
            ly.b_fc2 = vec(get(p + "output.dense.bias")).data_ptr()
        w.layers = ctypes.cast(self._layers, ctypes.POINTER(_lib.VitLayer))
        if (pre + "layernorm.weight") in sd:
            w.post_g = vec(get("layernorm.weight")).data_ptr()
            w.post_b = vec(get("layernorm.bias")).data_ptr()
        self.struct = w

    def workspace_bytes(self, batch: int) -> int:
        return int(_lib.lib().gvl_videomae_workspace_bytes(ctypes.byref(self.struct), int(batch)))
