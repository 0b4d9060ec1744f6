"""Drop-in for the projector part of the reference's `src/agent_core/qwen_reasoning_core.py`:
`MultiModalProjector` (:995-1027), `ProjectorBank` (:1030-1122) and the `ReasoningCoreConfig` fields they
read (:671-681).  Same attribute names (`net`, `siglip_proj`, ...), same state-dict layout
(`{"siglip": {"net.0.weight", ...}, "videomae": ..., "audio": ..., "video": ...}`), same `.to(device)` /
`project_*` / `load_weights` / `save_weights` behaviour — but `forward` on a CUDA device runs the two
tcgen05 GEMMs of libgvl_sm100a.so (bias + exact-erf GELU fused in the first epilogue) instead of
`torch.nn.Linear`.  Construction works without a GPU (the reference's tests build
`ProjectorBank(config).to(device)`); projecting on a CPU device raises — there is no CPU fallback.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass

import torch

from . import ops
from .weights import ProjectorPack

logger = logging.getLogger(__name__)


@dataclass
class ReasoningCoreConfig:
    """The subset of the reference's ReasoningCoreConfig this path honours (defaults as in :638-748)."""

    siglip_dim: int = 1152
    videomae_dim: int = 768
    internvideo_dim: int = 1408
    audiomae_dim: int = 1024
    llm_hidden_dim: int = 4096
    device: str = "cuda"
    dtype: torch.dtype = torch.bfloat16
    semantic_top_k: int = 10
    timestamp_window_sec: float = 30.0


class MultiModalProjector(torch.nn.Module):
    """Linear(encoder_dim, llm_dim) -> GELU (exact erf) -> Linear(llm_dim, llm_dim); parameters live in `net`
    exactly like the reference so `state_dict()` / `load_state_dict()` / `.to()` are interchangeable."""

    def __init__(self, encoder_dim: int, llm_dim: int = 4096):
        super().__init__()
        self.net = torch.nn.Sequential(
            torch.nn.Linear(encoder_dim, llm_dim),
            torch.nn.GELU(),
            torch.nn.Linear(llm_dim, llm_dim),
        )
        self.encoder_dim = encoder_dim
        self.llm_dim = llm_dim
        self._pack = None
        self._pack_key = None

    def _device_pack(self) -> ProjectorPack:
        params = [self.net[0].weight, self.net[0].bias, self.net[2].weight, self.net[2].bias]
        key = tuple((p.data_ptr(), p._version, p.device, p.dtype) for p in params)
        if self._pack is None or key != self._pack_key:
            sd = {"net.0.weight": params[0], "net.0.bias": params[1], "net.2.weight": params[2], "net.2.bias": params[3]}
            self._pack = ProjectorPack(sd, params[0].device)
            self._pack_key = key
        return self._pack

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (batch, encoder_dim) or (batch, seq, encoder_dim) -> (..., llm_dim), fp32 like the reference's
        fp32 modules (it calls `.float()` first, :2803); bf16 in -> bf16 GEMM operands, fp32 accumulate."""
        dev = self.net[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("MultiModalProjector.forward needs a CUDA (sm_100a) device; there is no CPU fallback")
        if x.shape[-1] != self.encoder_dim:
            raise RuntimeError(f"expected last dim {self.encoder_dim}, got {tuple(x.shape)}")
        lead = x.shape[:-1]
        x2 = x.to(dev).reshape(-1, self.encoder_dim).to(torch.bfloat16).contiguous()
        with torch.no_grad():
            out = ops.project(self._device_pack(), x2, out_dtype=torch.float32)
        return out.reshape(*lead, self.llm_dim)


class ProjectorBank:
    """Collection of projectors for all modalities (reference :1030-1122)."""

    def __init__(self, config: ReasoningCoreConfig):
        self.config = config
        self.device = config.device
        self.siglip_proj = MultiModalProjector(config.siglip_dim, config.llm_hidden_dim)
        self.videomae_proj = MultiModalProjector(config.videomae_dim, config.llm_hidden_dim)
        self.video_proj = MultiModalProjector(config.internvideo_dim, config.llm_hidden_dim)
        self.audio_proj = MultiModalProjector(config.audiomae_dim, config.llm_hidden_dim)
        self._initialized = False

    def to(self, device: str) -> "ProjectorBank":
        self.siglip_proj = self.siglip_proj.to(device)
        self.videomae_proj = self.videomae_proj.to(device)
        self.video_proj = self.video_proj.to(device)
        self.audio_proj = self.audio_proj.to(device)
        self.device = device
        return self

    def project_region(self, embeddings: torch.Tensor) -> torch.Tensor:
        return self.siglip_proj(embeddings.to(self.device))

    def project_videomae(self, embeddings: torch.Tensor) -> torch.Tensor:
        return self.videomae_proj(embeddings.to(self.device))

    def project_video(self, embeddings: torch.Tensor) -> torch.Tensor:
        return self.video_proj(embeddings.to(self.device))

    def project_audio(self, embeddings: torch.Tensor) -> torch.Tensor:
        return self.audio_proj(embeddings.to(self.device))

    def load_weights(self, path: str) -> None:
        state_dict = torch.load(path, map_location=self.device, weights_only=False)
        if "siglip" in state_dict:
            self.siglip_proj.load_state_dict(state_dict["siglip"])
        if "videomae" in state_dict:
            self.videomae_proj.load_state_dict(state_dict["videomae"])
        if "audio" in state_dict:
            self.audio_proj.load_state_dict(state_dict["audio"])
        if "video" in state_dict:
            self.video_proj.load_state_dict(state_dict["video"])
        self._initialized = True
        logger.info("Loaded projector weights from %s", path)

    def save_weights(self, path: str) -> None:
        torch.save({
            "siglip": self.siglip_proj.state_dict(),
            "videomae": self.videomae_proj.state_dict(),
            "audio": self.audio_proj.state_dict(),
            "video": self.video_proj.state_dict(),
        }, path)
        logger.info("Saved projector weights to %s", path)


def project_embeddings(projectors: ProjectorBank, region_embeddings=None, videomae_embeddings=None,
                       audio_embeddings=None, device=None) -> dict:
    """`PerceptionReasoningLoop.project_embeddings` (reference :2779-2814): `.to(device).float()` then the
    matching projector, under no_grad; returns {"siglip": (N, llm), "videomae": ..., "audio": ...}."""
    device = device or projectors.device
    projected = {}
    with torch.no_grad():
        if region_embeddings is not None:
            projected["siglip"] = projectors.project_region(region_embeddings.to(device).float())
        if videomae_embeddings is not None:
            projected["videomae"] = projectors.project_videomae(videomae_embeddings.to(device).float())
        if audio_embeddings is not None:
            projected["audio"] = projectors.project_audio(audio_embeddings.to(device).float())
    return projected
