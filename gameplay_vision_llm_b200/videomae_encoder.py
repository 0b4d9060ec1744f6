"""Drop-in for the reference's VideoMAE clip route, `run_videomae_encoder`
(scripts/extract_features.py:335-403; twin: scripts/realtime_inference.py:339-381).

Same contract: frames are taken 16 at a time (non-overlapping, the tail clip padded by repeating its last
frame), each clip goes through `VideoMAEImageProcessor` (shortest-edge-224 uint8 resize, center crop 224,
(x/255 - 0.5)/0.5) and `VideoMAEModel`, the 1568 token states are averaged, and the result is a dict with one
`{"start_time", "end_time", "embedding" (768,) fp32 on the CPU, "source_frame_count"}` entry per clip.  Here
the frames are a uint8 `(N, H, W, 3)` tensor (host or device), clips are batched, and every step is a kernel of
libgvl_sm100a.so: gvl_preprocess_u8_crop -> gvl_patchify_tubelet_bf16 -> gvl_videomae_forward (the SigLIP
tower's GEMM / LayerNorm / attention kernels, bf16 with fp32 accumulation; the reference runs the model in
fp32) -> optionally the 768 -> 4096 `videomae_proj` projector.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import ops
from .weights import ProjectorPack, VideoMAEPack, VideoMAESpec


def resize_geometry(H: int, W: int, shortest_edge: int = 224, crop: int = 224):
    """Output size of the processor's shortest-edge resize and the origin of its center crop."""
    if H <= W:
        out_h, out_w = shortest_edge, int(shortest_edge * W / H)
    else:
        out_h, out_w = int(shortest_edge * H / W), shortest_edge
    return out_h, out_w, int(round((out_h - crop) / 2.0)), int(round((out_w - crop) / 2.0))


class VideoMAEClipEncoder:
    """Batched clip encoder over device-resident weights (an HF `VideoMAEModel.state_dict()` or synthetic)."""

    def __init__(self, state_dict: dict, spec: VideoMAESpec | None = None, device: str | torch.device = "cuda",
                 image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), resample: int = ops.BILINEAR,
                 clips_per_batch: int = 32):
        self.spec = spec or VideoMAESpec.base()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("VideoMAEClipEncoder runs on a CUDA device only (no CPU fallback)")
        self.pack = VideoMAEPack(state_dict, self.spec, self.device)
        self.image_mean, self.image_std, self.resample = tuple(image_mean), tuple(image_std), resample
        self.clips_per_batch = int(clips_per_batch)
        self._ws: torch.Tensor | None = None

    def _workspace(self, clips: int) -> torch.Tensor:
        need = self.pack.workspace_bytes(clips)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def preprocess(self, frames: torch.Tensor) -> torch.Tensor:
        """uint8 [n,H,W,3] (device) -> bf16 pixel_values [n,3,S,S]."""
        s = self.spec
        _, H, W, _ = frames.shape
        out_h, out_w, y0, x0 = resize_geometry(H, W, s.image, s.image)
        return ops.preprocess_crop(frames, out_h, out_w, y0, x0, s.image, s.image, self.resample, self.image_mean,
                                   self.image_std, layout=ops.LAYOUT_BF16_CHW)

    def encode_clips(self, frames: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """uint8 [clips*frames_per_clip, H, W, 3] on the device -> clip embeddings [clips, hidden]."""
        s = self.spec
        if frames.shape[0] % s.frames:
            raise RuntimeError(f"encode_clips: frame count {frames.shape[0]} is not a multiple of {s.frames}")
        clips = frames.shape[0] // s.frames
        pv = self.preprocess(frames)
        patches = ops.patchify_tubelet(pv, s.frames, s.patch, s.tubelet)
        return ops.videomae_forward(self.pack, patches, self._workspace(clips), out_dtype=out_dtype)

    def run(self, frames: torch.Tensor, timestamps, projector: ProjectorPack | None = None) -> dict:
        """Mirror of `run_videomae_encoder(frames)`: `frames` uint8 [N,H,W,3], `timestamps` one float per frame."""
        s = self.spec
        n = int(frames.shape[0])
        embeddings = []
        per_batch = self.clips_per_batch * s.frames
        for start in range(0, n, per_batch):
            chunk = frames[start:min(n, start + per_batch)].to(self.device, non_blocking=True)
            real = chunk.shape[0]
            pad = (-real) % s.frames
            if pad:  # reference: `while len(clip_frames) < clip_size: clip_frames.append(clip_frames[-1])`
                chunk = torch.cat([chunk, chunk[-1:].expand(pad, -1, -1, -1)], 0).contiguous()
            emb = self.encode_clips(chunk)
            proj = ops.project(projector, emb.to(torch.bfloat16)) if projector is not None else None
            emb_cpu = emb.cpu()
            for c in range(chunk.shape[0] // s.frames):
                f0 = start + c * s.frames
                f1 = min(f0 + s.frames, n)
                item = {"start_time": float(timestamps[f0]), "end_time": float(timestamps[f1 - 1]),
                        "embedding": emb_cpu[c], "source_frame_count": s.frames}
                if proj is not None:
                    item["projected"] = proj[c].cpu()
                embeddings.append(item)
        return {"num_input_frames": n, "num_embeddings": len(embeddings), "embeddings": embeddings,
                "embedding_dim": s.hidden}
