"""Drop-in for the reference's VideoMAE clip route, `run_videomae_encoder`
(scripts/extract_features.py:335-403; twin: scripts/realtime_inference.py:339-381).

Same contract: frames are taken 16 at a time (non-overlapping, the tail clip padded by repeating its last
frame; the realtime twin takes windows of 16 with stride 8 and drops the tail), each clip goes through
`VideoMAEImageProcessor` (shortest-edge-224 uint8 resize, center crop 224, (x/255 - mean)/std with the
checkpoint's `preprocessor_config.json` constants — ImageNet mean / std for MCG-NJU/videomae-base; the class
default 0.5 / 0.5 only applies to a config without values) and `VideoMAEModel`, the 1568 token states are averaged, and the result is a dict with one
`{"start_time", "end_time", "embedding" (768,) fp32 on the CPU, "source_frame_count"}` entry per clip.  Here
the frames are a uint8 `(N, H, W, 3)` tensor (host or device), clips are batched, and every step is a kernel of
libgvl_sm100a.so: gvl_preprocess_u8_crop -> gvl_patchify_tubelet_bf16 -> gvl_videomae_forward (the SigLIP
tower's GEMM / LayerNorm / attention kernels, bf16 with fp32 accumulation; the reference runs the model in
fp32) -> optionally the 768 -> 4096 `videomae_proj` projector.  There is no CPU fallback.
"""
from __future__ import annotations

import torch

import json
import os

from . import ops
from .weights import ProjectorPack, VideoMAEPack, VideoMAESpec, resolve_device


def resize_geometry(H: int, W: int, shortest_edge: int = 224, crop: int = 224):
    """Output size of the processor's shortest-edge resize and the origin of its center crop."""
    if H <= W:
        out_h, out_w = shortest_edge, int(shortest_edge * W / H)
    else:
        out_h, out_w = int(shortest_edge * H / W), shortest_edge
    return out_h, out_w, int(round((out_h - crop) / 2.0)), int(round((out_w - crop) / 2.0))


# preprocessor_config.json of MCG-NJU/videomae-base (the checkpoint both reference routes load): ImageNet statistics
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def clip_windows(n_frames: int, clip: int = 16, stride: int | None = None, drop_tail: bool = False) -> list[tuple[int, int]]:
    """[start, end) frame ranges of the clips of an n-frame video.

    stride None / == clip, drop_tail False: `run_videomae_encoder` (scripts/extract_features.py:355-365) —
    non-overlapping clips, the last one short (the caller pads it by repeating its last frame).
    stride 8, drop_tail True: `extract_videomae_embeddings` (scripts/realtime_inference.py:352-355) —
    `range(0, n - clip + 1, clip // 2)`, windows that do not fit are dropped."""
    stride = clip if stride is None else int(stride)
    if drop_tail:
        return [(i, i + clip) for i in range(0, n_frames - clip + 1, stride)]
    return [(i, min(i + clip, n_frames)) for i in range(0, n_frames, stride)]


class VideoMAEClipEncoder:
    """Batched clip encoder over device-resident weights (an HF `VideoMAEModel.state_dict()` or synthetic)."""

    @classmethod
    def from_checkpoint(cls, path: str, device: str | torch.device = "cuda", **kwargs) -> "VideoMAEClipEncoder":
        """A local HF checkpoint directory (`VideoMAEModel.from_pretrained` + `VideoMAEImageProcessor.from_pretrained`,
        scripts/extract_features.py:348-350): weights from model.safetensors / pytorch_model.bin, geometry from
        config.json, mean / std / resample from preprocessor_config.json."""
        from .siglip_semantic_encoder import _load_state_dict
        sd = _load_state_dict(path)
        d = path if os.path.isdir(path) else os.path.dirname(path)
        spec_kw, pre_kw = {}, {}
        cfg_file, pre_file = os.path.join(d, "config.json"), os.path.join(d, "preprocessor_config.json")
        if os.path.isfile(cfg_file):
            with open(cfg_file) as f:
                c = json.load(f)
            names = {"hidden_size": "hidden", "intermediate_size": "intermediate", "num_hidden_layers": "layers",
                     "num_attention_heads": "heads", "image_size": "image", "patch_size": "patch", "num_frames": "frames",
                     "tubelet_size": "tubelet", "layer_norm_eps": "eps", "hidden_act": "act"}
            spec_kw = {ours: c[theirs] for theirs, ours in names.items() if theirs in c}
            if "use_mean_pooling" in c:
                spec_kw["final_norm"] = not c["use_mean_pooling"]
        if os.path.isfile(pre_file):
            with open(pre_file) as f:
                pc = json.load(f)
            pre_kw = {ours: pc[theirs] for theirs, ours in (("image_mean", "image_mean"), ("image_std", "image_std"),
                                                             ("resample", "resample")) if theirs in pc}
        pre_kw.update(kwargs)
        return cls(sd, VideoMAESpec(**spec_kw), device, **pre_kw)

    def __init__(self, state_dict: dict, spec: VideoMAESpec | None = None, device: str | torch.device = "cuda",
                 image_mean=IMAGENET_MEAN, image_std=IMAGENET_STD, resample: int = ops.BILINEAR,
                 clips_per_batch: int = 32):
        self.spec = spec or VideoMAESpec.base()
        self.device = resolve_device(device)  # raises off-CUDA (no CPU fallback)
        self.pack = VideoMAEPack(state_dict, self.spec, self.device)
        self.image_mean, self.image_std, self.resample = tuple(image_mean), tuple(image_std), resample
        self.clips_per_batch = int(clips_per_batch)
        self._ws: torch.Tensor | None = None

    def _workspace(self, clips: int) -> torch.Tensor:
        need = self.pack.workspace_bytes(clips)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def preprocess(self, frames: torch.Tensor, band: tuple[int, int] | None = None) -> torch.Tensor:
        """uint8 [n,H,W,3] (device) -> bf16 pixel_values [n,3,S,S].  band = (x0, full_width): `frames` holds only the
        source columns [x0, x0 + frames.shape[2]) the center crop reads (`source_band`)."""
        s = self.spec
        _, H, W, _ = frames.shape
        full_w = W if band is None else band[1]
        out_h, out_w, y0, x0 = resize_geometry(H, full_w, s.image, s.image)
        return ops.preprocess_crop(frames, out_h, out_w, y0, x0, s.image, s.image, self.resample, self.image_mean,
                                   self.image_std, layout=ops.LAYOUT_BF16_CHW, band=band)

    def source_band(self, H: int, W: int) -> tuple[int, int]:
        """(x0, width) of the source columns the processor's center crop reads (≈57 % of a 16:9 row): the only part of a
        host frame the feed has to move (`ops.copy_band_h2d`)."""
        s = self.spec
        _, out_w, _, x0 = resize_geometry(H, W, s.image, s.image)
        return ops.crop_source_band(W, out_w, x0, s.image, self.resample)

    def encode_clips(self, frames: torch.Tensor, out_dtype: torch.dtype = torch.float32,
                     band: tuple[int, int] | None = None) -> torch.Tensor:
        """uint8 [clips*frames_per_clip, H, W, 3] on the device -> clip embeddings [clips, hidden]."""
        s = self.spec
        if frames.shape[0] % s.frames:
            raise RuntimeError(f"encode_clips: frame count {frames.shape[0]} is not a multiple of {s.frames}")
        clips = frames.shape[0] // s.frames
        pv = self.preprocess(frames, band)
        patches = ops.patchify_tubelet(pv, s.frames, s.patch, s.tubelet)
        return ops.videomae_forward(self.pack, patches, self._workspace(clips), out_dtype=out_dtype)

    def run(self, frames: torch.Tensor, timestamps, projector: ProjectorPack | None = None) -> dict:
        """Mirror of `run_videomae_encoder(frames)`: `frames` uint8 [N,H,W,3], `timestamps` one float per frame."""
        s = self.spec
        n = int(frames.shape[0])
        embeddings = []
        per_batch = self.clips_per_batch * s.frames
        # host frames: only the column band the center crop reads crosses PCIe
        band = None
        if not frames.is_cuda and frames.is_contiguous() and frames.shape[2] > frames.shape[1]:
            bx0, bw = self.source_band(int(frames.shape[1]), int(frames.shape[2]))
            if bw < frames.shape[2]:
                band = (bx0, int(frames.shape[2]))
        for start in range(0, n, per_batch):
            host = frames[start:min(n, start + per_batch)]
            if band is not None:
                chunk = torch.empty((host.shape[0], host.shape[1], bw, 3), dtype=torch.uint8, device=self.device)
                ops.copy_band_h2d(chunk, host, bx0)
            else:
                chunk = host.to(self.device, non_blocking=True)
            real = chunk.shape[0]
            pad = (-real) % s.frames
            if pad:  # reference: `while len(clip_frames) < clip_size: clip_frames.append(clip_frames[-1])`
                chunk = torch.cat([chunk, chunk[-1:].expand(pad, -1, -1, -1)], 0).contiguous()
            emb = self.encode_clips(chunk, band=band)
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

    def run_realtime(self, frames: torch.Tensor, timestamps) -> list[dict]:
        """Mirror of `extract_videomae_embeddings(frames)` (scripts/realtime_inference.py:339-381): windows of 16
        frames with stride 8, windows that do not fit are dropped, one `{"timestamp": ts of the window's middle frame,
        "embedding": (768,) fp32 on the CPU}` per window.  Every frame is resized / normalized once although it
        belongs to two windows; the windows are then gathered on the device."""
        s = self.spec
        n = int(frames.shape[0])
        windows = clip_windows(n, s.frames, s.frames // 2, drop_tail=True)
        out: list[dict] = []
        for w0 in range(0, len(windows), self.clips_per_batch):
            batch = windows[w0:w0 + self.clips_per_batch]
            f0, f1 = batch[0][0], batch[-1][1]
            pv = self.preprocess(frames[f0:f1].to(self.device, non_blocking=True).contiguous())
            gather = torch.tensor([i - f0 for a, b in batch for i in range(a, b)], dtype=torch.long, device=self.device)
            patches = ops.patchify_tubelet(pv.index_select(0, gather), s.frames, s.patch, s.tubelet)
            emb = ops.videomae_forward(self.pack, patches, self._workspace(len(batch)), out_dtype=torch.float32).cpu()
            for c, (a, _) in enumerate(batch):
                out.append({"timestamp": float(timestamps[a + s.frames // 2]), "embedding": emb[c]})
        return out
