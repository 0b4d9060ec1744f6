"""The hot path as one object: uint8 frames -> SigLIP2 embedding (1152) -> projected embedding (4096) ->
timeline index, batched, stream-ordered, one process per GPU.

Replaces the reference's per-frame loop (scripts/extract_features.py:590-607 `run_siglip_encoder`, twin
scripts/realtime_inference.py:313-324): there every frame costs a CPU preprocess, one H2D, a batch-1
forward and a D2H sync; here a chunk of the timeline is resident (or streamed through a double-buffered
pinned-memory feed), each batch is three C-ABI calls (preprocess, tower, projector) on one stream, and
the only host synchronisation is at the end of the chunk.
"""
from __future__ import annotations

from struct import error as struct_error
from typing import Iterable, Iterator

import numpy as np
import torch

from . import ops
from .weights import ProjectorPack, SiglipPack, SiglipVisionSpec, resolve_device


def shard_range(n_items: int, rank: int, world: int, align: int = 1) -> tuple[int, int]:
    """Contiguous timeline chunk of `rank`: [r*ceil(n/world), min(n, (r+1)*ceil(n/world))), chunk size rounded
    up to a multiple of `align` (16 for VideoMAE clips) so timestamps stay ordered after the gather."""
    per = -(-n_items // world)
    per = -(-per // align) * align
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class _NvdecUnusable(Exception):
    """Hardware decode cannot be used for this file on this machine (decoder="auto" then takes the OpenCV feed)."""


class EmbeddingPipeline:
    def __init__(self, siglip_sd: dict, projector_sd: dict, spec: SiglipVisionSpec | None = None,
                 device: str | torch.device = "cuda", batch: int = 64, resample: int = ops.BILINEAR,
                 image_mean=(0.5, 0.5, 0.5), image_std=(0.5, 0.5, 0.5), fold_ln: bool = True):
        """fold_ln: run the tower with every token-level LayerNorm folded into the GEMM that consumes it
        (weights.SiglipPack); the 56 LayerNorm launches per batch disappear.  False keeps the reference's op order."""
        self.spec = spec or SiglipVisionSpec.so400m()
        self.device = resolve_device(device)  # raises off-CUDA: this path has no CPU fallback
        self.batch = int(batch)
        self.resample = resample
        self.image_mean, self.image_std = tuple(image_mean), tuple(image_std)
        with torch.cuda.device(self.device):
            self.pack = SiglipPack(siglip_sd, self.spec, self.device, fold_ln=fold_ln)
            self.proj = ProjectorPack(projector_sd, self.device)
            s = self.spec
            self.workspace = torch.empty(self.pack.workspace_bytes(self.batch), dtype=torch.uint8, device=self.device)
            self.patches = torch.empty((self.batch * s.tokens, s.patch_ld), dtype=torch.bfloat16, device=self.device)
            self.hidden = torch.empty((self.batch, self.proj.llm_dim), dtype=torch.bfloat16, device=self.device)
        self.llm_dim = self.proj.llm_dim

    # ---- one batch, everything already on the device ------------------------------------------------
    def embed(self, frames: torch.Tensor, out_index: torch.Tensor | None = None):
        """frames uint8 [B,H,W,3] on the device (B <= batch) -> (pooled bf16 [B,D], projected bf16 [B,llm]).
        `out_index` (bf16 [B, llm], e.g. a slice of the timeline index) receives the projection in place."""
        B = frames.shape[0]
        if B > self.batch:
            raise RuntimeError(f"batch of {B} frames exceeds the pipeline's capacity {self.batch}")
        s = self.spec
        patches = self.patches[: B * s.tokens]
        ops.preprocess(frames, s.image, s.image, self.resample, self.image_mean, self.image_std,
                       layout=ops.LAYOUT_BF16_PATCH, patch=s.patch, ld=s.patch_ld, out=patches)
        pooled = ops.siglip_forward(self.pack, patches, workspace=self.workspace)
        projected = ops.project(self.proj, pooled, out_dtype=torch.bfloat16, hidden=self.hidden[:B], out=out_index)
        return pooled, projected

    # ---- a resident chunk of the timeline ---------------------------------------------------------
    def embed_resident(self, frames: torch.Tensor, index: torch.Tensor | None = None) -> torch.Tensor:
        """frames uint8 [N,H,W,3] resident in HBM -> projected index bf16 [N, llm] (timestamp order)."""
        N = frames.shape[0]
        if index is None:
            index = torch.empty((N, self.llm_dim), dtype=torch.bfloat16, device=self.device)
        for i0 in range(0, N, self.batch):
            i1 = min(N, i0 + self.batch)
            self.embed(frames[i0:i1], out_index=index[i0:i1])
        return index

    # ---- host frames streamed through pinned memory -----------------------------------------------
    def embed_stream(self, host_batches: Iterable[torch.Tensor], index: torch.Tensor,
                     host_out: torch.Tensor | None = None, pooled_out: torch.Tensor | None = None,
                     host_pooled: torch.Tensor | None = None) -> int:
        """The call a user makes with decoded frames in host memory.

        host_batches yields pinned uint8 [b,H,W,3] tensors (b <= batch) in timeline order; rows of `index`
        (bf16 [>=N, llm], device) are filled in order.  `pooled_out` (bf16 [>=N, hidden], device) also keeps the
        1152-d SigLIP vectors — what the reference caches (scripts/extract_features.py:597-603, 1447-1468) — so ONE
        pass yields both the timeline index and the cache files.  `host_out` / `host_pooled` (pinned bf16 [>=N, llm] /
        [>=N, hidden]) receive asynchronous device->host copies of each batch's rows.  H2D copies run on one side
        stream and overlap the previous batch's compute (two device frame buffers); D2H copies run on a second side
        stream behind an event, so neither direction ever sits on the compute stream.  Returns the number of frames;
        the compute stream is made to wait for the last D2H, so the caller only synchronises its current stream.
        """
        compute = torch.cuda.current_stream(self.device)
        recycle = getattr(host_batches, "recycle", None)  # e.g. frame_ingest.FrameFeed: pinned ring buffers
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(self.device)
            self._d2h_stream = torch.cuda.Stream(self.device)
            self._dev_frames = [None, None]
            self._copied = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
        if host_pooled is not None and pooled_out is None:
            pooled_out = torch.empty((index.shape[0], self.spec.hidden), dtype=torch.bfloat16, device=self.device)
        done = 0
        d2h_used = False
        for step, hb in enumerate(host_batches):
            slot = step & 1
            b = hb.shape[0]
            buf = self._dev_frames[slot]
            if buf is None or buf.shape[1:] != hb.shape[1:] or buf.shape[0] < b:
                buf = torch.empty((self.batch,) + tuple(hb.shape[1:]), dtype=torch.uint8, device=self.device)
                self._dev_frames[slot] = buf
            with torch.cuda.stream(self._copy_stream):
                if step >= 2:
                    self._copy_stream.wait_event(self._consumed[slot])
                else:
                    self._copy_stream.wait_stream(compute)
                buf[:b].copy_(hb, non_blocking=True)
                self._copied[slot].record(self._copy_stream)
                if recycle is not None:  # the source may reuse the host buffer once this copy has run
                    done_ev = torch.cuda.Event()
                    done_ev.record(self._copy_stream)
                    recycle(hb, done_ev)
            compute.wait_event(self._copied[slot])
            pooled, _ = self.embed(buf[:b], out_index=index[done:done + b])
            if pooled_out is not None:  # `pooled` is a fresh allocation of the compute stream: park it before the D2H
                pooled_out[done:done + b].copy_(pooled, non_blocking=True)
            self._consumed[slot].record(compute)
            if host_out is not None or host_pooled is not None:
                d2h_used = True
                with torch.cuda.stream(self._d2h_stream):
                    self._d2h_stream.wait_event(self._consumed[slot])
                    if host_out is not None:
                        host_out[done:done + b].copy_(index[done:done + b], non_blocking=True)
                    if host_pooled is not None:
                        host_pooled[done:done + b].copy_(pooled_out[done:done + b], non_blocking=True)
            done += b
        if d2h_used:
            compute.wait_stream(self._d2h_stream)
        return done

    # ---- a video file: decode -> pinned ring -> device, overlapped -------------------------------------
    def embed_video(self, video_path: str, fps: float = 1.0, index: torch.Tensor | None = None,
                    host_out: torch.Tensor | None = None, return_pooled: bool = False, decoder: str = "auto"):
        """Replaces `extract_frames` + `run_siglip_encoder` (scripts/extract_features.py:230-264, 590-607) for one
        video: frames are sampled with the reference's rule (every int(video_fps / fps)-th frame, timestamp =
        idx / video_fps) and embedded batch by batch.  Returns (timestamps float64 [n], projected index bf16 [n, llm] on
        the device) and, with `return_pooled`, the 1152-d SigLIP rows bf16 [n, hidden] the reference's cache files hold.

        decoder = "nvdec": the GPU's hardware decoder (`nvdec_ingest.NvdecFeed`: the host only demuxes; sampled frames
        are converted to RGB in device memory and go straight into the preprocess kernel — no host frame, no H2D);
        "opencv": software decode on a background thread into pinned batches (`frame_ingest.FrameFeed`), H2D overlapped
        with compute; "auto": nvdec when the driver library, the container (MP4 / MOV) and the codec (H.264 / HEVC 8-bit
        4:2:0) allow it AND the engine has passed `nvdec_ingest.self_test` (a known-answer clip decoded in a child process,
        once per process), else opencv."""
        if decoder not in ("auto", "nvdec", "opencv"):
            raise ValueError("decoder must be 'auto', 'nvdec' or 'opencv'")
        if decoder != "opencv":
            try:
                return self._embed_video_nvdec(video_path, fps, index, host_out, return_pooled, trust=decoder == "nvdec")
            except _NvdecUnusable as exc:
                if decoder == "nvdec":
                    raise RuntimeError(f"hardware decode of {video_path} is not possible: {exc}") from exc
        from .frame_ingest import FrameFeed
        feed = FrameFeed(video_path, fps=fps, batch=self.batch, auto_release=False)
        n_plan = len(feed.timestamps)
        if index is None:
            index = torch.empty((n_plan, self.llm_dim), dtype=torch.bfloat16, device=self.device)
        pooled = torch.empty((n_plan, self.spec.hidden), dtype=torch.bfloat16, device=self.device) if return_pooled else None
        with torch.cuda.device(self.device):
            n = self.embed_stream(feed, index, host_out, pooled_out=pooled)
            torch.cuda.current_stream(self.device).synchronize()
        if return_pooled:
            return feed.timestamps[:n], index[:n], pooled[:n]
        return feed.timestamps[:n], index[:n]

    def _embed_video_nvdec(self, video_path, fps, index, host_out, return_pooled, trust: bool = False):
        from . import nvdec_ingest as nv
        if not nv.available():
            raise _NvdecUnusable("libnvcuvid.so.1 (GPU driver) did not load")
        try:
            track = nv.read_mp4_video_track(video_path)
        except (ValueError, KeyError, struct_error) as exc:  # not an MP4 / MOV container, or not H.264 / HEVC
            raise _NvdecUnusable(str(exc)) from exc
        ok, why = nv.usable(track.codec, self.device)
        if not ok:
            raise _NvdecUnusable(why)
        if not trust:  # decoder="auto": the engine must first reproduce a known-answer clip (in a child process)
            ok, why = nv.self_test(self.device)
            if not ok:
                raise _NvdecUnusable(why)
        feed = nv.NvdecFeed(video_path, fps=fps, batch=self.batch, device=self.device)
        with torch.cuda.device(self.device):
            n_plan = len(feed.timestamps)
            if index is None:
                index = torch.empty((n_plan, self.llm_dim), dtype=torch.bfloat16, device=self.device)
            pooled = torch.empty((n_plan, self.spec.hidden), dtype=torch.bfloat16, device=self.device) if return_pooled else None
            n = 0
            for frames in feed:
                b = frames.shape[0]
                if n + b > index.shape[0]:
                    raise RuntimeError(f"the file holds more sampled frames than its sample table announced ({n + b} > {n_plan})")
                pl, _ = self.embed(frames, out_index=index[n:n + b])
                if pooled is not None:
                    pooled[n:n + b].copy_(pl)
                n += b
            if host_out is not None:
                host_out[:n].copy_(index[:n], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            feed.close()
        if return_pooled:
            return feed.timestamps[:n], index[:n], pooled[:n]
        return feed.timestamps[:n], index[:n]

    def write_caches(self, video_path: str, out_dir: str, fps: float = 1.0, perception_cache_dir: str | None = None):
        """One pass over a video -> the reference's cache files: `{stem}_embeddings.pt` (C1,
        scripts/extract_features.py:1431-1468) and, if asked, the `PerceptionCache` directory (C3,
        benchmarks/perception_cache.py:203-283).  Returns (timestamps, index, pooled)."""
        import os

        from . import feature_cache as fc
        ts, index, pooled = self.embed_video(video_path, fps=fps, return_pooled=True)
        stem = os.path.splitext(os.path.basename(video_path))[0]
        os.makedirs(out_dir, exist_ok=True)
        fc.write_embeddings_pt(os.path.join(out_dir, f"{stem}_embeddings.pt"), fc.siglip_embedding_records(ts, pooled))
        if perception_cache_dir is not None:
            fc.write_perception_npz(perception_cache_dir, video_path, siglip=pooled, timestamps=np.asarray(ts),
                                    video_duration_sec=float(ts[-1]) if len(ts) else 0.0, perception_config={"fps": fps})
        return ts, index, pooled


def run_siglip_encoder(frames, device: str = "cuda", sam_results: list | None = None, entity_tracker=None,
                       encoder=None) -> list[dict]:
    """Drop-in for `run_siglip_encoder` (scripts/extract_features.py:502-610): frames = [(timestamp, PIL image or uint8
    array)], optional SAM detections [{"timestamp", "bbox", "entity_type", "entity_id", "description"}] -> the
    reference's list of dicts (`timestamp`, `embedding` (CPU tensor), `embedding_shape`, `entity_type`, `description`,
    and for detections `entity_id`, `bbox`) in the reference's order.

    With detections (and an entity tracker, :520) every detection's bounding box is encoded through the masked-region
    route exactly as the reference's one-call-per-detection loop would (:552-583), but all detections of a frame that
    share a target size go through one tower pass, and all frames WITHOUT detections (and every frame in the fallback
    branch, :590-607) are embedded together in batches of `config.batch_size` instead of one forward per frame.
    A failing frame / detection is logged and skipped like the reference (:547, :584, :606).  `encoder`: an already
    constructed `SigLIPSemanticEncoder` (the reference builds a fresh one per call, :514-515)."""
    import logging

    from .siglip_semantic_encoder import BoxMask, NaFlexConfig, SigLIPSemanticEncoder, _image_shape
    log = logging.getLogger(__name__)
    if encoder is None:
        encoder = SigLIPSemanticEncoder(NaFlexConfig(device=device))
    slots: list[list[dict]] = [[] for _ in frames]  # per frame, in detection order
    full_frame: list[int] = []
    note = "Full frame encoding"
    if sam_results and entity_tracker:
        note = "Full frame encoding (no SAM detection)"
        by_time: dict = {}
        for det in sam_results:
            by_time.setdefault(det.get("timestamp", 0), []).append(det)
        for idx, (timestamp, frame) in enumerate(frames):
            dets = [d for d in by_time.get(timestamp, []) if d.get("bbox") is not None]
            if not by_time.get(timestamp):
                full_frame.append(idx)
                continue
            frame_hw = _image_shape(frame)[:2]
            masks, kept = [], []
            for det in dets:
                entity_type = det.get("entity_type", "unknown")
                entity_id = det.get("entity_id", f"{entity_type}_{timestamp}")
                x1, y1, x2, y2 = [int(c) for c in det["bbox"]]
                masks.append((entity_id, BoxMask(frame_hw, y1, y2, x1, x2)))  # the reference's zeros + slice fill
                kept.append(det)
            try:
                regions = encoder.encode_regions_individually(frame, masks)
            except Exception as e:  # one bad detection must not lose the frame: fall back to one call per detection
                log.warning("SigLIP mask encoding failed at %.1fs (%s); retrying per detection", timestamp, e)
                regions = []
                for mk in masks:
                    try:
                        regions.append(encoder.encode_masked_regions(frame, [mk])[0])
                    except Exception as e2:
                        log.warning("SigLIP mask encoding failed for %s at %.1fs: %s", mk[0], timestamp, e2)
                        regions.append(None)
            for det, emb in zip(kept, regions):
                if emb is None:
                    continue
                entity_type = det.get("entity_type", "unknown")
                slots[idx].append({
                    "timestamp": timestamp, "embedding": emb.embedding.cpu(), "embedding_shape": list(emb.embedding.shape),
                    "entity_type": entity_type, "entity_id": emb.entity_id,
                    "description": det.get("description", f"Detected {entity_type}"), "bbox": det["bbox"]})
    else:
        full_frame = list(range(len(frames)))
    # frames encoded whole: runs of equal size go through the encoder's batched, pinned, threaded feed
    i = 0
    while i < len(full_frame):
        shape = _image_shape(frames[full_frame[i]][1])
        run = [full_frame[i]]
        while i + len(run) < len(full_frame) and _image_shape(frames[full_frame[i + len(run)]][1]) == shape:
            run.append(full_frame[i + len(run)])
        i += len(run)
        try:
            rows = dict(zip(run, encoder.encode_images([frames[j][1] for j in run]).cpu()))
        except Exception as e:  # one bad frame must not lose its run: the reference skips only the failing frame (:606)
            log.warning("SigLIP batch starting at %.1fs failed (%s); retrying frame by frame", frames[run[0]][0], e)
            rows = {}
            for j in run:
                try:
                    rows[j] = encoder.encode_images([frames[j][1]])[0].cpu()
                except Exception as e2:
                    log.warning("SigLIP failed at %.1fs: %s", frames[j][0], e2)
        for j, emb in rows.items():
            slots[j].append({"timestamp": frames[j][0], "embedding": emb, "embedding_shape": list(emb.shape),
                             "entity_type": "full_frame", "description": note})
    return [d for s in slots for d in s]


def run_videomae_encoder(frames, device: str = "cuda", encoder=None) -> dict:
    """Drop-in for `run_videomae_encoder` (scripts/extract_features.py:335-403): frames = [(timestamp, PIL image)] ->
    {"num_input_frames", "num_embeddings", "embeddings": [{"start_time", "end_time", "embedding" (768,) fp32 CPU,
    "source_frame_count"}], "embedding_dim"} — non-overlapping clips of 16 frames, the tail padded with its last frame.

    `encoder`: a `VideoMAEClipEncoder` (the reference loads MCG-NJU/videomae-base from the hub on every call; there is no
    hub here, so the caller passes one built from a local checkpoint or a state dict).  PIL frames are read through
    Pillow's zero-copy Arrow export into a pinned buffer one batch of clips at a time; a failure returns the reference's
    empty result (:399-403)."""
    import logging

    from .siglip_semantic_encoder import _image_shape, _rgbx_view, _to_uint8_hwc
    log = logging.getLogger(__name__)
    empty = {"num_input_frames": len(frames), "num_embeddings": 0, "embeddings": [], "embedding_dim": 768}
    try:
        if encoder is None:
            raise RuntimeError("run_videomae_encoder needs a VideoMAEClipEncoder (no hub access to load "
                               "MCG-NJU/videomae-base): VideoMAEClipEncoder.from_checkpoint(path, device)")
        if not frames:
            return dict(empty, embedding_dim=encoder.spec.hidden)
        spec = encoder.spec
        per_batch = encoder.clips_per_batch * spec.frames
        shape = _image_shape(frames[0][1])
        timestamps = [float(t) for t, _ in frames]
        embeddings: list[dict] = []
        stage = None
        for start in range(0, len(frames), per_batch):
            chunk = [f for _, f in frames[start:start + per_batch]]
            if any(_image_shape(f) != shape for f in chunk):
                raise ValueError("frames of one video must have one size")
            views = [_rgbx_view(f) for f in chunk]
            if all(v is not None for v in views):
                if stage is None or stage.shape[0] < len(chunk):
                    stage = torch.empty((per_batch,) + shape[:2] + (4,), dtype=torch.uint8).pin_memory()
                host = stage.numpy()
                for i, v in enumerate(views):
                    np.copyto(host[i], v)
                dev = stage[:len(chunk)].to(encoder.device, non_blocking=True)[..., :3].contiguous()
                torch.cuda.current_stream(encoder.device).synchronize()  # the pinned buffer is refilled next iteration
            else:
                dev = torch.from_numpy(np.stack([_to_uint8_hwc(f) for f in chunk])).to(encoder.device)
            part = encoder.run(dev, timestamps[start:start + len(chunk)])
            embeddings += part["embeddings"]
        log.info("VideoMAE: %d frames -> %d clip embeddings (%d-dim)", len(frames), len(embeddings), spec.hidden)
        return {"num_input_frames": len(frames), "num_embeddings": len(embeddings), "embeddings": embeddings,
                "embedding_dim": spec.hidden}
    except Exception as e:
        log.warning("VideoMAE failed: %s", e)
        return empty


def extract_siglip_embeddings(frames, device: str = "cuda", sam_results: list | None = None, encoder=None) -> list[dict]:
    """Drop-in for `extract_siglip_embeddings` of the online pipeline (scripts/realtime_inference.py:244-335): frames =
    [(timestamp, PIL image)], optional `sam_results` = [{"timestamp", "frame", "detections": [{"mask" | "bbox", "label"}]}]
    -> [{"timestamp", "label", "embedding" (CPU tensor)}] in the reference's order.

    The reference sends every detection through `encode_image` as its own image — the frame with everything outside
    the mask zeroed (:281-286), or `frame.crop(bbox)` (:287-290, PIL semantics: rounded box, black beyond the frame),
    resized to the model's square input by the processor like any frame.  Here the images are built the same way on the
    host and then embedded together: equal shapes share one preprocess launch (the general two-pass kernel up-scales
    small crops) and one tower pass per `config.batch_size` images instead of one forward per detection.  An image that
    fails is logged and skipped like the reference (:303, :324)."""
    import logging

    from .siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder, _image_shape, _to_uint8_hwc
    log = logging.getLogger(__name__)
    if encoder is None:
        encoder = SigLIPSemanticEncoder(NaFlexConfig(device=device))
    items: list[tuple] = []  # (timestamp, label, PIL image or uint8 array)
    if sam_results and any(d["detections"] for d in sam_results):
        for sam_frame in sam_results:
            timestamp, frame = sam_frame["timestamp"], sam_frame["frame"]
            if not sam_frame["detections"]:
                items.append((timestamp, "full_frame", frame))  # PIL frames stay PIL: encode_images reads them in place
                continue
            frame_np = None
            for det in sam_frame["detections"]:
                try:
                    if det.get("mask") and hasattr(det["mask"], "mask"):
                        frame_np = np.array(frame) if frame_np is None else frame_np
                        masked = frame_np.copy()
                        masked[~det["mask"].mask] = 0
                        img = masked
                    elif det.get("bbox"):
                        x1, y1, x2, y2 = det["bbox"]
                        img = _to_uint8_hwc(frame.crop((x1, y1, x2, y2)))
                    else:
                        img = frame
                    items.append((timestamp, det.get("label", "region"), img))
                except Exception as e:
                    log.debug("Region encoding failed: %s", e)
    else:
        for timestamp, frame in frames:
            items.append((timestamp, "full_frame", frame))
    # embed: images of one shape go through the batched entry; the output keeps the reference's order
    out: list[dict | None] = [None] * len(items)
    by_shape: dict[tuple, list[int]] = {}
    for i, (_, _, img) in enumerate(items):
        by_shape.setdefault(_image_shape(img), []).append(i)
    for idx in by_shape.values():
        try:
            emb = encoder.encode_images([items[i][2] for i in idx]).cpu()
            rows = {i: emb[j] for j, i in enumerate(idx)}
        except Exception as e:  # one bad image must not lose its whole group
            log.warning("SigLIP batch of %d images failed (%s); retrying one by one", len(idx), e)
            rows = {}
            for i in idx:
                try:
                    rows[i] = encoder.encode_images([items[i][2]])[0].cpu()
                except Exception as e2:
                    log.warning("SigLIP failed at %.1fs: %s", items[i][0], e2)
        for i, r in rows.items():
            out[i] = {"timestamp": items[i][0], "label": items[i][1], "embedding": r}
    return [d for d in out if d is not None]


def pinned_batches(frames: np.ndarray | torch.Tensor, batch: int) -> Iterator[torch.Tensor]:
    """Convenience: slice a host frame array into pinned batches."""
    t = torch.as_tensor(frames)
    for i0 in range(0, t.shape[0], batch):
        chunk = t[i0:i0 + batch]
        yield chunk if chunk.is_pinned() else chunk.contiguous().pin_memory()
