"""Frame ingest: the step in front of the hot path.

The reference decodes the WHOLE video into a Python list of PIL images before the first frame is encoded
(scripts/extract_features.py:230-264 `extract_frames`; 22 GB of RAM for one hour of 1080p at 1 fps) and then feeds
them to the encoder one by one.  Here the same frames — the same sampling rule and the same timestamps — are decoded
on a producer thread straight into a small ring of PINNED host batches that `EmbeddingPipeline.embed_stream` copies
to the device while the previous batch is being embedded: host memory stays at `depth` batches, decode overlaps
compute.

Sampling rule (extract_features.py:247-253): `frame_interval = int(video_fps / fps)`,
`indices = range(0, total_frames, frame_interval)`, `timestamp = idx / video_fps`.

The decoder is OpenCV (`cv2.VideoCapture`, BGR -> RGB); the reference's decord is not in this image and NVDEC has no
Python binding here — the feed's interface (pinned uint8 [b,H,W,3] batches in timeline order) is what an NVDEC
source would implement as well.
"""
from __future__ import annotations

import queue
import threading
from typing import Iterator, Optional

import numpy as np
import torch


def sample_plan(total_frames: int, video_fps: float, fps: float = 1.0) -> tuple[np.ndarray, np.ndarray]:
    """(frame indices int64, timestamps float64) the reference samples: every int(video_fps / fps)-th frame from 0,
    timestamp = idx / video_fps.  fps > video_fps makes the interval 0 and raises, as `range(0, n, 0)` does there."""
    frame_interval = int(video_fps / fps)
    if frame_interval <= 0:
        raise ValueError("range() arg 3 must not be zero")  # the reference's failure for fps > video_fps
    idx = np.arange(0, int(total_frames), frame_interval, dtype=np.int64)
    return idx, idx.astype(np.float64) / float(video_fps)


def _open(video_path: str):
    import cv2
    cap = cv2.VideoCapture(str(video_path))
    if not cap.isOpened():
        raise RuntimeError(f"cannot open video {video_path!r}")
    return cap


def video_info(video_path: str) -> tuple[int, float, int, int]:
    """(total_frames, avg fps, height, width)."""
    import cv2
    cap = _open(video_path)
    try:
        return (int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), float(cap.get(cv2.CAP_PROP_FPS)),
                int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)))
    finally:
        cap.release()


def _decode_sampled(video_path: str, indices: np.ndarray) -> Iterator[np.ndarray]:
    """RGB uint8 frames at the (ascending) indices: sequential decode, unwanted frames are grabbed without being
    retrieved (no colour conversion, no copy) — seeking is neither exact nor faster at these sampling ratios."""
    import cv2
    cap = _open(video_path)
    try:
        pos = 0
        for want in indices.tolist():
            while pos < want:
                if not cap.grab():
                    return
                pos += 1
            ok, bgr = cap.read()
            pos += 1
            if not ok:
                return
            yield cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    finally:
        cap.release()


def extract_frames(video_path: str, fps: float = 1.0, as_pil: bool = True) -> list:
    """Drop-in for scripts/extract_features.py:230 `extract_frames`: list of (timestamp_seconds, PIL RGB image)
    (numpy arrays with as_pil=False).  Holds every frame in memory like the reference; prefer `FrameFeed`."""
    total, video_fps, _, _ = video_info(video_path)
    idx, ts = sample_plan(total, video_fps, fps)
    out = []
    for t, frame in zip(ts.tolist(), _decode_sampled(video_path, idx)):
        if as_pil:
            from PIL import Image
            frame = Image.fromarray(frame)
        out.append((t, frame))
    return out


class FrameFeed:
    """Iterator of pinned uint8 [b,H,W,3] batches (b <= batch) of the sampled frames, in timeline order, decoded on a
    background thread into a ring of `depth` pinned buffers.  `timestamps` (float64, one per sampled frame) is
    known up front from the container's frame count; `frames_decoded` says how many frames really arrived.

    Buffer lifetime: a batch belongs to the consumer until it is handed back with `recycle(batch, event)`; the
    producer waits for `event` (e.g. the CUDA event recorded behind the asynchronous host-to-device copy) before it
    overwrites the buffer.  `EmbeddingPipeline.embed_stream` does this for any source that has a `recycle` method.
    With auto_release=True (synchronous consumers) taking the next batch hands the previous one back."""

    def __init__(self, video_path: str, fps: float = 1.0, batch: int = 64, depth: int = 3, pin: Optional[bool] = None,
                 auto_release: bool = True):
        if depth < 2:
            raise ValueError("FrameFeed needs a ring of at least 2 batches")
        self.video_path, self.batch, self.depth = str(video_path), int(batch), int(depth)
        total, self.video_fps, self.height, self.width = video_info(self.video_path)
        self.indices, self.timestamps = sample_plan(total, self.video_fps, fps)
        self.frames_decoded = 0
        self.auto_release = bool(auto_release)
        pin = torch.cuda.is_available() if pin is None else pin
        self._ring = [torch.empty((self.batch, self.height, self.width, 3), dtype=torch.uint8, pin_memory=pin)
                      for _ in range(self.depth)]
        self._slot_of = {t.data_ptr(): i for i, t in enumerate(self._ring)}
        self._free: queue.Queue = queue.Queue()
        self._ready: queue.Queue = queue.Queue()
        for i in range(self.depth):
            self._free.put((i, None))
        self._last: Optional[int] = None
        self._error: Optional[BaseException] = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._produce, name="gvl-frame-feed", daemon=True)
        self._started = False

    def __len__(self) -> int:
        return -(-len(self.indices) // self.batch)

    def recycle(self, batch: torch.Tensor, event=None) -> None:
        """Hand a yielded batch back; the buffer is overwritten only after `event.synchronize()` (if given)."""
        slot = self._slot_of.get(batch.data_ptr())
        if slot is None:
            raise RuntimeError("recycle: not a batch of this feed")
        if slot == self._last:
            self._last = None
        self._free.put((slot, event))

    def _produce(self) -> None:
        try:
            slot, fill = None, 0
            for frame in _decode_sampled(self.video_path, self.indices):
                if slot is None:
                    item = self._free.get()
                    if item is None or self._stop.is_set():
                        return
                    slot, event = item
                    fill = 0
                    if event is not None:
                        event.synchronize()
                if frame.shape != (self.height, self.width, 3):
                    raise RuntimeError(f"frame of shape {frame.shape} in a {self.height}x{self.width} video")
                self._ring[slot][fill].copy_(torch.from_numpy(frame))
                fill += 1
                if fill == self.batch:
                    self._ready.put((slot, fill))
                    slot = None
            if slot is not None and fill:
                self._ready.put((slot, fill))
        except BaseException as e:  # surfaced in the consumer thread
            self._error = e
        finally:
            self._ready.put(None)

    def __iter__(self) -> Iterator[torch.Tensor]:
        if self._started:
            raise RuntimeError("a FrameFeed can be iterated once")
        self._started = True
        self._thread.start()
        try:
            while True:
                if self.auto_release and self._last is not None:
                    self._free.put((self._last, None))
                    self._last = None
                item = self._ready.get()
                if item is None:
                    break
                slot, n = item
                self._last = slot
                self.frames_decoded += n
                yield self._ring[slot][:n]
            if self._error is not None:
                raise self._error
        finally:
            self._stop.set()
            self._free.put(None)
