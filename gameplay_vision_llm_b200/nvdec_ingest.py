"""Hardware-decoded frame ingest: container demux on the host, picture decode on the GPU's NVDEC
engines, sampled frames delivered as uint8 RGB batches that never leave device memory.

Replaces `extract_frames` (scripts/extract_features.py:230-264; twin scripts/realtime_inference.py:120-158), where decord
decodes the whole file on the CPU, every sampled frame becomes a PIL image held in host RAM (22 GB for an hour of 1080p
at 1 fps) and each one is later copied to the GPU by the processor.  Same sampling rule and timestamps
(`frame_ingest.sample_plan`): every int(video_fps / fps)-th frame in presentation order, t = idx / video_fps.

Host side (this file): a minimal ISO-BMFF (.mp4 / .mov) reader for one H.264 / HEVC video track (sample tables ->
byte ranges, avcC / hvcC parameter sets, length-prefixed NAL units -> Annex-B), or a raw Annex-B file; chunks of samples
go to `gvl_nvdec_feed` (include/gvl.h K10).  Device side: csrc/nvdec.cu.  No fallback inside: `NvdecFeed` raises when the
driver library or the codec is missing; `EmbeddingPipeline.embed_video(decoder="auto")` then picks the OpenCV feed.
"""
from __future__ import annotations

import ctypes
import mmap
import os
import struct
from dataclasses import dataclass, field
from typing import Iterator, Optional

import numpy as np
import torch

from . import _lib
from .frame_ingest import sample_plan
from .weights import resolve_device

CODEC_H264, CODEC_HEVC = 4, 8
_START = b"\x00\x00\x00\x01"


def available() -> bool:
    """libnvcuvid.so.1 present with every entry point (no device call)."""
    return bool(_lib.lib().gvl_nvdec_available())


def decoder_caps(codec: int = CODEC_H264) -> dict:
    sup, mw, mh, ne = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().gvl_nvdec_caps(codec, ctypes.byref(sup), ctypes.byref(mw), ctypes.byref(mh), ctypes.byref(ne)),
               "gvl_nvdec_caps")
    return {"supported": bool(sup.value), "max_width": mw.value, "max_height": mh.value, "engines": ne.value}


def usable(codec: int = CODEC_H264, device=None) -> tuple[bool, str]:
    """(True, "") when this process can decode `codec` on the GPU's NVDEC, else (False, why).  The library being present
    is not enough: a container whose GPU is exposed for compute only (NVIDIA_DRIVER_CAPABILITIES without `video`, or a
    sandbox that proxies just the compute ioctls) loads libnvcuvid but gets CUDA_ERROR_NO_DEVICE from every call."""
    if not available():
        return False, "libnvcuvid.so.1 (part of the GPU driver) did not load"
    try:
        if device is not None:
            with torch.cuda.device(resolve_device(device)):
                caps = decoder_caps(codec)
        else:
            caps = decoder_caps(codec)
    except RuntimeError as exc:
        return False, str(exc)
    if not caps["supported"]:
        return False, f"this GPU's NVDEC does not decode codec id {codec}"
    return True, ""


_SELF_TEST: dict = {}

_SELF_TEST_SCRIPT = """
import sys, tempfile, os, numpy as np, torch
from gameplay_vision_llm_b200 import nvdec_ingest as nv, synth_video as sv
dev = sys.argv[1]
rng = np.random.default_rng(0)
y = rng.integers(16, 236, (3, 48, 64), dtype=np.uint8)
cb = rng.integers(16, 241, (3, 24, 32), dtype=np.uint8)
cr = rng.integers(16, 241, (3, 24, 32), dtype=np.uint8)
path = os.path.join(tempfile.mkdtemp(), "selftest.mp4")
sv.write_h264_mp4(path, y, cb, cr, fps=(30, 1), skip_every=2)
ts, frames = nv.extract_frames_nvdec(path, fps=30.0, device=dev)
assert frames.shape == (6, 48, 64, 3), tuple(frames.shape)
got = frames.cpu().numpy().astype(np.int16)
Y = (255 / 219) * (y.astype(np.float64) - 16)
Cb = cb.repeat(2, 1).repeat(2, 2).astype(np.float64) - 128
Cr = cr.repeat(2, 1).repeat(2, 2).astype(np.float64) - 128
k = 255 / 224
want = np.clip(np.rint(np.stack([Y + 1.402 * k * Cr, Y - 0.344136 * k * Cb - 0.714136 * k * Cr, Y + 1.772 * k * Cb], -1)), 0, 255)
for j in range(6):
    assert np.abs(got[j] - want[j // 2]).max() <= 1, j
print("nvdec self-test ok")
"""


def self_test(device="cuda") -> tuple[bool, str]:
    """Power-on self-test of the hardware-decode path, once per process and device, in a CHILD process: a 64 x 48
    known-answer H.264 clip (every macroblock I_PCM) is decoded on the engine and compared with the planes that were
    written into it.  `embed_video(decoder="auto")` only trusts NVDEC after this passes — the binding declares the
    driver's structs without the SDK headers (csrc/cuvid_abi.h), and a layout that did not match the installed driver
    must cost a fallback to the software feed, not a crash or wrong pixels in the caller's process."""
    import subprocess
    import sys
    dev = str(resolve_device(device))
    if dev not in _SELF_TEST:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        try:
            r = subprocess.run([sys.executable, "-c", _SELF_TEST_SCRIPT, dev], capture_output=True, text=True, timeout=180,
                               env=env)
            ok = r.returncode == 0 and "nvdec self-test ok" in r.stdout
            why = "" if ok else f"self-test failed (exit {r.returncode}): {(r.stderr or r.stdout).strip()[-300:]}"
        except Exception as exc:  # timeout, no interpreter, ...
            ok, why = False, f"self-test could not run: {exc}"
        _SELF_TEST[dev] = (ok, why)
    return _SELF_TEST[dev]


# ------------------------------------------------------------------------------------------------ ISO-BMFF reader
def _boxes(buf, start: int, end: int) -> Iterator[tuple[bytes, int, int]]:
    """(type, payload_start, payload_end) of the boxes in buf[start:end]."""
    pos = start
    while pos + 8 <= end:
        size, kind = struct.unpack_from(">I4s", buf, pos)
        head = 8
        if size == 1:
            size = struct.unpack_from(">Q", buf, pos + 8)[0]
            head = 16
        elif size == 0:
            size = end - pos
        if size < head or pos + size > end:
            raise ValueError(f"corrupt box '{kind.decode('latin1')}' at offset {pos}")
        yield kind, pos + head, pos + size
        pos += size


def _child(buf, start: int, end: int, kind: bytes) -> Optional[tuple[int, int]]:
    for k, a, b in _boxes(buf, start, end):
        if k == kind:
            return a, b
    return None


@dataclass
class VideoTrack:
    codec: int                      # cuvid codec id
    width: int
    height: int
    timescale: int
    sample_offsets: np.ndarray      # int64 [n] file offsets
    sample_sizes: np.ndarray        # int64 [n]
    sample_deltas: np.ndarray       # int64 [n] durations in `timescale` units (decode order)
    nal_length_size: int = 4        # 0 = the samples already are Annex-B (raw elementary stream)
    parameter_sets: list = field(default_factory=list)  # SPS / PPS (/ VPS) NAL units without start codes

    @property
    def n_frames(self) -> int:
        return int(self.sample_sizes.shape[0])

    @property
    def avg_fps(self) -> float:
        total = int(self.sample_deltas.sum())
        return self.n_frames * self.timescale / total if total > 0 else 0.0


def read_mp4_video_track(path: str) -> VideoTrack:
    """The first H.264 / HEVC video track of an ISO-BMFF file (ISO/IEC 14496-12 sample tables, 14496-15 avcC / hvcC)."""
    size = os.path.getsize(path)
    with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as buf:
        moov = _child(buf, 0, size, b"moov")
        if moov is None:
            raise ValueError(f"{path}: no 'moov' box (not an MP4 / MOV file, or a fragmented one without a movie header)")
        for kind, a, b in _boxes(buf, *moov):
            if kind != b"trak":
                continue
            mdia = _child(buf, a, b, b"mdia")
            if mdia is None:
                continue
            hdlr = _child(buf, *mdia, b"hdlr")
            if hdlr is None or bytes(buf[hdlr[0] + 8:hdlr[0] + 12]) != b"vide":
                continue
            mdhd = _child(buf, *mdia, b"mdhd")
            version = buf[mdhd[0]]
            timescale = struct.unpack_from(">I", buf, mdhd[0] + (20 if version == 1 else 12))[0]
            minf = _child(buf, *mdia, b"minf")
            stbl = _child(buf, *minf, b"stbl")
            tables = {k: (x, y) for k, x, y in _boxes(buf, *stbl)}
            # ---- sample description
            sa, sb = tables[b"stsd"]
            entry = sa + 8
            esize, fmt = struct.unpack_from(">I4s", buf, entry)
            width, height = struct.unpack_from(">HH", buf, entry + 32)
            codec, psets, nls = None, [], 4
            for k, x, y in _boxes(buf, entry + 86, entry + esize):
                if k == b"avcC" and fmt in (b"avc1", b"avc3"):
                    codec = CODEC_H264
                    nls = (buf[x + 4] & 3) + 1
                    pos = x + 5
                    for mask in (31, 255):  # SPS list, then PPS list
                        count = buf[pos] & mask
                        pos += 1
                        for _ in range(count):
                            ln = struct.unpack_from(">H", buf, pos)[0]
                            psets.append(bytes(buf[pos + 2:pos + 2 + ln]))
                            pos += 2 + ln
                elif k == b"hvcC" and fmt in (b"hvc1", b"hev1"):
                    codec = CODEC_HEVC
                    nls = (buf[x + 21] & 3) + 1
                    pos = x + 23
                    for _ in range(buf[x + 22]):
                        count = struct.unpack_from(">H", buf, pos + 1)[0]
                        pos += 3
                        for _ in range(count):
                            ln = struct.unpack_from(">H", buf, pos)[0]
                            psets.append(bytes(buf[pos + 2:pos + 2 + ln]))
                            pos += 2 + ln
            if codec is None:
                raise ValueError(f"{path}: video track codec '{fmt.decode('latin1')}' is not H.264 / HEVC")
            # ---- sample sizes
            za = tables[b"stsz"][0]
            uniform, n = struct.unpack_from(">II", buf, za + 4)
            sizes = (np.full(n, uniform, np.int64) if uniform else
                     np.frombuffer(buf, dtype=">u4", count=n, offset=za + 12).astype(np.int64))
            # ---- chunk offsets and the sample-to-chunk map
            if b"co64" in tables:
                ca = tables[b"co64"][0]
                chunks = np.frombuffer(buf, dtype=">u8", count=struct.unpack_from(">I", buf, ca + 4)[0], offset=ca + 8).astype(np.int64)
            else:
                ca = tables[b"stco"][0]
                chunks = np.frombuffer(buf, dtype=">u4", count=struct.unpack_from(">I", buf, ca + 4)[0], offset=ca + 8).astype(np.int64)
            ka = tables[b"stsc"][0]
            runs = np.frombuffer(buf, dtype=">u4", count=3 * struct.unpack_from(">I", buf, ka + 4)[0],
                                 offset=ka + 8).astype(np.int64).reshape(-1, 3)  # copies: nothing may reference the map
            per_chunk = np.zeros(len(chunks), np.int64)
            for i, (first, spc, _) in enumerate(runs):
                last = int(runs[i + 1][0]) - 1 if i + 1 < len(runs) else len(chunks)
                per_chunk[int(first) - 1:last] = int(spc)
            chunk_of = np.repeat(np.arange(len(chunks)), per_chunk)[:n]
            if chunk_of.shape[0] != n:
                raise ValueError(f"{path}: sample-to-chunk table covers {chunk_of.shape[0]} of {n} samples")
            first_of_chunk = np.concatenate([[0], np.cumsum(per_chunk)[:-1]])
            within = np.cumsum(sizes) - sizes
            offsets = chunks[chunk_of] + within - within[first_of_chunk[chunk_of]]
            # ---- durations
            ta = tables[b"stts"][0]
            tt = np.frombuffer(buf, dtype=">u4", count=2 * struct.unpack_from(">I", buf, ta + 4)[0],
                               offset=ta + 8).astype(np.int64).reshape(-1, 2)
            deltas = np.repeat(tt[:, 1], tt[:, 0])[:n]
            return VideoTrack(codec, int(width), int(height), int(timescale), offsets, sizes, deltas, nls, psets)
    raise ValueError(f"{path}: no video track")


def read_annexb_file(path: str, codec: int = CODEC_H264, fps: float = 30.0, step: int = 4096) -> VideoTrack:
    """A raw Annex-B elementary stream cut into `step`-byte pieces (the parser finds the pictures itself).  The frame
    count is unknown until decoded; `fps` supplies the time base the container would.  Test / tooling input: a stream
    whose pictures are much smaller than `step` / 30 bytes can overflow one feed's output slack."""
    size = os.path.getsize(path)
    offs = np.arange(0, size, step, dtype=np.int64)
    sizes = np.minimum(step, size - offs)
    return VideoTrack(codec, 0, 0, int(round(fps * 1000)), offs, sizes, np.full(len(offs), 1000, np.int64), 0, [])


def annexb_chunks(path: str, track: VideoTrack, samples_per_chunk: int) -> Iterator[tuple[bytes, int]]:
    """(Annex-B bytes, number of samples) in decode order; the parameter sets go in front of the first sample."""
    with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as buf:
        n, nls = track.n_frames, track.nal_length_size
        first = True
        for s0 in range(0, n, samples_per_chunk):
            parts = []
            if first:
                parts += [_START + p for p in track.parameter_sets]
                first = False
            for s in range(s0, min(n, s0 + samples_per_chunk)):
                pos, end = int(track.sample_offsets[s]), int(track.sample_offsets[s] + track.sample_sizes[s])
                if nls == 0:
                    parts.append(bytes(buf[pos:end]))
                    continue
                while pos + nls <= end:
                    ln = int.from_bytes(buf[pos:pos + nls], "big")
                    pos += nls
                    if ln <= 0 or pos + ln > end:
                        raise ValueError(f"{path}: corrupt NAL length in sample {s}")
                    parts.append(_START)
                    parts.append(bytes(buf[pos:pos + ln]))
                    pos += ln
            yield b"".join(parts), min(n, s0 + samples_per_chunk) - s0


# ------------------------------------------------------------------------------------------------ the device feed
class NvdecFeed:
    """Iterates uint8 RGB batches [b, H, W, 3] (b <= batch) that live in device memory, in presentation order, holding the
    frames the reference's sampling rule keeps.  `timestamps` (float64, idx / video_fps) is complete up front for
    container files.  The yielded tensor is a view into one of two ring buffers: it is valid until the next-but-one
    batch is requested (consume it — e.g. `EmbeddingPipeline.embed` — on the same stream before iterating further)."""

    SLACK = 32  # frames one feed may emit beyond `batch` (reorder queue flushes)

    def __init__(self, video_path: str, fps: float = 1.0, batch: int = 64, device="cuda", max_display_delay: int = 2,
                 raw_codec: Optional[int] = None, raw_fps: float = 30.0, colour_matrix: int = -1, full_range: int = -1):
        if not available():
            raise RuntimeError("NVDEC is not available: libnvcuvid.so.1 (part of the GPU driver) did not load")
        self.device = resolve_device(device)
        self.path, self.batch = video_path, int(batch)
        if raw_codec is not None:
            self.track = read_annexb_file(video_path, raw_codec, raw_fps)
            self.video_fps, total = float(raw_fps), None
        else:
            self.track = read_mp4_video_track(video_path)
            self.video_fps, total = self.track.avg_fps, self.track.n_frames
        if fps > self.video_fps:
            raise ValueError("fps exceeds the video frame rate")  # the reference divides by int(video_fps / fps) == 0
        self.interval = int(self.video_fps / fps)
        if total is not None:
            self.indices, self.timestamps = sample_plan(total, self.video_fps, fps)
        else:
            self.indices, self.timestamps = None, None
        self._handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().gvl_nvdec_open(self.track.codec, int(max_display_delay), ctypes.byref(self._handle)),
                       "gvl_nvdec_open")
        _lib.check(_lib.lib().gvl_nvdec_sampling(self._handle, 0, self.interval, int(colour_matrix), int(full_range)),
                   "gvl_nvdec_sampling")
        self.height, self.width = self.track.height, self.track.width
        self.frames_decoded = 0
        self.frames_kept = 0

    def close(self) -> None:
        if self._handle:
            _lib.lib().gvl_nvdec_close(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        if self.indices is None:
            raise TypeError("a raw elementary stream has no frame count before it is decoded")
        return -(-len(self.indices) // self.batch)

    def info(self) -> dict:
        arr = (ctypes.c_int32 * 8)()
        _lib.check(_lib.lib().gvl_nvdec_info(self._handle, arr), "gvl_nvdec_info")
        keys = ("coded_width", "coded_height", "width", "height", "fps_num", "fps_den", "matrix_coefficients", "full_range")
        return dict(zip(keys, list(arr)))

    def _probe_size(self, first_chunk: bytes) -> None:
        """Raw streams carry no container size: parse the first chunk without an output buffer is not possible (it may
        display frames), so read the size from the SPS through a throw-away parser."""
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().gvl_nvdec_open(self.track.codec, 0, ctypes.byref(h)), "gvl_nvdec_open")
        try:
            _lib.lib().gvl_nvdec_sampling(h, 1 << 60, 1, -1, -1)  # keep nothing
            _lib.check(_lib.lib().gvl_nvdec_feed(h, first_chunk, len(first_chunk), 0, None, 0, 0, 0, None, None, None),
                       "gvl_nvdec_feed (probe)")
            arr = (ctypes.c_int32 * 8)()
            _lib.check(_lib.lib().gvl_nvdec_info(h, arr), "gvl_nvdec_info")
            self.width, self.height = int(arr[2]), int(arr[3])
        finally:
            _lib.lib().gvl_nvdec_close(h)

    def __iter__(self) -> Iterator[torch.Tensor]:
        lib = _lib.lib()
        B, S = self.batch, self.SLACK
        # a feed of k samples can display at most k + (frames the decoder still holds) pictures; keep the kept ones < S
        per_feed = 1 if self.track.nal_length_size == 0 else max(1, min(64, (S - 2) * self.interval - 20))
        chunks = annexb_chunks(self.path, self.track, per_feed)
        pending = None
        if self.height == 0:
            pending = next(chunks, None)
            if pending is None:
                return
            self._probe_size(pending[0])
        H, W = self.height, self.width
        with torch.cuda.device(self.device):
            ring = [torch.empty((B + S, H, W, 3), dtype=torch.uint8, device=self.device) for _ in range(2)]
            cur, filled = 0, 0
            stream = torch.cuda.current_stream(self.device)
            written, shown = ctypes.c_int(0), ctypes.c_longlong(0)

            def feed(data: bytes, eos: int):
                nonlocal filled
                buf = ring[cur]
                _lib.check(lib.gvl_nvdec_feed(self._handle, data, len(data), eos, buf[filled:].data_ptr(), B + S - filled, H, W,
                                              ctypes.byref(written), ctypes.byref(shown), int(stream.cuda_stream)),
                           "gvl_nvdec_feed")
                filled += written.value
                self.frames_decoded = int(shown.value)

            def drain(final: bool):
                nonlocal cur, filled
                while filled >= B or (final and filled > 0):
                    n = min(B, filled)
                    out = ring[cur][:n]
                    rest = filled - n
                    if rest:  # frames beyond the batch move to the front of the other buffer
                        ring[cur ^ 1][:rest].copy_(ring[cur][n:filled])
                    cur ^= 1
                    filled = rest
                    self.frames_kept += n
                    yield out

            while True:
                item = pending if pending is not None else next(chunks, None)
                pending = None
                if item is None:
                    break
                feed(item[0], 0)
                yield from drain(False)
            feed(b"", 1)
            yield from drain(True)
        if self.indices is None:
            self.timestamps = np.arange(self.frames_kept, dtype=np.float64) * self.interval / self.video_fps


def extract_frames_nvdec(video_path: str, fps: float = 1.0, device="cuda", **kw) -> tuple[np.ndarray, torch.Tensor]:
    """(timestamps float64 [n], frames uint8 [n, H, W, 3] on the device): the result of the reference's `extract_frames`
    (scripts/extract_features.py:230-264) without the host round trip.  Holds all sampled frames in device memory
    (6.2 MB per 1080p frame) — for long videos iterate `NvdecFeed` and embed batch by batch instead."""
    feed = NvdecFeed(video_path, fps=fps, device=device, **kw)
    parts = [b.clone() for b in feed]
    feed.close()
    frames = torch.cat(parts, 0) if parts else torch.empty((0, feed.height, feed.width, 3), dtype=torch.uint8, device=feed.device)
    return np.asarray(feed.timestamps[:frames.shape[0]]), frames
