"""Host side of the hardware-decoded ingest on the CPU: the known-answer H.264 writer is a valid
stream (OpenCV's ffmpeg decodes it to exactly the stored luma), the ISO-BMFF reader recovers its samples, the Annex-B
conversion is what a decoder accepts, and the library reports NVDEC's absence instead of crashing.  No GPU."""
import os

import numpy as np
import pytest

from gameplay_vision_llm_b200 import _lib, synth
from gameplay_vision_llm_b200 import nvdec_ingest as nv
from gameplay_vision_llm_b200 import synth_video as sv

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def clip(tmp_path_factory):
    d = tmp_path_factory.mktemp("h264")
    rgb = np.concatenate([synth.scene_frames_np(30 * i, 1, 100, 176) for i in range(9)])
    y, cb, cr = sv.rgb_to_ycbcr420(rgb)
    mp4, es = str(d / "clip.mp4"), str(d / "clip.h264")
    assert sv.write_h264_mp4(mp4, y, cb, cr, fps=(30, 1), skip_every=2) == 18
    assert sv.write_h264_annexb(es, y, cb, cr, fps=(30, 1), skip_every=2) == 18
    return mp4, es, y, cb, cr, rgb


def _decode(path, raw=False):
    cap = cv2.VideoCapture(path)
    if raw:
        cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
    out = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        out.append(f.copy())
    meta = (cap.get(cv2.CAP_PROP_FPS), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)))
    cap.release()
    return out, meta


def test_synthetic_stream_is_valid_h264_with_the_known_answer(clip):
    mp4, es, y, cb, cr, rgb = clip
    frames, (fps, w, h) = _decode(mp4)
    if len(frames) != 18:
        pytest.skip("this OpenCV build has no H.264 decoder")
    assert (fps, w, h) == (30.0, 176, 100)  # container frame rate, SPS cropping 112 -> 100 rows
    luma, _ = _decode(mp4, raw=True)     # raw mode hands back the decoder's luma plane
    if luma[0].shape == (100, 176):
        for j in range(18):
            assert np.array_equal(luma[j], y[j // 2]), f"frame {j}: I_PCM luma must be lossless"
    # colour: swscale BT.601 limited range, within its table rounding of the source up to chroma subsampling
    for j in (0, 1, 16, 17):
        assert np.abs(frames[j][..., ::-1].astype(int) - rgb[j // 2].astype(int)).mean() < 8
    # the raw elementary stream decodes to the same pictures
    es_frames, _ = _decode(es)
    assert len(es_frames) == 18 and all(np.array_equal(a, b) for a, b in zip(frames, es_frames))


def test_mp4_reader_recovers_the_sample_tables(clip):
    mp4, es, y, *_ = clip
    t = nv.read_mp4_video_track(mp4)
    assert (t.codec, t.width, t.height, t.n_frames, t.nal_length_size) == (nv.CODEC_H264, 176, 100, 18, 4)
    assert t.avg_fps == 30.0 and t.timescale == 30000 and (t.sample_deltas == 1000).all()
    assert [p[0] & 31 for p in t.parameter_sets] == [7, 8]  # SPS, PPS
    coded = 11 * 7 * 384  # 176 x 112 in macroblocks, 384 raw bytes each
    assert (t.sample_sizes[0::2] > coded).all() and (t.sample_sizes[1::2] < 32).all()  # I_PCM pictures / skipped pictures
    assert (np.diff(t.sample_offsets) == t.sample_sizes[:-1]).all()  # one contiguous chunk
    with open(mp4, "rb") as f:
        f.seek(int(t.sample_offsets[1]))
        ln = int.from_bytes(f.read(4), "big")
        assert ln == t.sample_sizes[1] - 4 and (f.read(1)[0] & 31) == 1  # a non-IDR slice NAL


@pytest.mark.parametrize("per_chunk", [1, 5, 64])
def test_annexb_conversion_is_what_a_decoder_accepts(clip, tmp_path, per_chunk):
    mp4 = clip[0]
    t = nv.read_mp4_video_track(mp4)
    chunks = list(nv.annexb_chunks(mp4, t, per_chunk))
    assert sum(n for _, n in chunks) == 18 and len(chunks) == -(-18 // per_chunk)
    stream = b"".join(c for c, _ in chunks)
    assert stream.startswith(b"\x00\x00\x00\x01\x67") and stream.count(b"\x00\x00\x00\x01") == 2 + 18
    out = tmp_path / "demuxed.h264"
    out.write_bytes(stream)
    want, _ = _decode(mp4)
    got, _ = _decode(str(out))
    if len(want) != 18:
        pytest.skip("this OpenCV build has no H.264 decoder")
    assert len(got) == 18 and all(np.array_equal(a, b) for a, b in zip(want, got))


def test_reader_rejects_what_it_cannot_demux(tmp_path):
    bad = tmp_path / "x.mp4"
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(ValueError):
        nv.read_mp4_video_track(str(bad))
    avi = tmp_path / "x.avi"
    w = cv2.VideoWriter(str(avi), cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (96, 64))
    w.write(np.zeros((64, 96, 3), np.uint8))
    w.release()
    with pytest.raises(ValueError):
        nv.read_mp4_video_track(str(avi))


def test_library_reports_nvdec_absence_without_a_gpu():
    lib = _lib.lib()
    assert lib.gvl_nvdec_available() in (0, 1)
    if not nv.available():  # this container: no GPU driver at all
        ok, why = nv.usable()
        assert not ok and "libnvcuvid" in why
        with pytest.raises(RuntimeError, match="NVDEC is not available"):
            nv.NvdecFeed("whatever.mp4")


def test_mp4_reader_parses_an_hevc_sample_entry(tmp_path):
    """The hvcC branch (ISO/IEC 14496-15 §8.3.3): parameter-set arrays, NAL length size, 32-bit chunk offsets and a
    multi-run sample-to-chunk table — on a hand-built file (dummy NAL payloads: only the container is under test)."""
    import struct

    from gameplay_vision_llm_b200.synth_video import _box, _full
    vps, sps, pps = b"\x40\x01VPS", b"\x42\x01SPS-", b"\x44\x01PPS--"
    arrays = b"".join(bytes([0x80 | t]) + struct.pack(">H", 1) + struct.pack(">H", len(n)) + n
                      for t, n in ((32, vps), (33, sps), (34, pps)))
    hvcc = _box(b"hvcC", bytes(21) + bytes([0xFC | 2]) + bytes([3]) + arrays)  # lengthSizeMinusOne = 2 -> 3-byte lengths
    entry = _box(b"hvc1", b"\x00" * 6, struct.pack(">H", 1), b"\x00" * 16, struct.pack(">HH", 640, 360),
                 struct.pack(">II", 0x00480000, 0x00480000), b"\x00" * 4, struct.pack(">H", 1), b"\x00" * 32,
                 struct.pack(">Hh", 0x18, -1), hvcc)
    nal = lambda payload: len(payload).to_bytes(3, "big") + payload  # noqa: E731
    samples = [nal(b"\x26\x01" + bytes([i]) * (5 + i)) for i in range(5)]
    sizes = [len(x) for x in samples]
    ftyp = _box(b"ftyp", b"isom", struct.pack(">I", 0x200), b"isomiso2hvc1")

    def moov(off0, off1):
        stbl = _box(b"stbl", _full(b"stsd", 0, struct.pack(">I", 1), entry),
                    _full(b"stts", 0, struct.pack(">IIIII", 2, 2, 1001, 3, 2002)),      # two duration runs
                    _full(b"stsc", 0, struct.pack(">IIIIIII", 2, 1, 2, 1, 2, 3, 1)),    # chunk 1: 2 samples, chunk 2: 3
                    _full(b"stsz", 0, struct.pack(">II", 0, 5), b"".join(struct.pack(">I", z) for z in sizes)),
                    _full(b"stco", 0, struct.pack(">III", 2, off0, off1)))
        minf = _box(b"minf", _full(b"vmhd", 1, b"\x00" * 8), stbl)
        mdhd = _full(b"mdhd", 1 << 24, struct.pack(">QQIQHH", 0, 0, 24000, 8008, 0x55C4, 0))  # version 1: 64-bit times
        hdlr = _full(b"hdlr", 0, struct.pack(">I4s", 0, b"vide"), b"\x00" * 12, b"V\x00")
        audio = _box(b"trak", _box(b"mdia", _full(b"hdlr", 0, struct.pack(">I4s", 0, b"soun"), b"\x00" * 12, b"A\x00")))
        return _box(b"moov", audio, _box(b"trak", _box(b"mdia", mdhd, hdlr, minf)))

    gap = b"free-space"
    head = len(ftyp) + len(moov(0, 0)) + 8
    off0 = head
    off1 = head + sizes[0] + sizes[1] + len(gap)
    path = tmp_path / "h.mp4"
    path.write_bytes(ftyp + moov(off0, off1) + struct.pack(">I4s", 8 + sum(sizes) + len(gap), b"mdat") +
                     samples[0] + samples[1] + gap + samples[2] + samples[3] + samples[4])
    t = nv.read_mp4_video_track(str(path))
    assert (t.codec, t.width, t.height, t.timescale, t.n_frames, t.nal_length_size) == (nv.CODEC_HEVC, 640, 360, 24000, 5, 3)
    assert t.parameter_sets == [vps, sps, pps]
    assert t.sample_deltas.tolist() == [1001, 1001, 2002, 2002, 2002] and abs(t.avg_fps - 5 * 24000 / 8008) < 1e-12
    assert t.sample_offsets.tolist() == [off0, off0 + sizes[0], off1, off1 + sizes[2], off1 + sizes[2] + sizes[3]]
    stream = b"".join(c for c, _ in nv.annexb_chunks(str(path), t, 2))
    start = b"\x00\x00\x00\x01"
    want = start + vps + start + sps + start + pps + b"".join(start + b"\x26\x01" + bytes([i]) * (5 + i) for i in range(5))
    assert stream == want
