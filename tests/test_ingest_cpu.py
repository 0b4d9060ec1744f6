"""Frame ingest: the reference's sampling rule and a bounded-memory pinned feed, on CPU."""
import numpy as np
import pytest
import torch

cv2 = pytest.importorskip("cv2")

from gameplay_vision_llm_b200 import frame_ingest as fi  # noqa: E402


def reference_sampling(total_frames, video_fps, fps):
    """scripts/extract_features.py:247-253, literally."""
    frame_interval = int(video_fps / fps)
    sample_indices = list(range(0, total_frames, frame_interval))
    return sample_indices, [idx / video_fps for idx in sample_indices]


@pytest.mark.parametrize("total,video_fps,fps", [(300, 30.0, 1.0), (301, 29.97, 2.0), (7, 60.0, 1.0), (100, 24.0, 24.0),
                                                  (1000, 59.94, 0.5), (0, 30.0, 1.0)])
def test_sample_plan_matches_reference_rule(total, video_fps, fps):
    idx, ts = fi.sample_plan(total, video_fps, fps)
    want_idx, want_ts = reference_sampling(total, video_fps, fps)
    assert idx.tolist() == want_idx and ts.tolist() == want_ts


def test_sample_plan_raises_like_the_reference_when_fps_exceeds_video_fps():
    with pytest.raises(ValueError):
        reference_sampling(10, 30.0, 60.0)
    with pytest.raises(ValueError):
        fi.sample_plan(10, 30.0, 60.0)


@pytest.fixture(scope="module")
def tiny_video(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("vid") / "clip.avi")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (96, 64))
    assert w.isOpened()
    rng = np.random.default_rng(0)
    for i in range(95):
        f = np.full((64, 96, 3), (i * 2) % 256, np.uint8)
        f[:, : (i * 3) % 96, 1] = 255 - i
        f[rng.integers(0, 64), :, 2] = 200
        w.write(f)
    w.release()
    return path


def decode_all_rgb(path):
    cap = cv2.VideoCapture(path)
    out = []
    while True:
        ok, bgr = cap.read()
        if not ok:
            break
        out.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    cap.release()
    return out


def test_extract_frames_returns_the_sampled_frames_and_timestamps(tiny_video):
    every = decode_all_rgb(tiny_video)
    assert len(every) == 95
    got = fi.extract_frames(tiny_video, fps=2.0, as_pil=False)      # interval int(30 / 2) = 15 -> frames 0, 15, ..., 90
    assert [t for t, _ in got] == [i / 30.0 for i in range(0, 95, 15)]
    for (t, frame), i in zip(got, range(0, 95, 15)):
        assert np.array_equal(frame, every[i])
    pil = fi.extract_frames(tiny_video, fps=2.0)
    assert pil[1][1].size == (96, 64) and pil[1][1].mode == "RGB" and np.array_equal(np.asarray(pil[1][1]), every[15])


@pytest.mark.parametrize("batch,depth,auto", [(4, 2, True), (3, 3, True), (4, 3, False), (64, 2, True)])
def test_frame_feed_batches_in_timeline_order_with_a_bounded_ring(tiny_video, batch, depth, auto):
    every = decode_all_rgb(tiny_video)
    feed = fi.FrameFeed(tiny_video, fps=10.0, batch=batch, depth=depth, pin=False, auto_release=auto)  # every 3rd frame
    want = list(range(0, 95, 3))
    assert feed.timestamps.tolist() == [i / 30.0 for i in want] and len(feed) == -(-len(want) // batch)
    seen = 0
    for hb in feed:
        assert hb.dtype == torch.uint8 and hb.shape[1:] == (64, 96, 3) and hb.shape[0] <= batch
        for k in range(hb.shape[0]):
            assert np.array_equal(hb[k].numpy(), every[want[seen + k]])
        seen += hb.shape[0]
        if not auto:
            feed.recycle(hb)            # explicit hand-back (what embed_stream does, with a CUDA event)
    assert seen == len(want) == feed.frames_decoded
    assert len(feed._ring) == depth     # memory stays at `depth` batches however long the video is


def test_frame_feed_surfaces_decoder_errors(tmp_path):
    with pytest.raises(RuntimeError):
        fi.FrameFeed(str(tmp_path / "missing.mp4"))
