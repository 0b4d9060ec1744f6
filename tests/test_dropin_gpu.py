"""The reference-facing classes on the GPU: same call sequence as the reference's scripts, checked against
the oracle; plus the streaming pipeline and the timeline index."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gameplay_vision_llm_b200 import ops, synth  # noqa: E402
from gameplay_vision_llm_b200.pipeline import EmbeddingPipeline, pinned_batches  # noqa: E402
from gameplay_vision_llm_b200.projector import ProjectorBank, ReasoningCoreConfig, project_embeddings  # noqa: E402
from gameplay_vision_llm_b200.siglip_semantic_encoder import (NaFlexConfig, SemanticEmbedding,  # noqa: E402
                                                               SigLIPSemanticEncoder)
from gameplay_vision_llm_b200.timeline import TimelineEmbeddingIndex  # noqa: E402
from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_projector_state_dict,  # noqa: E402
                                                synth_siglip_state_dict)
from oracle import preprocess_ref, siglip_ref  # noqa: E402

DEV = "cuda:0"
SPEC = SiglipVisionSpec(hidden=216, intermediate=400, layers=2, heads=3, image=140, patch=14)


def _encoder():
    sd = synth_siglip_state_dict(SPEC, seed=0)
    enc = SigLIPSemanticEncoder(NaFlexConfig(device=DEV, state_dict=sd, base_resolution=SPEC.image, embedding_dim=SPEC.hidden,
                                             batch_size=4))
    return enc, sd


def test_encode_image_reference_call_sequence():
    from PIL import Image
    enc, sd = _encoder()
    frames = synth.scene_frames_np(0, 3, 270, 480)
    want = siglip_ref.vision_forward(sd, torch.from_numpy(preprocess_ref.pixel_values(frames, SPEC.image, SPEC.image, 2)),
                                     SPEC.heads, SPEC.patch, SPEC.eps)
    for i in range(3):
        emb = enc.encode_image(Image.fromarray(frames[i]))          # extract_features.py:595
        assert emb.shape == (SPEC.hidden,) and emb.dtype == torch.bfloat16 and emb.is_cuda
        cos = torch.nn.functional.cosine_similarity(emb.float().cpu(), want[i], dim=0)
        assert cos > 0.999
    # the inner seam the reference uses (:474-481)
    inputs = enc.encoder._processor(images=[Image.fromarray(frames[0])], return_tensors="pt").to(enc.encoder._model.device)
    pv = inputs["pixel_values"]
    assert np.array_equal(pv.cpu().numpy(), preprocess_ref.pixel_values(frames[:1], SPEC.image, SPEC.image, 2))
    batched = enc.encode_frames(frames)
    single = torch.stack([enc.encode_image(Image.fromarray(f)) for f in frames])
    assert torch.nn.functional.cosine_similarity(batched.float(), single.float(), dim=1).min() > 0.9999


def test_find_similar_regions_order_and_ties():
    enc, _ = _encoder()
    g = torch.Generator().manual_seed(0)
    vecs = torch.randn(20, SPEC.hidden, generator=g)
    vecs[7] = vecs[2]
    cands = [SemanticEmbedding(embedding=v.to(torch.bfloat16), entity_id=str(i)) for i, v in enumerate(vecs)]
    res = enc.find_similar_regions(cands[2], cands, top_k=5)
    assert [r[0].entity_id for r in res[:2]] == ["2", "7"] and abs(res[0][1] - 1.0) < 1e-3
    want = siglip_ref.cosine_topk(vecs.to(torch.bfloat16).float().numpy(), vecs[2:3].to(torch.bfloat16).float().numpy(), 5)[1][0]
    assert [int(r[0].entity_id) for r in res] == want.tolist()
    assert abs(enc.compute_similarity(cands[0], cands[1]) -
               float(torch.nn.functional.cosine_similarity(vecs[0].bfloat16().float(), vecs[1].bfloat16().float(), dim=0))) < 1e-4


def test_projector_bank_matches_reference_module(tmp_path):
    bank = ProjectorBank(ReasoningCoreConfig(device=DEV))
    bank.siglip_proj.load_state_dict(synth_projector_state_dict(1152, 4096, 1))
    bank.videomae_proj.load_state_dict(synth_projector_state_dict(768, 4096, 2))
    path = tmp_path / "projector_weights.pt"
    bank.save_weights(str(path))
    bank2 = ProjectorBank(ReasoningCoreConfig(device=DEV)).to(DEV)
    bank2.load_weights(str(path))
    g = torch.Generator().manual_seed(3)
    x = torch.randn(10, 1152, generator=g).to(torch.bfloat16)          # embeddings as stored in *_embeddings.pt
    out = project_embeddings(bank2, region_embeddings=x, videomae_embeddings=torch.randn(2, 5, 768, generator=g))
    assert out["siglip"].shape == (10, 4096) and out["siglip"].dtype == torch.float32
    assert out["videomae"].shape == (2, 5, 4096)
    want = siglip_ref.projector_forward(synth_projector_state_dict(1152, 4096, 1), x.float())
    cos = torch.nn.functional.cosine_similarity(out["siglip"].cpu(), want, dim=1)
    assert cos.min() > 0.9999 and (out["siglip"].cpu() - want).abs().max() < 0.05


def test_pipeline_stream_equals_resident_and_index_search():
    sd = synth_siglip_state_dict(SPEC, seed=0)
    psd = synth_projector_state_dict(SPEC.hidden, 512, seed=1)
    pipe = EmbeddingPipeline(sd, psd, SPEC, DEV, batch=4)
    frames = synth.scene_frames_np(0, 10, 180, 320, frames_per_scene=5)
    idx = TimelineEmbeddingIndex(10, 512, fps=2.0, device=DEV)
    pipe.embed_resident(torch.from_numpy(frames).to(DEV), index=idx.local_rows())
    resident = idx.index().clone()
    host_out = torch.empty((10, 512), dtype=torch.bfloat16).pin_memory()
    streamed = torch.zeros((10, 512), dtype=torch.bfloat16, device=DEV)
    n = pipe.embed_stream(pinned_batches(frames, 4), streamed, host_out)
    torch.cuda.synchronize()
    assert n == 10 and torch.equal(streamed, resident) and torch.equal(host_out, resident.cpu())
    # oracle embeddings -> same nearest neighbours
    want = siglip_ref.projector_forward(psd, siglip_ref.vision_forward(
        sd, torch.from_numpy(preprocess_ref.pixel_values(frames, SPEC.image, SPEC.image, 2)), SPEC.heads, SPEC.patch, SPEC.eps))
    assert torch.nn.functional.cosine_similarity(resident.float().cpu(), want, dim=1).min() > 0.999
    scores, ids = idx.search(resident[3:4], top_k=3)
    assert ids[0, 0].item() == 3
    hits = idx.retrieve_by_semantic(resident[6], top_k=2)
    assert hits[0][0] == 3.0  # frame 6 at 2 fps
    sel, emb = idx.window(2.0, window_sec=1.0)
    assert sel.tolist() == [3, 4, 5] and emb.shape == (3, 512)
    sel, emb = idx.window(None)          # realtime_inference.py:1000: no timestamp -> the first 20 rows
    assert sel.tolist() == list(range(10))
    # retrieval service: time window fused with the ranking, and hybrid_retrieve's merge order
    e64, ts = resident.float().cpu().numpy(), idx.timestamps
    q = resident[6:7]
    lo, hi = siglip_ref.window_rows(ts, 2.0, 1.0, True)            # rows 2..6 (inclusive bounds)
    assert idx.retrieve_by_timestamp(2.0, 1.0) == list(range(lo, hi)) == [2, 3, 4, 5, 6]
    _, want_i = siglip_ref.cosine_topk_windowed(e64, q.float().cpu().numpy(), 3, [lo], [hi])
    _, got_i = idx.search(q, top_k=3, windows=[(2.0, 1.0)], inclusive=True)
    assert got_i[0].tolist() == want_i[0].tolist() and got_i[0, 0].item() == 6
    _, sem5 = siglip_ref.cosine_topk(e64, q.float().cpu().numpy(), 5)[:2]
    assert idx.hybrid_retrieve(q, timestamp=0.5, window=0.5) == siglip_ref.hybrid_merge([0, 1, 2], sem5[0].tolist())
    assert idx.hybrid_retrieve(q) == siglip_ref.cosine_topk(e64, q.float().cpu().numpy(), 10)[1][0].tolist()


def test_embed_video_matches_embedding_the_decoded_frames(tmp_path):
    """video file -> sampled frames (reference rule) -> pinned ring -> device, overlapped; the result must
    equal embedding the same decoded frames from a resident tensor, with the reference's timestamps."""
    cv2 = pytest.importorskip("cv2")
    from gameplay_vision_llm_b200 import frame_ingest as fi
    path = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (320, 180))
    frames = synth.scene_frames_np(0, 50, 180, 320, frames_per_scene=10)
    for f in frames:
        w.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    w.release()
    sd = synth_siglip_state_dict(SPEC, seed=0)
    psd = synth_projector_state_dict(SPEC.hidden, 512, seed=1)
    pipe = EmbeddingPipeline(sd, psd, SPEC, DEV, batch=4)
    ts, index = pipe.embed_video(path, fps=5.0)          # every 6th frame: 0, 6, ..., 48 -> 9 frames, batches 4 + 4 + 1
    assert ts.tolist() == [i / 30.0 for i in range(0, 50, 6)] and index.shape == (9, 512)
    decoded = np.stack([f for _, f in fi.extract_frames(path, fps=5.0, as_pil=False)])
    want = pipe.embed_resident(torch.from_numpy(decoded).to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(index, want)


def test_one_streaming_pass_yields_index_and_reference_cache_files(tmp_path):
    """the reference caches the 1152-d SigLIP vector (scripts/extract_features.py:597-603,
    1447-1468), not the projection — `embed_stream` hands both out in one pass (pooled rows via the D2H side stream),
    and `write_caches` turns a video into `{stem}_embeddings.pt` + the PerceptionCache directory."""
    cv2 = pytest.importorskip("cv2")
    from gameplay_vision_llm_b200 import feature_cache as fc
    sd = synth_siglip_state_dict(SPEC, seed=0)
    psd = synth_projector_state_dict(SPEC.hidden, 512, seed=1)
    pipe = EmbeddingPipeline(sd, psd, SPEC, DEV, batch=4)
    frames = synth.scene_frames_np(0, 10, 180, 320, frames_per_scene=5)
    dev_frames = torch.from_numpy(frames).to(DEV)
    want_pooled = torch.cat([pipe.embed(dev_frames[i:i + 4])[0].clone() for i in range(0, 10, 4)])
    want_index = pipe.embed_resident(dev_frames).clone()
    index = torch.zeros((10, 512), dtype=torch.bfloat16, device=DEV)
    pooled = torch.zeros((10, SPEC.hidden), dtype=torch.bfloat16, device=DEV)
    host_idx = torch.zeros((10, 512), dtype=torch.bfloat16).pin_memory()
    host_pooled = torch.zeros((10, SPEC.hidden), dtype=torch.bfloat16).pin_memory()
    n = pipe.embed_stream(pinned_batches(frames, 4), index, host_idx, pooled_out=pooled, host_pooled=host_pooled)
    torch.cuda.current_stream().synchronize()  # the compute stream waits for the D2H side stream
    assert n == 10 and torch.equal(index, want_index) and torch.equal(pooled, want_pooled)
    assert torch.equal(host_idx, want_index.cpu()) and torch.equal(host_pooled, want_pooled.cpu())
    # host_pooled alone (no device buffer passed) works too
    host_pooled2 = torch.zeros_like(host_pooled)
    pipe.embed_stream(pinned_batches(frames, 4), index, host_pooled=host_pooled2)
    torch.cuda.current_stream().synchronize()
    assert torch.equal(host_pooled2, want_pooled.cpu())
    # a video file -> cache files in the reference's layouts
    path = str(tmp_path / "clip.avi")
    w = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (320, 180))
    for f in synth.scene_frames_np(0, 20, 180, 320, frames_per_scene=10):
        w.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    w.release()
    ts, idx2, pooled2 = pipe.write_caches(path, str(tmp_path / "out"), fps=10.0, perception_cache_dir=str(tmp_path / "pc"))
    assert len(ts) == 7 and idx2.shape == (7, 512) and pooled2.shape == (7, SPEC.hidden)
    data = torch.load(tmp_path / "out" / "clip_embeddings.pt", weights_only=False)
    assert [e["timestamp"] for e in data["siglip"]] == ts.tolist()
    assert torch.equal(torch.stack([e["embedding"] for e in data["siglip"]]), pooled2.cpu())
    back = fc.load_perception_npz(str(tmp_path / "pc"), path)
    assert np.array_equal(back["siglip_embeddings"], pooled2.float().cpu().numpy())
    assert back["metadata"]["has_siglip"] and back["metadata"]["num_frames"] == 7


def test_timeline_norm_cache_follows_appended_rows():
    """rows written through `local_rows()` after a first search must be ranked with fresh 1/|e| values
    (the streaming / realtime use) — tensor path, 4096+ rows so that AUTO picks it."""
    g = torch.Generator().manual_seed(2)
    n, d = 4200, 256
    rows = (torch.randn(n, d, generator=g) * torch.rand(n, 1, generator=g) * 4).to(torch.bfloat16).to(DEV)
    idx = TimelineEmbeddingIndex(n, d, device=DEV)
    idx.local_rows()[: n // 2].copy_(rows[: n // 2])  # second half still zero
    q = rows[n // 2 + 5: n // 2 + 21]                  # 16 queries that are NOT indexed yet
    _, first = idx.search(q, top_k=4, mode=ops.TOPK_TENSOR)
    assert (first[:, 0].cpu() != torch.arange(n // 2 + 5, n // 2 + 21)).all()
    idx.local_rows()[n // 2:].copy_(rows[n // 2:])    # append; local_rows() drops the cached norms
    _, second = idx.search(q, top_k=4, mode=ops.TOPK_TENSOR)
    _, want, _ = siglip_ref.cosine_topk(rows.float().cpu().numpy(), q.float().cpu().numpy(), 4)
    assert np.array_equal(second.cpu().numpy(), want)
    with pytest.raises(RuntimeError, match="top_k"):
        idx.search(q, top_k=65)


def test_ops_bind_to_the_tensors_device():
    """a pipeline on cuda:1 must launch on cuda:1 even when the process's current device is cuda:0
    (weights, streams and TMA descriptors all follow the tensors).  Needs two GPUs; on one GPU the same code path is
    exercised with the current device == the tensors' device, and a cross-device call must raise, not fault."""
    sd = synth_siglip_state_dict(SPEC, seed=0)
    psd = synth_projector_state_dict(SPEC.hidden, 512, seed=1)
    frames = synth.scene_frames_np(0, 3, 180, 320)
    pipe0 = EmbeddingPipeline(sd, psd, SPEC, "cuda:0", batch=4)
    want = pipe0.embed_resident(torch.from_numpy(frames).to("cuda:0")).cpu()
    assert EmbeddingPipeline(sd, psd, SPEC, "cuda", batch=4).device == torch.device("cuda", torch.cuda.current_device())
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible: the cuda:1 half needs --gpus 2")
    assert torch.cuda.current_device() == 0
    pipe1 = EmbeddingPipeline(sd, psd, SPEC, "cuda:1", batch=4)
    got = pipe1.embed_resident(torch.from_numpy(frames).to("cuda:1"))
    torch.cuda.synchronize("cuda:1")
    assert got.device == torch.device("cuda:1") and torch.equal(got.cpu(), want) and torch.cuda.current_device() == 0
    idx = TimelineEmbeddingIndex(3, 512, device="cuda:1")
    idx.local_rows().copy_(got)
    assert idx.search(got[1:2], top_k=1)[1].item() == 1
    with pytest.raises(RuntimeError, match="different devices|weights on"):
        pipe1.embed(torch.from_numpy(frames).to("cuda:0"))



def test_encode_image_accepts_small_images_like_the_reference():
    """`encode_image(region_img)` on a bounding-box crop (scripts/realtime_inference.py:281-295): the HF processor
    up-scales whatever it is given; so does the drop-in (general two-pass kernel), bit-identical pixel_values."""
    from PIL import Image
    spec = SiglipVisionSpec(hidden=216, intermediate=400, layers=2, heads=3, image=140, patch=14)
    sd = synth_siglip_state_dict(spec, seed=0)
    enc = SigLIPSemanticEncoder(NaFlexConfig(device=DEV, base_resolution=140, state_dict=sd, num_attention_heads=3))
    crop = synth.scene_frames_np(4, 1, 270, 480)[0][40:97, 100:231]  # 57 x 131: up-scaled vertically and horizontally
    emb = enc.encode_image(Image.fromarray(crop))
    pv = preprocess_ref.pixel_values(crop[None], 140, 140, 2)
    got_pv = enc.encoder._processor(images=[Image.fromarray(crop)], return_tensors="pt")["pixel_values"]
    assert np.array_equal(got_pv.cpu().numpy().view(np.uint32), pv.view(np.uint32))
    want = siglip_ref.vision_forward(sd, torch.from_numpy(pv), spec.heads, spec.patch, spec.eps)
    cos = torch.nn.functional.cosine_similarity(emb.float().cpu(), want[0], dim=0).item()
    assert cos > 0.999, cos


def test_extract_siglip_embeddings_equals_the_realtime_loop():
    """`pipeline.extract_siglip_embeddings` (scripts/realtime_inference.py:244-335) against the reference's loop written
    out literally on the same encoder: masked frames, PIL crops (including one that leaves the frame and a tiny one that
    is up-scaled), frames without detections, and the full-frame fallback."""
    from types import SimpleNamespace

    from PIL import Image

    from gameplay_vision_llm_b200.pipeline import extract_siglip_embeddings
    spec = SiglipVisionSpec(hidden=216, intermediate=400, layers=2, heads=3, image=140, patch=14)
    sd = synth_siglip_state_dict(spec, seed=0)
    enc = SigLIPSemanticEncoder(NaFlexConfig(device=DEV, base_resolution=140, state_dict=sd, num_attention_heads=3, batch_size=4))
    raw = synth.scene_frames_np(9, 4, 270, 480)
    pil = [Image.fromarray(f) for f in raw]
    mask = np.zeros((270, 480), np.bool_)
    mask[60:200, 100:300] = True
    sam = [
        {"timestamp": 0.0, "frame": pil[0], "detections": [{"mask": SimpleNamespace(mask=mask), "label": "player"},
                                                           {"bbox": (40, 30, 300, 120), "label": "hud"},
                                                           {"bbox": (400.4, 200.6, 520, 300)},      # leaves the frame
                                                           {"label": "whole"}]},
        {"timestamp": 0.5, "frame": pil[1], "detections": []},
        {"timestamp": 1.0, "frame": pil[2], "detections": [{"bbox": (10, 10, 50, 34), "label": "icon"},  # 40 x 24: up-scaled
                                                           {"bbox": (45, 35, 305, 125), "label": "hud"}]},
    ]
    got = extract_siglip_embeddings([(i * 0.5, p) for i, p in enumerate(pil)], DEV, sam_results=sam, encoder=enc)
    want = []
    for sf in sam:  # the reference's loop (:271-311), literally
        frame, frame_np = sf["frame"], np.array(sf["frame"])
        if sf["detections"]:
            for det in sf["detections"]:
                if det.get("mask") and hasattr(det["mask"], "mask"):
                    masked = frame_np.copy()
                    masked[~det["mask"].mask] = 0
                    region_img = Image.fromarray(masked)
                elif det.get("bbox"):
                    x1, y1, x2, y2 = det["bbox"]
                    region_img = frame.crop((x1, y1, x2, y2))
                else:
                    region_img = frame
                want.append({"timestamp": sf["timestamp"], "label": det.get("label", "region"),
                             "embedding": enc.encode_image(region_img).cpu()})
        else:
            want.append({"timestamp": sf["timestamp"], "label": "full_frame", "embedding": enc.encode_image(frame).cpu()})
    assert [(g["timestamp"], g["label"]) for g in got] == [(w["timestamp"], w["label"]) for w in want] and len(got) == 7
    assert all(torch.equal(g["embedding"], w["embedding"]) for g, w in zip(got, want))
    plain = extract_siglip_embeddings([(i * 0.5, p) for i, p in enumerate(pil)], DEV, encoder=enc)
    assert [d["label"] for d in plain] == ["full_frame"] * 4
    assert all(torch.equal(d["embedding"], enc.encode_image(p).cpu()) for d, p in zip(plain, pil))


def test_encode_images_equals_the_per_frame_loop():
    """`encode_images` (threaded PIL conversion into a pinned double buffer, batches of config.batch_size) must return
    what the reference's loop `[encode_image(f) for f in frames]` returns, bit for bit, across buffer reuse."""
    from PIL import Image
    spec = SiglipVisionSpec(hidden=216, intermediate=400, layers=2, heads=3, image=140, patch=14)
    sd = synth_siglip_state_dict(spec, seed=0)
    enc = SigLIPSemanticEncoder(NaFlexConfig(device=DEV, base_resolution=140, state_dict=sd, num_attention_heads=3, batch_size=4))
    raw = synth.scene_frames_np(11, 14, 270, 480)  # 14 frames: 3 full batches of 4 + a tail of 2, ring reused twice
    pil = [Image.fromarray(f) for f in raw]
    got = enc.encode_images(pil)
    want = torch.stack([enc.encode_image(p) for p in pil])
    assert got.shape == (14, spec.hidden) and torch.equal(got, want)
    assert torch.equal(enc.encode_images(list(raw)), want)           # arrays instead of PIL images
    assert torch.equal(enc.encode_images(pil[:2]), want[:2])         # the small-list path
    with pytest.raises(ValueError, match="one size"):
        enc.encode_images(pil[:4] + [Image.fromarray(raw[0][:100])])
