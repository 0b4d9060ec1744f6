"""VideoMAE clip route (SURVEY.md §8f.1 / BASELINE.json configs[3]) through the C ABI against the oracle and the
committed HF golden vectors: preprocessing bit-exact, embeddings within the stated bf16 tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from gameplay_vision_llm_b200 import ops, synth  # noqa: E402
from gameplay_vision_llm_b200.videomae_encoder import VideoMAEClipEncoder  # noqa: E402
from gameplay_vision_llm_b200.weights import (ProjectorPack, VideoMAEPack, VideoMAESpec,  # noqa: E402
                                                synth_projector_state_dict, synth_videomae_state_dict)
from oracle import preprocess_ref, siglip_ref, videomae_ref  # noqa: E402

DEV = "cuda:0"


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.double().cpu(), b.double().cpu(), dim=-1)


@pytest.mark.parametrize("H,W,size", [(1080, 1920, 224), (123, 211, 48), (720, 1280, 224), (400, 300, 64)])
def test_preprocess_crop_bit_exact(H, W, size):
    frames = synth.noise_frames(3, H, W, seed=H)
    oh, ow, y0, x0 = videomae_ref.resize_geometry(H, W, size, size)
    dev = frames.to(DEV)
    # the processor's class defaults (mean = std = 0.5) and the checkpoint's constants (ImageNet mean / std)
    for mean, std in (((0.5,) * 3, (0.5,) * 3), (videomae_ref.IMAGENET_MEAN, videomae_ref.IMAGENET_STD)):
        pv = videomae_ref.pixel_values(frames.numpy(), size, size, image_mean=mean, image_std=std)
        got = ops.preprocess_crop(dev, oh, ow, y0, x0, size, size, image_mean=mean, image_std=std,
                                  layout=ops.LAYOUT_F32_CHW).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), pv.view(np.uint32)), "fp32 pixel_values differ"
        got16 = ops.preprocess_crop(dev, oh, ow, y0, x0, size, size, image_mean=mean, image_std=std,
                                    layout=ops.LAYOUT_BF16_CHW).cpu()
        assert torch.equal(got16.view(torch.int16), torch.from_numpy(pv).to(torch.bfloat16).view(torch.int16))
    pv = videomae_ref.pixel_values(frames.numpy(), size, size, image_mean=(0.5,) * 3, image_std=(0.5,) * 3)
    u8 = ops.preprocess_crop(dev, oh, ow, y0, x0, size, size, layout=ops.LAYOUT_U8_CHW).cpu().numpy()
    assert np.array_equal(u8, np.rint(pv * 127.5 + 127.5).astype(np.uint8))


@pytest.mark.parametrize("H,W,size", [(1080, 1920, 224), (720, 1280, 224), (123, 211, 48), (300, 1000, 64)])
def test_preprocess_crop_band_equals_whole_frames(H, W, size):
    """Only the source-column band the center crop reads is moved to the device (one strided H2D copy of pinned whole
    frames): the band call must give the whole-frame call's bits, and the band must be what ATen's taps read."""
    frames = synth.noise_frames(3, H, W, seed=W)
    oh, ow, y0, x0 = videomae_ref.resize_geometry(H, W, size, size)
    bx0, bw = ops.crop_source_band(W, ow, x0, size)
    xmin, xsize, _, _ = preprocess_ref.resize_taps(W, ow, preprocess_ref.BILINEAR)
    lo, hi = int(xmin[x0]), int(xmin[x0 + size - 1] + xsize[x0 + size - 1])
    assert bx0 % 16 == 0 and lo - 16 < bx0 <= lo and hi <= bx0 + bw <= min(W, hi + 15)
    want = ops.preprocess_crop(frames.to(DEV), oh, ow, y0, x0, size, size, layout=ops.LAYOUT_BF16_CHW)
    band_dev = torch.empty((3, H, bw, 3), dtype=torch.uint8, device=DEV)
    ops.copy_band_h2d(band_dev, frames.pin_memory(), bx0)
    torch.cuda.synchronize()
    assert torch.equal(band_dev.cpu(), frames[:, :, bx0:bx0 + bw, :])
    got = ops.preprocess_crop(band_dev, oh, ow, y0, x0, size, size, layout=ops.LAYOUT_BF16_CHW, band=(bx0, W))
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    with pytest.raises(RuntimeError, match="band"):  # a band that misses columns the crop reads is refused
        ops.preprocess_crop(band_dev[:, :, 16:, :].contiguous(), oh, ow, y0, x0, size, size, layout=ops.LAYOUT_BF16_CHW,
                            band=(bx0 + 16, W))


def test_run_host_frames_band_feed_equals_device_frames():
    """`run` on host frames ships only the crop band; its embeddings equal those of device-resident whole frames."""
    spec = VideoMAESpec.tiny()
    enc = VideoMAEClipEncoder(synth_videomae_state_dict(spec, seed=2), spec, DEV, clips_per_batch=2)
    n = 3 * spec.frames + 1
    frames = torch.from_numpy(synth.scene_frames_np(7, n, 270, 480))
    ts = [i / 2.0 for i in range(n)]
    a = enc.run(frames.pin_memory(), ts)
    b = enc.run(frames.to(DEV), ts)
    assert a["num_embeddings"] == b["num_embeddings"] == 4
    for x, y in zip(a["embeddings"], b["embeddings"]):
        assert torch.equal(x["embedding"], y["embedding"]) and x["start_time"] == y["start_time"]


def test_patchify_tubelet_bit_exact():
    g = torch.Generator().manual_seed(0)
    pv = torch.randn(2 * 4, 3, 48, 64, generator=g).to(torch.bfloat16)
    want = videomae_ref.tubelet_patches(pv.float(), 4, 16, 2).reshape(-1, 3 * 2 * 256).to(torch.bfloat16)
    got = ops.patchify_tubelet(pv.to(DEV), 4, 16, 2).cpu()
    assert got.shape == want.shape and torch.equal(got.view(torch.int16), want.view(torch.int16))


def test_mean_tokens():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3 * 1568, 768, generator=g).to(torch.bfloat16)
    got = ops.mean_tokens(x.to(DEV), 3, 1568).cpu()
    want = x.float().view(3, 1568, 768).mean(1)
    assert (got - want).abs().max() < 1e-5


def test_tiny_encoder_vs_oracle_and_golden(golden_dir):
    gold = np.load(f"{golden_dir}/golden_videomae.npz")
    spec = VideoMAESpec.tiny()
    sd = synth_videomae_state_dict(spec, seed=2)
    frames = synth.noise_frames(2 * spec.frames, 123, 211, seed=9)
    seams = {}
    pv = torch.from_numpy(videomae_ref.pixel_values(frames.numpy(), spec.image, spec.image))
    want = videomae_ref.encoder_forward(sd, pv, spec.frames, spec.heads, spec.patch, spec.tubelet, spec.eps, seams=seams)
    enc = VideoMAEClipEncoder(sd, spec, DEV)
    dev_pv = enc.preprocess(frames.to(DEV))
    patches = ops.patchify_tubelet(dev_pv, spec.frames, spec.patch, spec.tubelet)
    pooled, tokens = ops.videomae_forward(enc.pack, patches, return_tokens=True)
    torch.cuda.synchronize()
    ctok = _cos(tokens.float().view(2, spec.tokens, -1), seams["last_hidden_state"])
    print(f"tiny videomae: token cos min {ctok.min():.6f} pooled cos {_cos(pooled, want).tolist()}")
    assert ctok.min() > 0.999 and _cos(pooled, want).min() > 0.999
    assert _cos(pooled, torch.from_numpy(gold["tiny_pooled"])).min() > 0.999
    assert (pooled.cpu() - torch.from_numpy(gold["tiny_pooled"])).abs().max() < 0.05


def test_tiny_encoder_without_final_layernorm_vs_golden(golden_dir):
    """`use_mean_pooling=True` checkpoints (fine-tuned VideoMAE): VideoMAEModel has no final LayerNorm; the base
    checkpoint the reference loads (use_mean_pooling=False) has one — both variants are pinned on HF."""
    import dataclasses
    gold = np.load(f"{golden_dir}/golden_videomae.npz")
    spec = dataclasses.replace(VideoMAESpec.tiny(), final_norm=False)
    sd = synth_videomae_state_dict(spec, seed=2)
    assert "layernorm.weight" not in sd
    enc = VideoMAEClipEncoder(sd, spec, DEV)
    frames = synth.noise_frames(2 * spec.frames, 123, 211, seed=9)
    patches = ops.patchify_tubelet(enc.preprocess(frames.to(DEV)), spec.frames, spec.patch, spec.tubelet)
    pooled, tokens = ops.videomae_forward(enc.pack, patches, return_tokens=True)
    torch.cuda.synchronize()
    want_tok, want = torch.from_numpy(gold["tiny_nofinal_last_hidden_state"]), torch.from_numpy(gold["tiny_nofinal_pooled"])
    assert _cos(tokens.float().view(2, spec.tokens, -1), want_tok).min() > 0.999
    assert _cos(pooled, want).min() > 0.999 and (pooled.cpu() - want).abs().max() < 0.05


def test_base_encoder_vs_hf_golden(golden_dir):
    """VideoMAE-base geometry, one 16-frame 1080p clip (two scenes), synthetic weights seed 2, vs HF fp32."""
    gold = np.load(f"{golden_dir}/golden_videomae.npz")
    spec = VideoMAESpec.base()
    sd = synth_videomae_state_dict(spec, seed=2)
    enc = VideoMAEClipEncoder(sd, spec, DEV)
    clip = torch.cat([synth.scene_frames(0, 8, device=DEV), synth.scene_frames(30, 8, device=DEV)], 0)
    pooled = enc.encode_clips(clip)
    pp = ProjectorPack(synth_projector_state_dict(spec.hidden, 4096, seed=3), DEV)
    proj = ops.project(pp, pooled.to(torch.bfloat16))
    torch.cuda.synchronize()
    want, want_proj = torch.from_numpy(gold["base_pooled"]), torch.from_numpy(gold["base_projected"])
    e = (pooled.cpu() - want).abs().max().item()
    print(f"videomae-base: pooled cos {_cos(pooled, want).tolist()} max_abs {e:.4f}; projected cos "
          f"{_cos(proj, want_proj).tolist()}")
    assert _cos(pooled, want).min() >= 0.999 and _cos(proj, want_proj).min() >= 0.999
    assert e < 0.1  # mean over 1568 tokens of |x| ~ 1 states: bf16 noise averages out, stated bound


def test_run_mirrors_reference_clip_bookkeeping():
    """`run_videomae_encoder` semantics (scripts/extract_features.py:355-390): non-overlapping clips, tail padded
    with its last frame, start/end timestamps, one fp32 CPU embedding per clip."""
    spec = VideoMAESpec.tiny()
    sd = synth_videomae_state_dict(spec, seed=2)
    enc = VideoMAEClipEncoder(sd, spec, DEV, clips_per_batch=2)
    n = 3 * spec.frames + 1  # 13 frames -> 4 clips, the last one padded from a single frame
    frames = synth.noise_frames(n, 90, 120, seed=4)
    ts = [i / 2.0 for i in range(n)]
    out = enc.run(frames, ts, projector=ProjectorPack(synth_projector_state_dict(spec.hidden, 256, seed=3), DEV))
    assert out["num_input_frames"] == n and out["num_embeddings"] == 4 and out["embedding_dim"] == spec.hidden
    last = out["embeddings"][-1]
    assert last["start_time"] == ts[12] and last["end_time"] == ts[12] and last["source_frame_count"] == spec.frames
    assert out["embeddings"][0]["start_time"] == 0.0 and out["embeddings"][0]["end_time"] == ts[3]
    assert last["embedding"].dtype == torch.float32 and last["embedding"].device.type == "cpu"
    assert last["projected"].shape == (256,)
    # the padded clip equals a clip made of 4 copies of the last frame
    rep = enc.encode_clips(frames[-1:].expand(spec.frames, -1, -1, -1).contiguous().to(DEV)).cpu()[0]
    assert torch.allclose(rep, last["embedding"], atol=1e-6)


def test_run_realtime_mirrors_stride8_windows():
    """`extract_videomae_embeddings` semantics (scripts/realtime_inference.py:352-375): windows of `frames` with stride
    frames/2, the tail that does not fill a window dropped, timestamp of the window's middle frame; every window's
    embedding equals encoding that window alone."""
    spec = VideoMAESpec.tiny()  # 4-frame clips -> stride 2
    enc = VideoMAEClipEncoder(synth_videomae_state_dict(spec, seed=2), spec, DEV, clips_per_batch=3)
    n = 11  # windows start at 0, 2, 4, 6 (8..11 would need frame 11: dropped)
    frames = synth.noise_frames(n, 90, 120, seed=5)
    ts = [i / 4.0 for i in range(n)]
    out = enc.run_realtime(frames, ts)
    assert [o["timestamp"] for o in out] == [ts[2], ts[4], ts[6], ts[8]]
    for o, start in zip(out, (0, 2, 4, 6)):
        alone = enc.encode_clips(frames[start:start + spec.frames].to(DEV)).cpu()[0]
        assert o["embedding"].dtype == torch.float32 and o["embedding"].device.type == "cpu"
        assert torch.equal(alone, o["embedding"])
    assert enc.run_realtime(frames[:3], ts[:3]) == []


def test_run_videomae_encoder_caller_dropin_from_pil_frames():
    """`pipeline.run_videomae_encoder([(ts, PIL frame)], device, encoder)` — the reference's caller-level signature
    (scripts/extract_features.py:335-403) — returns exactly what `VideoMAEClipEncoder.run` returns for the same frames
    as a tensor (PIL frames read through Pillow's Arrow export, several batches of clips, padded tail)."""
    from PIL import Image

    from gameplay_vision_llm_b200.pipeline import run_videomae_encoder
    spec = VideoMAESpec.tiny()
    enc = VideoMAEClipEncoder(synth_videomae_state_dict(spec, seed=2), spec, DEV, clips_per_batch=2)
    n = 5 * spec.frames + 3  # 23 frames -> 6 clips over 3 batches, the last clip padded
    frames = synth.noise_frames(n, 90, 120, seed=6)
    ts = [i / 3.0 for i in range(n)]
    want = enc.run(frames, ts)
    got = run_videomae_encoder([(t, Image.fromarray(f.numpy())) for t, f in zip(ts, frames)], DEV, encoder=enc)
    keys = ("num_input_frames", "num_embeddings", "embedding_dim")
    assert {k: got[k] for k in keys} == {k: want[k] for k in keys} and got["num_embeddings"] == 6
    for g, w in zip(got["embeddings"], want["embeddings"]):
        assert (g["start_time"], g["end_time"], g["source_frame_count"]) == (w["start_time"], w["end_time"], w["source_frame_count"])
        assert torch.equal(g["embedding"], w["embedding"])
    # failures degrade to the reference's empty result instead of raising (:399-403)
    assert run_videomae_encoder([(0.0, Image.fromarray(frames[0].numpy()))], DEV, encoder=None)["num_embeddings"] == 0
