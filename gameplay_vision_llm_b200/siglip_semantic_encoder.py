"""Drop-in for the hot-path half of the reference's `src/perception/siglip_semantic_encoder.py`.

Same names, argument meaning and error behaviour as the reference for everything on the path
(SURVEY.md §8b): `NaFlexConfig` (:59-83), `SemanticEmbedding` (:37-56), `SigLIPEncoder._load_model`
(:178-210) with its `_model` / `_processor` seam, `SigLIPSemanticEncoder.encode_image` (:445-483),
`compute_similarity` / `find_similar_regions` (:604-638).  The arithmetic runs in libgvl_sm100a.so:

* `_processor(images=[...], return_tensors="pt")`  -> preprocess kernel (bit-exact with HF's CPU path)
* `_model.get_image_features(pixel_values=...)`    -> patchify + tcgen05 tower + MAP head
* `encode_frames(uint8 [B,H,W,3])`                 -> the batched fast entry the reference lacks

* `encode_masked_regions(frame, masks)`             -> the masked-region variant (reference :485-562):
  Pillow-bicubic crops of the resident frame (bit-exact), HF `interpolate_pos_encoding` for non-square grids, the
  reference's pooling and REN `projection` head — all on the device

Differences, on purpose: a model that cannot be loaded raises (the reference silently degrades to a
random `Placeholder`, :206-210 — a CPU/placeholder fallback would void every parity claim); non-square regions are
encoded with HF's own position-table interpolation instead of raising inside HF (`NaFlexConfig.interpolate_pos_encoding
= False` restores the reference's failure for them).
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops
from .weights import SiglipPack, SiglipVisionSpec, synth_siglip_state_dict

logger = logging.getLogger(__name__)


@dataclass
class SemanticEmbedding:
    """Semantic embedding of a region / frame (reference :37-56)."""

    embedding: torch.Tensor  # (embedding_dim,)
    entity_id: Optional[str] = None
    confidence: float = 1.0
    original_bbox: Optional[tuple[float, float, float, float]] = None
    aspect_ratio: Optional[float] = None

    def __repr__(self) -> str:
        return f"SemanticEmbedding(dim={self.embedding.shape[-1]}, entity={self.entity_id}, conf={self.confidence:.2f})"


@dataclass
class NaFlexConfig:
    """Reference fields (:59-83) first; the fields after `use_tf32` are this build's additions."""

    model_name: str = "google/siglip2-so400m-patch14-384"
    device: str = "cuda"
    dtype: torch.dtype = torch.bfloat16
    base_resolution: int = 384
    min_resolution: int = 128
    max_resolution: int = 768
    preserve_aspect_ratio: bool = True
    embedding_dim: int = 1152
    use_cls_token: bool = True
    pool_strategy: str = "mean"
    batch_size: int = 16
    use_amp: bool = True
    use_tf32: bool = True
    # --- additions ---
    resample: int = 2  # PIL BILINEAR; the checkpoint's preprocessor_config.json value (unknowable offline)
    image_mean: tuple = (0.5, 0.5, 0.5)
    image_std: tuple = (0.5, 0.5, 0.5)
    synthetic_seed: Optional[int] = None  # random-init weights in the HF layout (no network here)
    state_dict: Optional[dict] = None  # an already loaded HF state_dict (vision_model.* keys)
    num_attention_heads: Optional[int] = None  # only needed for a bare state dict of an unpublished tower width
    # masked-region route: HF's `interpolate_pos_encoding=True` for patch grids other than the checkpoint's.  False =
    # the reference as published: HF raises on the position add for every non-square region
    interpolate_pos_encoding: bool = True
    # fold every token-level LayerNorm into the GEMM that consumes it (weights.SiglipPack(fold_ln=True), what
    # EmbeddingPipeline runs): the LayerNorm kernels disappear.  False keeps the reference's op order.
    fold_layernorm: bool = False


class BatchFeature(dict):
    """Minimal stand-in for transformers.BatchFeature: a dict with `.to(device)` and attribute access."""

    def to(self, device=None, *args, **kwargs):
        return BatchFeature({k: (v.to(device, *args, **kwargs) if torch.is_tensor(v) else v) for k, v in self.items()})

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc


def _to_uint8_hwc(image) -> np.ndarray:
    """PIL image / numpy array / tensor -> uint8 [H,W,3] (`convert_rgb` + `pil_to_tensor` of the HF processor)."""
    if hasattr(image, "convert"):  # PIL; an image that already is RGB is read in place (convert() would copy it first)
        return np.asarray(image if getattr(image, "mode", None) == "RGB" else image.convert("RGB"), dtype=np.uint8)
    arr = image.detach().cpu().numpy() if torch.is_tensor(image) else np.asarray(image)
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[-1] != 3:
        raise ValueError("images must be PIL RGB images or uint8 [H,W,3] arrays")
    return arr


def _rgbx_view(image):
    """Zero-copy uint8 [H, W, 4] view of a PIL "RGB" image's own pixel storage (Pillow keeps RGB as 4 bytes per pixel and
    exports it through the Arrow C data interface since 11.2), or None when that is not available.  `np.asarray(image)`
    instead packs the pixels into a bytes object under the GIL (1.7 ms per 1080p frame, not parallelisable)."""
    if getattr(image, "mode", None) != "RGB" or not hasattr(image, "__arrow_c_array__"):
        return None
    try:
        import pyarrow as pa
        flat = pa.array(image).flatten().to_numpy(zero_copy_only=True)
    except Exception:
        return None
    w, h = image.size
    return flat.reshape(h, w, 4) if flat.size == h * w * 4 else None


def _image_shape(image) -> tuple[int, int, int]:
    """(H, W, 3) of a PIL image / array without converting it."""
    if hasattr(image, "convert"):
        w, h = image.size
        return (h, w, 3)
    return tuple(image.shape)


class GvlSiglipProcessor:
    """`AutoProcessor` seam: `processor(images=[...], return_tensors="pt") -> {"pixel_values": fp32 [B,3,S,S]}`.

    The result lives on the GPU already (the reference moves it there right after, :477) and is bit-identical
    to HF SiglipImageProcessor's CPU output for the same `resample` / mean / std."""

    def __init__(self, size: int, resample: int, image_mean, image_std, device: torch.device):
        self.size = {"height": size, "width": size}
        self.resample, self.image_mean, self.image_std = resample, tuple(image_mean), tuple(image_std)
        self.device = device
        self._stage: Optional[tuple] = None  # (pinned uint8 [H,W,4], event of the copy that last read it)

    def _upload_rgbx(self, view: np.ndarray) -> torch.Tensor:
        """One PIL "RGB" image (its zero-copy RGBX storage) -> uint8 [1,H,W,3] on the device: a flat memcpy into a
        reused pinned buffer, an asynchronous copy, the padding byte dropped on the device — instead of Pillow's
        `tobytes` packing (1.7 ms per 1080p frame) followed by a pageable copy."""
        if self._stage is None or tuple(self._stage[0].shape) != view.shape:
            self._stage = (torch.empty(view.shape, dtype=torch.uint8).pin_memory(), None)
        host, event = self._stage
        if event is not None:
            event.synchronize()
        np.copyto(host.numpy(), view)
        with torch.cuda.device(self.device):
            dev = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        self._stage = (host, ev)
        return dev[None, ..., :3].contiguous()

    def __call__(self, images=None, return_tensors: str = "pt", **kwargs) -> BatchFeature:
        if images is None:
            raise ValueError("images is required")
        if not isinstance(images, (list, tuple)):
            images = [images]
        s = self.size["height"]
        if len(images) == 1:  # the per-frame loop's call (`encode_image`, reference :474-477)
            view = _rgbx_view(images[0])
            if view is not None:
                return BatchFeature({"pixel_values": ops.preprocess(self._upload_rgbx(view), s, s, self.resample,
                                                                    self.image_mean, self.image_std,
                                                                    layout=ops.LAYOUT_F32_CHW)})
        arrays = [_to_uint8_hwc(im) for im in images]
        outs = []
        # frames of equal shape go through one launch (HF groups by shape the same way)
        i = 0
        while i < len(arrays):
            j = i
            while j < len(arrays) and arrays[j].shape == arrays[i].shape:
                j += 1
            batch = torch.from_numpy(np.stack(arrays[i:j])).to(self.device, non_blocking=False)
            outs.append(ops.preprocess(batch, s, s, self.resample, self.image_mean, self.image_std,
                                       layout=ops.LAYOUT_F32_CHW))
            i = j
        pv = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
        return BatchFeature({"pixel_values": pv})


class GvlSiglipModel:
    """`AutoModel` seam: `.device`, `.eval()`, `.get_image_features(pixel_values=...) -> Tensor [B, D]` (the
    transformers-4.57 return type the reference's `.squeeze(0)` expects, :483)."""

    def __init__(self, pack: SiglipPack):
        self.pack = pack
        self.spec = pack.spec
        self.device = pack.device
        self.dtype = torch.bfloat16
        self._workspace: Optional[torch.Tensor] = None

    def eval(self):
        return self

    def _ws(self, batch: int) -> torch.Tensor:
        return self._ws_bytes(self.pack.workspace_bytes(batch))

    def _ws_bytes(self, need: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def forward_patches(self, patches: torch.Tensor) -> torch.Tensor:
        return ops.siglip_forward(self.pack, patches, workspace=self._ws(patches.shape[0] // self.spec.tokens))

    def get_image_features(self, pixel_values: torch.Tensor = None, **kwargs) -> torch.Tensor:
        if pixel_values is None:
            raise ValueError("pixel_values is required")
        pv = pixel_values.to(self.device, torch.float32)
        if pv.shape[-1] != self.spec.image or pv.shape[-2] != self.spec.image:
            raise RuntimeError(f"pixel_values must be {self.spec.image}x{self.spec.image} (fixed position embedding)")
        return self.forward_patches(ops.patchify(pv, self.spec.patch, self.spec.patch_ld))


def _load_state_dict(model_name: str) -> dict:
    """A local HF checkpoint directory or file (safetensors / torch).  There is no hub access."""
    path = model_name
    if os.path.isdir(path):
        for cand in ("model.safetensors", "pytorch_model.bin"):
            if os.path.exists(os.path.join(path, cand)):
                path = os.path.join(path, cand)
                break
    if not os.path.isfile(path):
        raise RuntimeError(
            f"cannot load SigLIP weights '{model_name}': not a local checkpoint (no network in this build). Pass a "
            "local path, NaFlexConfig.state_dict, or NaFlexConfig.synthetic_seed for random-init weights.")
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def _load_checkpoint_configs(model_name: str) -> tuple[Optional[dict], Optional[dict]]:
    """(vision config, preprocessor config) from `config.json` / `preprocessor_config.json` beside a local
    checkpoint — what `AutoModel` / `AutoProcessor.from_pretrained` would read (reference :184-204)."""
    import json
    d = model_name if os.path.isdir(model_name) else os.path.dirname(model_name)
    out = []
    for name in ("config.json", "preprocessor_config.json"):
        f = os.path.join(d, name)
        cfg = None
        if d and os.path.isfile(f):
            with open(f) as fh:
                cfg = json.load(fh)
        out.append(cfg)
    vision = out[0]
    if vision is not None and "vision_config" in vision:
        vision = vision["vision_config"]
    return vision, out[1]


class AspectPreservingResizer:
    """NaFlex target size (reference :86-163).  The resize itself runs in `gvl_region_patches_pil_u8`."""

    def __init__(self, config: NaFlexConfig):
        self.config = config

    def compute_optimal_size(self, original_h: int, original_w: int) -> tuple[int, int]:
        """(target_h, target_w): base resolution along the longer side, aspect preserved, bounded by
        min/max_resolution, rounded down to the 14-pixel patch (reference :97-135, same float arithmetic)."""
        aspect_ratio = original_w / original_h
        base = self.config.base_resolution
        if aspect_ratio >= 1:
            target_w = min(self.config.max_resolution, max(self.config.min_resolution, base))
            target_h = max(self.config.min_resolution, int(target_w / aspect_ratio))
        else:
            target_h = min(self.config.max_resolution, max(self.config.min_resolution, base))
            target_w = max(self.config.min_resolution, int(target_h * aspect_ratio))
        patch_size = 14
        target_h = (target_h // patch_size) * patch_size
        target_w = (target_w // patch_size) * patch_size
        return max(patch_size, target_h), max(patch_size, target_w)

    def resize_with_aspect_ratio(self, image):
        """(resized PIL image, aspect ratio) like the reference (:137-163); the resize runs in the region kernel and
        returns the bytes `image.resize((w, h), Image.Resampling.BICUBIC)` returns."""
        from PIL import Image
        arr = _to_uint8_hwc(image)
        h, w = arr.shape[:2]
        th, tw = self.target_size(h, w)
        dev = ops.resolve_device(self.config.device)
        _, u8 = ops.region_patches(torch.from_numpy(np.require(arr, requirements="CW")).to(dev), [(0, 0, w, h)], [(th, tw)],
                                   ((th + 13) // 14 * 14, (tw + 13) // 14 * 14), None, want_patches=False, want_u8=True)
        return Image.fromarray(u8[0, :th, :tw].cpu().numpy()), w / h

    def target_size(self, region_h: int, region_w: int) -> tuple[int, int]:
        """(target_h, target_w) of `resize_with_aspect_ratio` (:137-163), square fallback included."""
        if self.config.preserve_aspect_ratio:
            return self.compute_optimal_size(region_h, region_w)
        return self.config.base_resolution, self.config.base_resolution


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # hard-coded in the reference's prepare_region_tensor (:363-364)
IMAGENET_STD = (0.229, 0.224, 0.225)


class BoxMask:
    """The mask `m = np.zeros((H, W), bool); m[y1:y2, x1:x2] = True` — how scripts/extract_features.py:563-565 turns a
    SAM bounding box into a mask — kept as its four slice bounds.  `region_bbox` reads the extent of the set pixels off
    the bounds (Python slice semantics: negative indices wrap, out-of-range ones clip, like the array assignment)
    instead of scanning H x W bytes that were written a moment earlier; `np.asarray(box_mask)` materialises the array
    for any other consumer."""

    def __init__(self, shape, y1: int, y2: int, x1: int, x2: int):
        self.shape = (int(shape[0]), int(shape[1]))
        self.bounds = (int(y1), int(y2), int(x1), int(x2))

    def extent(self):
        """(x_min, x_max, y_min, y_max) of the set pixels (inclusive), or None when the slices select nothing."""
        y1, y2, x1, x2 = self.bounds
        rows = range(*slice(y1, y2).indices(self.shape[0]))
        cols = range(*slice(x1, x2).indices(self.shape[1]))
        if len(rows) == 0 or len(cols) == 0:
            return None
        return cols[0], cols[-1], rows[0], rows[-1]

    def __array__(self, dtype=None, copy=None):
        y1, y2, x1, x2 = self.bounds
        m = np.zeros(self.shape, dtype=np.bool_)
        m[y1:y2, x1:x2] = True
        return m if dtype is None else m.astype(dtype)


class RegionExtractor:
    """Mask -> expanded bounding box (reference :292-344); the crop / resize / normalisation of :346-367 run on the
    device."""

    def __init__(self, config: NaFlexConfig):
        self.config = config
        self.resizer = AspectPreservingResizer(config)

    @staticmethod
    def region_bbox(frame_shape, mask, expand_ratio: float = 0.1) -> tuple[int, int, int, int]:
        """The box of `extract_masked_region` (:318-338).  Row / column occupancy instead of `np.where` over the whole
        mask: the same min / max, without materialising every set pixel's coordinates."""
        if isinstance(mask, BoxMask):
            ext = mask.extent()
            cols, rows = ((), ()) if ext is None else ((ext[0], ext[1]), (ext[2], ext[3]))
            cols, rows = np.asarray(cols, np.int64), np.asarray(rows, np.int64)
        elif (mask := np.asarray(mask)).dtype == np.bool_ and mask.ndim == 2 and mask.flags.c_contiguous and \
                mask.shape[1] % 8 == 0:
            words = mask.view(np.uint8).view(np.uint64)  # 8 pixels per word: OR-reductions instead of byte scans
            rows = np.flatnonzero(np.bitwise_or.reduce(words, axis=1))
            cols = (np.flatnonzero(np.bitwise_or.reduce(words[rows[0]:rows[-1] + 1], axis=0).view(np.uint8))
                    if rows.size else rows)
        else:
            cols = np.flatnonzero(mask.any(axis=0))
            rows = np.flatnonzero(mask.any(axis=1)) if cols.size else cols
        if cols.size == 0:
            h, w = frame_shape[:2]
            cx, cy = w // 2, h // 2
            size = min(h, w) // 4
            return (cx - size, cy - size, cx + size, cy + size)
        x_min, x_max = cols[0], cols[-1]
        y_min, y_max = rows[0], rows[-1]
        width = x_max - x_min
        height = y_max - y_min
        x_min = max(0, int(x_min - width * expand_ratio))
        y_min = max(0, int(y_min - height * expand_ratio))
        x_max = min(frame_shape[1], int(x_max + width * expand_ratio))
        y_max = min(frame_shape[0], int(y_max + height * expand_ratio))
        return (x_min, y_min, x_max, y_max)

    def extract_masked_region(self, frame, mask, expand_ratio: float = 0.1):
        """(PIL crop, bbox) like the reference (:301-344)."""
        from PIL import Image
        bbox = self.region_bbox(frame.shape, mask, expand_ratio)
        x1, y1, x2, y2 = bbox
        return Image.fromarray(np.asarray(frame)[y1:y2, x1:x2]), bbox

    def prepare_region_tensor(self, region) -> tuple[torch.Tensor, float]:
        """PIL region -> (fp32 CHW tensor on the CPU, aspect ratio) like the reference (:346-367): the resize runs in
        the region kernel (bit-identical to `PIL.Image.resize(BICUBIC)`), the three fp32 operations are the
        reference's."""
        resized, aspect_ratio = self.resizer.resize_with_aspect_ratio(region)
        tensor = torch.from_numpy(np.array(resized)).float().permute(2, 0, 1) / 255.0
        mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
        std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
        return (tensor - mean) / std, aspect_ratio


class RenProjection(torch.nn.Sequential):
    """The REN-style `projection` head (reference :415-420): Linear(D, D) -> GELU (erf) -> Linear(D, D), parameters
    under "0.*" / "2.*" like the reference's nn.Sequential; `forward` runs the two tcgen05 GEMMs of `gvl_project`."""

    def __init__(self, dim: int):
        super().__init__(torch.nn.Linear(dim, dim), torch.nn.GELU(), torch.nn.Linear(dim, dim))
        self.dim = dim
        self._pack = None
        self._pack_key = None

    def _device_pack(self, device) -> "ops.ProjectorPack":
        params = [self[0].weight, self[0].bias, self[2].weight, self[2].bias]
        key = tuple((p.data_ptr(), p._version, p.device, p.dtype) for p in params) + (str(device),)
        if self._pack is None or key != self._pack_key:
            sd = {"net.0.weight": params[0], "net.0.bias": params[1], "net.2.weight": params[2], "net.2.bias": params[3]}
            self._pack = ops.ProjectorPack(sd, device)
            self._pack_key = key
        return self._pack

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("RenProjection.forward needs a CUDA (sm_100a) tensor; there is no CPU fallback")
        x2 = x.reshape(-1, self.dim).to(torch.bfloat16).contiguous()
        with torch.no_grad():
            out = ops.project(self._device_pack(x.device), x2, out_dtype=torch.float32)
        return out.reshape(*x.shape[:-1], self.dim)


class SigLIPEncoder:
    """Lazy loader with the reference's attribute seam (`_model`, `_processor`, `_load_model`)."""

    def __init__(self, config: NaFlexConfig):
        self.config = config
        self._model: Optional[GvlSiglipModel] = None
        self._processor: Optional[GvlSiglipProcessor] = None

    def _load_model(self) -> None:
        if self._model is not None:
            return
        cfg = self.config
        device = ops.resolve_device(cfg.device)
        vision_cfg = pre_cfg = None
        if cfg.state_dict is not None:
            sd = cfg.state_dict
        elif cfg.synthetic_seed is not None:
            sd = None
        else:
            sd = _load_state_dict(cfg.model_name)
            vision_cfg, pre_cfg = _load_checkpoint_configs(cfg.model_name)
        heads = cfg.num_attention_heads or (vision_cfg or {}).get("num_attention_heads")
        spec = SiglipVisionSpec.so400m() if sd is None else spec_from_state_dict(
            sd, cfg.base_resolution, heads=heads, eps=(vision_cfg or {}).get("layer_norm_eps"))
        if sd is None:
            sd = synth_siglip_state_dict(spec, cfg.synthetic_seed)
        resample, mean, std = cfg.resample, cfg.image_mean, cfg.image_std
        if pre_cfg is not None:  # the checkpoint's own preprocessing constants win over the dataclass defaults
            resample = int(pre_cfg.get("resample", resample))
            mean = tuple(pre_cfg.get("image_mean", mean))
            std = tuple(pre_cfg.get("image_std", std))
            self.config.resample, self.config.image_mean, self.config.image_std = resample, mean, std
        logger.info("Loading SigLIP encoder: %s (%d layers, hidden %d)", cfg.model_name, spec.layers, spec.hidden)
        with torch.cuda.device(device):
            self._model = GvlSiglipModel(SiglipPack(sd, spec, device, fold_ln=bool(cfg.fold_layernorm)))
        self._processor = GvlSiglipProcessor(spec.image, resample, mean, std, device)

    def forward(self, pixel_values: torch.Tensor):
        """(sequence_output, pooled_output) like the reference's `SigLIPEncoder.forward` (:246-289).  pixel_values
        [B,3,h,w]; h x w other than the checkpoint's square goes through the re-sampled position table (HF
        `interpolate_pos_encoding=True`) unless `config.interpolate_pos_encoding` is off."""
        self._load_model()
        m = self._model
        B, _, h, w = pixel_values.shape
        gh, gw = h // m.spec.patch, w // m.spec.patch
        if (gh, gw) != (m.spec.grid, m.spec.grid) and not self.config.interpolate_pos_encoding:
            raise RuntimeError(f"pixel_values {h}x{w} do not match the checkpoint's {m.spec.tokens} positions")
        view = ops.siglip_grid_view(m.pack, gh, gw)
        patches = ops.patchify(pixel_values.to(m.device, torch.float32), m.spec.patch, m.spec.patch_ld)
        pooled, tokens = ops.siglip_forward(view, patches, workspace=m._ws_bytes(view.workspace_bytes(B)), return_tokens=True)
        return tokens.view(B, view.tokens, m.spec.hidden), pooled

    __call__ = forward


# attention heads of the published SigLIP / SigLIP2 vision towers by width (a state dict does not carry the count)
_KNOWN_HEADS = {768: 12, 1024: 16, 1152: 16}


def spec_from_state_dict(sd: dict, image: int = 384, heads: Optional[int] = None,
                         eps: Optional[float] = None) -> SiglipVisionSpec:
    """Tower geometry from the HF tensor shapes.  The head count is not recoverable from shapes: it comes from the
    checkpoint's config.json (`heads`), else from the table of published towers, else from the one supported head
    dim (64 / 72) that divides the width; a width both divide (e.g. 576) raises — guessing would silently change the
    arithmetic."""
    pre = "vision_model." if any(k.startswith("vision_model.") for k in sd) else ""
    w = sd[pre + "embeddings.patch_embedding.weight"]
    hidden, patch = w.shape[0], w.shape[-1]
    tokens = sd[pre + "embeddings.position_embedding.weight"].shape[0]
    grid = int(round(tokens ** 0.5))
    layers = 1 + max(int(k[len(pre):].split(".")[2]) for k in sd if k.startswith(pre + "encoder.layers."))
    inter = sd[pre + "encoder.layers.0.mlp.fc1.weight"].shape[0]
    if heads is None:
        heads = _KNOWN_HEADS.get(hidden)
    if heads is None:
        # the attention kernel supports head dims 64 and 72 only: a width only one of them divides is unambiguous
        fits = [hd for hd in (64, 72) if hidden % hd == 0]
        if len(fits) == 1:
            heads = hidden // fits[0]
    if heads is None:
        raise RuntimeError(f"cannot infer the attention head count of a {hidden}-wide tower from its state dict: pass "
                           "NaFlexConfig.num_attention_heads or keep config.json beside the checkpoint")
    if hidden % heads or hidden // heads not in (64, 72):
        raise RuntimeError(f"head dim {hidden / heads:g} is not supported by the attention kernel (64 and 72 are)")
    img = image if image // patch == grid else grid * patch
    kw = {} if eps is None else {"eps": float(eps)}
    return SiglipVisionSpec(hidden=hidden, intermediate=inter, layers=layers, heads=int(heads), image=img, patch=patch,
                            **kw)


class SigLIPSemanticEncoder:
    """Main interface (reference :370-638), hot-path subset."""

    def __init__(self, config: Optional[NaFlexConfig] = None, device: Optional[str] = None):
        self.config = config or NaFlexConfig()
        if device:
            self.config.device = device
        self.encoder = SigLIPEncoder(self.config)
        self.region_extractor = RegionExtractor(self.config)
        # same constructor order and default init as the reference (:412-420), so a seeded construction gives the
        # reference's projection weights; they stay on the host as the parameter container (bf16 device copy on use)
        self.projection = RenProjection(self.config.embedding_dim)
        self._region_lut: Optional[torch.Tensor] = None
        logger.info("SigLIPSemanticEncoder initialized with device=%s", self.config.device)

    # ---- the reference's per-frame entry ----------------------------------------------------------
    def encode_image(self, image) -> torch.Tensor:
        """PIL image -> embedding (embedding_dim,) bf16 on the device; same call sequence as the reference
        (:462-483): processor -> `.to(model.device)` -> `get_image_features(**inputs)` under no_grad."""
        self.encoder._load_model()
        inputs = self.encoder._processor(images=[image], return_tensors="pt").to(self.encoder._model.device)
        with torch.no_grad():
            embedding = self.encoder._model.get_image_features(**inputs)
        return embedding.squeeze(0)

    # ---- batched fast entry -----------------------------------------------------------------------
    def encode_frames(self, frames) -> torch.Tensor:
        """uint8 frames [B,H,W,3] (tensor on any device, or numpy) -> embeddings bf16 [B, D] on the device.
        One fused resize/normalize/patchify launch + one tower pass per `config.batch_size` frames."""
        self.encoder._load_model()
        m = self.encoder._model
        t = torch.as_tensor(frames)
        if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
            raise ValueError("frames must be uint8 [B,H,W,3]")
        t = t.to(m.device).contiguous()
        outs = []
        bs = max(1, int(self.config.batch_size))
        for i in range(0, t.shape[0], bs):
            patches = ops.preprocess(t[i:i + bs], m.spec.image, m.spec.image, self.config.resample, self.config.image_mean,
                                     self.config.image_std, layout=ops.LAYOUT_BF16_PATCH, patch=m.spec.patch,
                                     ld=m.spec.patch_ld)
            outs.append(m.forward_patches(patches))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def encode_images(self, images) -> torch.Tensor:
        """A list of PIL images / uint8 [H,W,3] arrays of ONE size -> embeddings bf16 [N, D] on the device: what
        `[encode_image(im) for im in images]` returns, for callers that keep the reference's data model (a list of PIL
        frames, scripts/extract_features.py:230-264).  The PIL -> array conversions (1.7 ms per 1080p frame, the largest
        host cost of the per-frame loop) run on a small thread pool straight into a pinned double buffer; each batch of
        `config.batch_size` frames crosses PCIe in one asynchronous copy while the previous batch is still in the tower."""
        if len(images) == 0:
            raise ValueError("encode_images: no images")
        self.encoder._load_model()
        m = self.encoder._model
        shape = _image_shape(images[0])
        if any(_image_shape(im) != shape for im in images):
            raise ValueError("encode_images: all images must have one size (group them by size first)")
        bs = max(1, int(self.config.batch_size))
        if len(images) < 4:  # a handful of images (odd-sized crops): not worth a pinned ring of their size
            return self.encode_frames(np.stack([_to_uint8_hwc(im) for im in images]))
        # PIL "RGB" images hand out their RGBX storage without a copy: the pool then only does flat 8 MB memcpys (GIL
        # released) into the pinned buffer, and the padding byte is dropped on the device
        views = [_rgbx_view(im) for im in images]
        rgbx = all(v is not None for v in views)
        ring = self._pinned_ring(bs, shape[:2] + (4,) if rgbx else shape)
        pool = self._convert_pool()
        outs = []
        with torch.cuda.device(m.device):
            stream = torch.cuda.current_stream(m.device)
            for k, i0 in enumerate(range(0, len(images), bs)):
                chunk = views[i0:i0 + bs] if rgbx else images[i0:i0 + bs]
                host, event = ring[k % 2]
                if event is not None:
                    event.synchronize()  # the copy that last read this buffer has finished
                view = host.numpy()
                if rgbx:
                    list(pool.map(lambda a: np.copyto(view[a[0]], a[1]), enumerate(chunk)))
                else:
                    list(pool.map(lambda a: np.copyto(view[a[0]], _to_uint8_hwc(a[1])), enumerate(chunk)))
                dev = host[:len(chunk)].to(m.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
                ring[k % 2] = (host, ev)
                if rgbx:
                    dev = dev[..., :3].contiguous()
                patches = ops.preprocess(dev, m.spec.image, m.spec.image, self.config.resample, self.config.image_mean,
                                         self.config.image_std, layout=ops.LAYOUT_BF16_PATCH, patch=m.spec.patch,
                                         ld=m.spec.patch_ld)
                outs.append(m.forward_patches(patches))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def _pinned_ring(self, bs: int, shape) -> list:
        key = (bs,) + tuple(shape)
        if getattr(self, "_ring_key", None) != key:
            self._ring = [(torch.empty((bs,) + tuple(shape), dtype=torch.uint8).pin_memory(), None) for _ in range(2)]
            self._ring_key = key
        return self._ring

    def _convert_pool(self):
        if getattr(self, "_pool", None) is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1), thread_name_prefix="gvl-pil")
        return self._pool

    # ---- masked-region variant (reference :485-602) ---------------------------------
    def _pool_features(self, tokens: torch.Tensor, pooled: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """`_pool_features` (:426-443): tokens bf16 [B*T, D] (post-layernorm), pooled bf16 [B, D] (MAP head) -> fp32."""
        if self.config.pool_strategy == "mean":
            return ops.mean_tokens(tokens, B, T, torch.float32)
        if self.config.pool_strategy == "max":
            return ops.max_tokens(tokens, B, T, torch.float32)
        return pooled.float()

    def encode_masked_regions(self, frame, masks) -> list[SemanticEmbedding]:
        """frame: RGB uint8 (H, W, 3) (numpy, or a tensor already on the device); masks: [(entity_id, bool (H, W))].
        One H2D copy of the frame, then per `config.batch_size` regions: one crop/resize/normalise/pad/im2col launch,
        one tower pass on the batch canvas, pooling, the projection head, one D2H copy — same batching, padding and
        outputs (fp32 CPU embeddings) as the reference (:503-562)."""
        if not masks:
            return []
        frame_dev, boxes, sizes, metadata = self._prepare_regions(frame, masks)
        bs = max(1, int(self.config.batch_size))
        chunks = [list(range(i, min(i + bs, len(boxes)))) for i in range(0, len(boxes), bs)]
        embeddings = self._encode_region_chunks(frame_dev, boxes, sizes, metadata, chunks)
        logger.debug("Encoded %d masked regions", len(embeddings))
        return embeddings

    def encode_regions_individually(self, frame, masks, max_tokens: int = 32768) -> list[SemanticEmbedding]:
        """What `[encode_masked_regions(frame, [m])[0] for m in masks]` returns — the way scripts/extract_features.py
        :552-583 calls the encoder, one detection per call, so no region is ever zero-padded to a neighbour's size —
        computed as ONE ragged tower pass (`gvl_siglip_forward_ragged`): the token rows of all regions are concatenated,
        every GEMM / LayerNorm runs once over all of them, and only the position add, the attention and the pooling run
        per distinct target size.  Every kernel's per-row arithmetic is independent of the rows around it, so the
        embeddings are bit-identical to the one-per-call results.  `max_tokens` bounds one pass (activation memory)."""
        if not masks:
            return []
        frame_dev, boxes, sizes, metadata = self._prepare_regions(frame, masks)
        m = self.encoder._model
        spec = m.spec
        by_size: dict[tuple[int, int], list[int]] = {}
        for i, sz in enumerate(sizes):
            by_size.setdefault(sz, []).append(i)
        # passes of at most max_tokens rows; a size group may be split across passes
        passes, cur, cur_tokens = [], {}, 0
        for (th, tw), idx in by_size.items():
            gh, gw = th // spec.patch, tw // spec.patch
            if (gh, gw) != (spec.grid, spec.grid) and not self.config.interpolate_pos_encoding:
                raise RuntimeError(f"a {gh}x{gw} patch grid does not match the checkpoint's {spec.tokens} positions "
                                   "(the reference fails here; NaFlexConfig.interpolate_pos_encoding=True encodes it)")
            for i in idx:
                if cur and cur_tokens + gh * gw > max_tokens:
                    passes.append(list(cur.items()))
                    cur, cur_tokens = {}, 0
                cur.setdefault((th, tw), []).append(i)
                cur_tokens += gh * gw
        if cur:
            passes.append(list(cur.items()))
        results: dict[int, SemanticEmbedding] = {}
        for groups in passes:
            shapes = [(len(idx), th // spec.patch, tw // spec.patch) for (th, tw), idx in groups]
            M = sum(b * gh * gw for b, gh, gw in shapes)
            order = [i for _, idx in groups for i in idx]  # regions of equal size adjacent: the ragged pass's groups
            # one crop / resize / normalise / im2col launch for the whole pass, every region on its own canvas
            patches, _ = ops.region_patches(frame_dev, [boxes[i] for i in order], [sizes[i] for i in order], None,
                                            self._region_lut, patch=spec.patch, ld=spec.patch_ld)
            need = ops.siglip_ragged_workspace_bytes(m.pack, M, sum(x[0] for x in shapes))
            pooled, tokens = ops.siglip_forward_ragged(m.pack, patches, shapes, workspace=m._ws_bytes(need), return_tokens=True)
            if self.config.pool_strategy in ("mean", "max"):
                feats, r0 = [], 0
                for b, gh, gw in shapes:
                    feats.append(self._pool_features(tokens[r0:r0 + b * gh * gw], None, b, gh * gw))
                    r0 += b * gh * gw
                feats = torch.cat(feats, 0)
            else:
                feats = pooled.float()
            out = self.projection(feats).cpu()
            for emb, i in zip(out, order):
                meta = metadata[i]
                results[i] = SemanticEmbedding(embedding=emb, entity_id=meta["entity_id"], confidence=1.0,
                                               original_bbox=meta["bbox"], aspect_ratio=meta["aspect_ratio"])
        return [results[i] for i in range(len(boxes))]

    def _prepare_regions(self, frame, masks):
        self.encoder._load_model()
        m = self.encoder._model
        if m.spec.patch != 14:
            raise RuntimeError("encode_masked_regions: the reference rounds region sizes to 14-pixel patches (:131)")
        view = _rgbx_view(frame) if hasattr(frame, "convert") else None
        if torch.is_tensor(frame):
            frame_dev = frame.to(m.device).contiguous()
        elif view is not None:  # a PIL RGB frame: its own storage -> pinned buffer -> device, no `tobytes` packing
            frame_dev = self.encoder._processor._upload_rgbx(view)[0]
        else:
            frame_dev = torch.from_numpy(np.require(_to_uint8_hwc(frame), requirements="CW")).to(m.device)
        if frame_dev.dtype != torch.uint8 or frame_dev.dim() != 3 or frame_dev.shape[2] != 3:
            raise ValueError("frame must be RGB uint8 (H, W, 3)")
        shape = tuple(frame_dev.shape)
        if self._region_lut is None or self._region_lut.device != m.device:
            self._region_lut = ops.region_lut(IMAGENET_MEAN, IMAGENET_STD).to(m.device)
        boxes, sizes, metadata = [], [], []
        for entity_id, mask in masks:
            bbox = self.region_extractor.region_bbox(shape, mask.cpu().numpy() if torch.is_tensor(mask) else mask)
            x1, y1, x2, y2 = bbox
            if x2 <= x1 or y2 <= y1:
                raise ValueError(f"region '{entity_id}': empty bounding box {bbox}")  # PIL raises here in the reference
            boxes.append(bbox)
            sizes.append(self.region_extractor.resizer.target_size(y2 - y1, x2 - x1))
            metadata.append({"entity_id": entity_id, "bbox": bbox, "aspect_ratio": (x2 - x1) / (y2 - y1)})
        return frame_dev, boxes, sizes, metadata

    def _encode_region_chunks(self, frame_dev, boxes, sizes, metadata, chunks) -> list[SemanticEmbedding]:
        """One launch sequence per chunk (a list of region indices); the chunk's canvas is its largest region."""
        m = self.encoder._model
        spec = m.spec
        embeddings: list[SemanticEmbedding] = []
        for chunk in chunks:
            chunk_sizes = [sizes[i] for i in chunk]
            max_h = max(s[0] for s in chunk_sizes)
            max_w = max(s[1] for s in chunk_sizes)
            gh, gw = max_h // spec.patch, max_w // spec.patch
            if (gh, gw) != (spec.grid, spec.grid) and not self.config.interpolate_pos_encoding:
                raise RuntimeError(f"a {gh}x{gw} patch grid does not match the checkpoint's {spec.tokens} positions "
                                   "(the reference fails here; NaFlexConfig.interpolate_pos_encoding=True encodes it)")
            patches, _ = ops.region_patches(frame_dev, [boxes[i] for i in chunk], chunk_sizes, (max_h, max_w),
                                            self._region_lut, patch=spec.patch, ld=spec.patch_ld)
            view = ops.siglip_grid_view(m.pack, gh, gw)
            B = len(chunk)
            pooled, tokens = ops.siglip_forward(view, patches, workspace=m._ws_bytes(view.workspace_bytes(B)),
                                                return_tokens=True)
            feats = self._pool_features(tokens, pooled, B, view.tokens)
            out = self.projection(feats).cpu()
            for emb, i in zip(out, chunk):
                meta = metadata[i]
                embeddings.append(SemanticEmbedding(embedding=emb, entity_id=meta["entity_id"], confidence=1.0,
                                                    original_bbox=meta["bbox"], aspect_ratio=meta["aspect_ratio"]))
        return embeddings

    def encode_with_context(self, frame, mask, context_radius: int = 50) -> tuple[SemanticEmbedding, SemanticEmbedding]:
        """Region + surrounding-context embeddings (reference :564-602): the context mask is the mask dilated
        `context_radius // 3` times (scipy's default cross structuring element) minus the mask."""
        import scipy.ndimage as ndi
        mask = np.asarray(mask.cpu() if torch.is_tensor(mask) else mask)
        region = self.encode_masked_regions(frame, [("region", mask)])
        dilated = ndi.binary_dilation(mask, iterations=context_radius // 3)
        context = self.encode_masked_regions(frame, [("context", dilated & ~mask)])
        return region[0], context[0]

    # ---- similarity (reference :604-638) ------------------------------------------------------------
    @staticmethod
    def _sim_dtype(*tensors: torch.Tensor) -> torch.dtype:
        """bf16 when every side already is bf16 (what encode_image returns: the cast to fp32 is exact either way),
        fp32 otherwise — the reference computes on `.float()` copies, so wider inputs are never narrowed."""
        return torch.bfloat16 if all(t.dtype == torch.bfloat16 for t in tensors) else torch.float32

    def compute_similarity(self, emb1: SemanticEmbedding, emb2: SemanticEmbedding) -> float:
        dt = self._sim_dtype(emb1.embedding, emb2.embedding)
        e1 = emb1.embedding.to(self._sim_device(emb1.embedding)).to(dt).reshape(1, -1).contiguous()
        e2 = emb2.embedding.to(e1.device).to(dt).reshape(1, -1).contiguous()
        scores, _ = ops.topk_cosine(e2, e1, 1, eps=1e-8)
        return float(scores.item())

    def find_similar_regions(self, query: SemanticEmbedding, candidates: Sequence[SemanticEmbedding],
                             top_k: int = 5) -> list[tuple[SemanticEmbedding, float]]:
        """Cosine similarity of `query` to every candidate, best first; equal scores keep the lower index first
        (the reference's stable `list.sort(reverse=True)`, :637)."""
        if not candidates:
            return []
        k_total = min(top_k, len(candidates))
        if k_total > 64:
            raise RuntimeError("find_similar_regions: top_k > 64 is not supported by the selection kernel")
        dev = self._sim_device(query.embedding)
        dt = self._sim_dtype(query.embedding, *(c.embedding for c in candidates))
        index = torch.stack([c.embedding.to(dev).to(dt).reshape(-1) for c in candidates]).contiguous()
        q = query.embedding.to(dev).to(dt).reshape(1, -1).contiguous()
        scores, idx = ops.topk_cosine(index, q, k_total, eps=1e-8)
        return [(candidates[i], float(s)) for s, i in zip(scores[0].tolist(), idx[0].tolist())]

    def _sim_device(self, t: torch.Tensor) -> torch.device:
        if t.is_cuda:
            return t.device
        return ops.resolve_device(self.config.device)


def create_siglip_encoder(model_name: str = "google/siglip2-so400m-patch14-384", device: str = "cuda",
                          preserve_aspect_ratio: bool = True, **kwargs) -> SigLIPSemanticEncoder:
    """Factory with the reference's signature (:641-662)."""
    return SigLIPSemanticEncoder(NaFlexConfig(model_name=model_name, device=device,
                                              preserve_aspect_ratio=preserve_aspect_ratio, **kwargs))
