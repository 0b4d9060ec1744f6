#!/usr/bin/env python
"""bench.py — frames/s of the SigLIP2-so400m + ProjectorBank embedding path on N B200s (BASELINE.json).

    python bench.py --gpus 1 --steps 57 --warmup 3            # our arm (default)
    python bench.py --impl reference --steps 1 --warmup 1     # the reference's CPU path, same metric
    torchrun ... bench.py --gpus N ...                        # one rank per GPU

A *step* is one pass of the hot path over one batch of 64 synthetic 1080p frames per GPU
(preprocess -> 27-layer tower -> MAP head -> projector -> row of the timeline index).  The default 57
steps are BASELINE.json configs[1]: one hour of 1 fps gameplay (3600 frames) rounded up to whole
batches (57 x 64 = 3648).  With N > 1 each rank owns a contiguous chunk of the timeline with the same
per-rank work ("weak" scaling) and the projected index is all-gathered over NCCL inside the timed
region.  `--scaling strong [--frames 3600]` is BASELINE.json configs[2]: the SAME N-frame timeline split as
[r*ceil(N/W), ...) across the W ranks, uneven tail batch and padded all-gather included (`--steps` is then
derived: ceil(ceil(N/W) / batch) batches per rank).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_H, FRAME_W = 1080, 1920
FRAME_BYTES = FRAME_H * FRAME_W * 3
PRE_BYTES_PER_FRAME = FRAME_BYTES + 729 * 588 * 2  # BASELINE.md §3 (patch layout written directly)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines, self.first = gpu_index, None, [], 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            self.t0 = time.time()
        except OSError:
            self.proc = None

    def live(self) -> bool:
        """nvidia-smi needs ~0.5 s to print its first line.  A caller with a short timed region (the strong-scaled
        3600-frame timeline on 8 GPUs lasts 0.3 s) keeps warming up until this is true, then calls mark()."""
        return self.proc is None or bool(self.lines) or time.time() - self.t0 > 3.0 or self.proc.poll() is not None

    def mark(self) -> None:
        self.first = len(self.lines)  # samples from here on belong to the timed region

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:] or self.lines[-1:]:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference(n_frames: int, batch: int = 8, warmup_batches: int = 1) -> dict:
    """The reference's CPU path for this metric on the host cores (oracle/hf_baseline.py)."""
    from oracle import hf_baseline
    return hf_baseline.run(n_frames=n_frames, batch=batch, warmup_batches=warmup_batches, frame_hw=(FRAME_H, FRAME_W))


MODEL_TEXT = "through SigLIP2-so400m-patch14-384 + ProjectorBank 1152->4096->4096"
WORKLOAD = ("1 h synthetic 1080p gameplay @1 fps (BASELINE.json configs[1]: 3600 frames, rounded up to 57 batches of 64) "
            + MODEL_TEXT)


def workload_text(total_frames: int, world: int, batch: int, steps: int, strong: bool) -> str:
    """States the frames actually timed (the text must follow --steps, not the default)."""
    if strong:
        return (f"BASELINE.json configs[{1 if world == 1 else 2}]: the {total_frames}-frame synthetic 1080p timeline (1 h @1 fps "
                f"at 3600) split into contiguous chunks of ceil({total_frames}/{world}) frames per GPU, batches of {batch} with "
                f"the uneven tail batch, {MODEL_TEXT}")
    if total_frames >= 72000:
        return ("10 h synthetic 1080p gameplay @2 fps (BASELINE.json configs[4]: 72 000 frames, rounded up to "
                f"{total_frames}) {MODEL_TEXT}, then cosine top-16 retrieval over the gathered index")
    if steps * batch == 3648:
        return WORKLOAD if world == 1 else WORKLOAD + f", {world} such chunks (one per GPU)"
    return (f"{total_frames} synthetic 1080p frames = {steps} steps x {batch} frames per GPU x {world} GPU(s) timed — a "
            f"{steps}/57 slice of BASELINE.json configs[1] (1 h @1 fps, 57 batches of 64) {MODEL_TEXT}")


def source_sha256(*rel_paths: str) -> str:
    """Hash of kernel sources: ties a committed ncu capture (profiles/gemm_traffic.json) to the tree that produced it."""
    import hashlib
    h = hashlib.sha256()
    for rel in rel_paths:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


GEMM_SOURCES = ("gameplay_vision_llm_b200/csrc/gemm.cu", "gameplay_vision_llm_b200/csrc/common.cuh")


def main_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(64, max(1, args.steps) * 8)  # bounded sample: at most 8 batches of 8 frames (~2 min on 8 cores)
    res = cpu_reference(n_frames=n, batch=8, warmup_batches=1 if args.warmup > 0 else 0)
    line = {
        "impl": "reference", "metric": "frames/s SigLIP2+ProjectorBank", "value": res["frames_per_s"], "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["seconds"] / (n / 8) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": [FRAME_H, FRAME_W, 3],
                   "sample": "bounded sample of the workload: one step = one batch of 8 frames on the host CPU in fp32, all "
                             "host threads (BASELINE.json configs[0] procedure: HF SiglipImageProcessor + SiglipVisionModel "
                             "+ projector)", "frames": n, "batch": 8},
        "cpu_baseline": {"value": res["frames_per_s"], "unit": "frames/s", "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def main_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from gameplay_vision_llm_b200 import _lib, synth
    from gameplay_vision_llm_b200.pipeline import EmbeddingPipeline, shard_range
    from gameplay_vision_llm_b200.timeline import TimelineEmbeddingIndex
    from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_projector_state_dict,
                                                    synth_siglip_state_dict)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().gvl_check_device(local_rank), "gvl_check_device")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    peaks = load_peaks()
    spec = SiglipVisionSpec.so400m()
    B, W = args.batch, args.warmup
    strong = args.scaling == "strong"
    # the timeline: weak = K batches of B frames per rank (per-GPU work fixed); strong = --frames in total (configs[2])
    total_frames = int(args.frames) if strong else args.steps * B * world
    lo, hi = shard_range(total_frames, rank, world)
    n_local = hi - lo
    per_rank = -(-total_frames // world)
    K = -(-per_rank // B)  # batches per rank (the last one may be short; trailing ranks may own fewer rows)
    pipe = EmbeddingPipeline(synth_siglip_state_dict(spec, seed=0), synth_projector_state_dict(spec.hidden, 4096, seed=1),
                             spec, dev, batch=B, fold_ln=not args.no_fold_ln)

    # this rank's chunk of the timeline, resident in HBM (22.7 GB at 57 x 64 frames)
    frames = torch.empty((max(n_local, 1), FRAME_H, FRAME_W, 3), dtype=torch.uint8, device=dev)
    for i0 in range(0, n_local, 16):
        n = min(16, n_local - i0)
        frames[i0:i0 + n] = synth.scene_frames(lo + i0, n, FRAME_H, FRAME_W, device=dev)
    # the timeline index: every rank contributes per_rank rows to the (padded) gather, rows [0, total) in timestamp order
    timeline = TimelineEmbeddingIndex(total_frames, 4096, fps=1.0, device=dev, rank=rank, world=world)
    index_local = timeline.local_rows()
    spans = [(i0, min(n_local, i0 + B)) for i0 in range(0, n_local, B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_region():
        for i0, i1 in spans:
            pipe.embed(frames[i0:i1], out_index=index_local[i0:i1])
        timeline.all_gather()  # NCCL all_gather_into_tensor over NVLink (no-op on one GPU)

    sampler = ClockSampler(local_rank)
    sampler.start()
    extra = 0
    # warm-up: W batches (+ one all-gather).  --scaling strong only (a timed region of a few hundred ms on 8 GPUs): keep
    # warming up, under load, until the clock sampler prints its first line, so the region is not over before it
    for s in range(W + (200 if strong else 0)):
        if s >= W:
            alive = torch.tensor([1 if sampler.live() else 0], device=dev)
            if world > 1:
                dist.all_reduce(alive, op=dist.ReduceOp.MIN)  # every rank leaves the warm-up at the same step
            if int(alive.item()):
                break
            extra += 1
        i0, i1 = spans[s % len(spans)]
        pipe.embed(frames[i0:i1], out_index=index_local[i0:i1])
    timeline.all_gather()
    barrier()

    # ---- timed region 1: device-resident inputs (value) ----
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    barrier()
    e0.record()
    run_region()
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = total_frames / (ms * 1e-3)

    # ---- the exchange step alone: achieved all-gather bandwidth over NVLink ----
    allgather = None
    if world > 1:
        for _ in range(2):
            timeline.all_gather()
        barrier()
        reps = 10
        e0.record()
        for _ in range(reps):
            timeline.all_gather()
        e1.record()
        barrier()
        ag_ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(ag_ms, op=dist.ReduceOp.MAX)
        ag_ms = float(ag_ms.item())
        ag_bytes = world * per_rank * 4096 * 2
        allgather = {"bytes_total": ag_bytes, "ms": round(ag_ms, 4), "algbw_gbs": round(ag_bytes / (ag_ms * 1e-3) / 1e9, 1),
                     "busbw_gbs": round(ag_bytes * (world - 1) / world / (ag_ms * 1e-3) / 1e9, 1),
                     "nvlink_peak_gbs": 770.0, "note": "busbw = bytes_total x (W-1)/W / time (each GPU receives W-1 shards); "
                     "peak = measured peer copy per direction per GPU (B200_PROFILING.md)"}

    # ---- timed region 2: same steps with per-launch CUDA events (roofline of the dominant kernel) ----
    _lib.prof_enable(True)
    run_region()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_summary()
    gemm = prof.get("gemm", {"ms": 0.0, "launches": 0, "work": 0.0})
    gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else 0.0
    prof_total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    # DRAM bytes per GEMM launch come from a committed ncu capture; it only counts when it was taken on THIS tree's
    # GEMM sources (hash recorded by tools/ncu_traffic.py), otherwise the line says null and why
    traffic, traffic_note = None, "no ncu capture committed (profiles/gemm_traffic.json)"
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("source_sha256") == source_sha256(*GEMM_SOURCES):
            traffic, traffic_note = tj.get("dram_bytes_per_launch"), f"ncu capture {tj.get('capture', '?')} of this tree's gemm.cu"
        else:
            traffic_note = "stale: profiles/gemm_traffic.json was captured on different gemm.cu / common.cuh sources"
    roofline = {"kernel": "gemm_bf16_cg2_kernel", "bound": "tensor", "achieved": round(gemm_tflops, 2),
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": round(gemm_tflops / peaks["bf16_tflops_sustained"], 4), "traffic": traffic,
                "traffic_note": traffic_note, "algorithmic_bytes_per_launch": round(gemm_alg_bytes(spec, B)),
                "peak_source": f"{peaks['source']} (sustained; burst {peaks['bf16_tflops']})",
                "launches": gemm["launches"], "avg_launch_ms": round(gemm["ms"] / max(1, gemm["launches"]), 4),
                "share_of_step": round(gemm["ms"] / prof_total_ms, 4)}
    kernels = {k: {"ms_per_step": round(v["ms"] / K, 4), "launches_per_step": round(v["launches"] / K, 2),
                   "share": round(v["ms"] / prof_total_ms, 4)} for k, v in prof.items()}
    if "preprocess" in prof and prof["preprocess"]["ms"]:
        gbs = prof["preprocess"]["work"] / (prof["preprocess"]["ms"] * 1e-3) / 1e9
        kernels["preprocess"].update({"achieved_gbs": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4)})
    if "attention" in prof and prof["attention"]["ms"]:
        kernels["attention"]["achieved_tflops"] = round(prof["attention"]["work"] / (prof["attention"]["ms"] * 1e-3) / 1e12, 2)
    model_tflops = value / world * spec.flops_per_frame() / 1e12

    # ---- timed region 3: end to end through the public API with HOST frames (e2e) ----
    ring = [frames[i0:i1].cpu().pin_memory() for i0, i1 in spans[:4]]
    host_out = torch.empty((max(n_local, 1), 4096), dtype=torch.bfloat16).pin_memory()

    def host_batches(n_spans):
        for s, (i0, i1) in enumerate(spans[:n_spans]):
            yield ring[s % len(ring)][: i1 - i0]

    pipe.embed_stream(host_batches(min(W, 2)), index_local, host_out)
    barrier()
    e0.record()
    pipe.embed_stream(host_batches(len(spans)), index_local, host_out)
    timeline.all_gather()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = total_frames / (float(ms_e2e.item()) * 1e-3)
    index_full = timeline.index()

    # ---- side measurement (not part of value / e2e): configs[4] retrieval over a 72 000-row timeline index ----
    retrieval = None
    if rank == 0 and world == 1 and not args.no_retrieval:
        del frames, ring
        torch.cuda.empty_cache()
        retrieval = retrieval_probe(dev)
    elif rank == 0 and world > 1 and not args.no_retrieval:
        # multi-GPU runs: retrieval over the index this run has just built and all-gathered (with --steps 141 on 8 GPUs
        # that is BASELINE.json configs[4]: 72 192 frames = 10 h at 2 fps, top-16 for 128 queries)
        retrieval = retrieval_on_gathered(index_full, dev)

    # ---- side measurement: the reference's caller-level function on its own data model (a list of PIL frames) ----
    caller = None
    if rank == 0 and world == 1 and not args.no_caller_api:
        try:
            caller = caller_api_probe(dev, synth_siglip_state_dict(spec, seed=0))
        except Exception as exc:  # Pillow / pyarrow missing on the box: the side measurement is simply absent
            caller = {"unavailable": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference(n_frames=args.cpu_frames, batch=8, warmup_batches=1)
        line = {
            "metric": "frames/s SigLIP2+ProjectorBank", "value": round(value, 2), "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": round(ms / K, 3), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_text(total_frames, world, B, K, strong), "frames_total": total_frames, "warmup_extra_steps": extra,
                       "frames_per_gpu": per_rank, "batch": B, "frame": [FRAME_H, FRAME_W, 3],
                       "weights": "random init, seeds 0/1", "layernorm": "separate kernels" if args.no_fold_ln else
                       "folded into the consuming GEMM epilogues", "sharding": "contiguous timeline chunk per rank, "
                       "NCCL all-gather of the projected index inside the timed region" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (398 MB of frames per step, never reused)"},
            "model_tflops_per_gpu": round(model_tflops, 1),
            "model_frac_of_sustained_peak": round(model_tflops / peaks["bf16_tflops_sustained"], 4),
            "roofline": roofline, "kernels": kernels, "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": B * FRAME_BYTES,
                    "d2h_bytes_per_step": B * 4096 * 2},
            "gpu_launches": int(launches),
        }
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu["frames_per_s"], "unit": "frames/s", "cores": cpu["cores"],
                                    "kind": cpu["kind"], "sample": cpu["sample"]}
        if retrieval is not None:
            line["retrieval"] = retrieval
        if caller is not None:
            line["reference_caller_api"] = caller
        if allgather is not None:
            line["allgather"] = allgather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def gemm_alg_bytes(spec, batch: int) -> float:
    """Average algorithmic bytes of one GEMM launch of a step: A + W + out (+ residual), bf16."""
    M, D, I, T = batch * spec.tokens, spec.hidden, spec.intermediate, spec.tokens
    per_layer = [(M, 3 * D, D, 0), (M, D, D, 1), (M, I, D, 0), (M, D, I, 1)]
    shapes = [(M, D, spec.patch_ld, 0)] + per_layer * spec.layers + [
        (M, 2 * D, D, 0), (batch, D, D, 0), (batch, I, D, 0), (batch, D, I, 1), (batch, 4096, D, 0), (batch, 4096, 4096, 0)]
    total = sum(2 * (m * k + n * k + m * n * (1 + res)) for m, n, k, res in shapes)
    return total / len(shapes)


def retrieval_on_gathered(index_full, dev, n_queries: int = 128, k: int = 16) -> dict:
    """Cosine top-k over the gathered timeline index of this run.  Queries are 128 rows of the index, so each query's
    best hit must be its own row at cosine 1 (or an identical earlier frame): checked."""
    import torch

    from gameplay_vision_llm_b200 import ops

    peaks = load_peaks()
    n, dim = index_full.shape
    g = torch.Generator(device=dev).manual_seed(7)
    rows = torch.randperm(n, device=dev, generator=g)[:n_queries].sort().values
    queries = index_full[rows].contiguous()
    inv = ops.row_inv_norm(index_full)
    scores, idx = ops.topk_cosine(index_full, queries, k, inv_norm=inv)
    torch.cuda.synchronize()
    hit = idx[:, 0].to(torch.int64)
    same = (index_full[hit] == queries).all(dim=1)
    if not bool(((hit <= rows) & same).all()) or not bool((scores[:, 0] > 0.9999).all()):
        raise RuntimeError("retrieval over the gathered index: a query did not retrieve its own row first")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ops.topk_cosine(index_full, queries, k, inv_norm=inv)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = n * dim * 2
    return {"workload": f"cosine top-{k} of {n_queries} queries over the gathered ({n}, {dim}) bf16 timeline index of this run",
            "ms_per_query_batch": round(ms, 4), "queries_per_s": round(n_queries / (ms * 1e-3), 1),
            "index_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1), "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
            "self_retrieval": "every query's first hit is its own row (or an identical earlier frame), cosine > 0.9999",
            "path": "fused tcgen05 scoring with per-CTA candidate lists + float64 re-score of everything within the margin (bit-identical to the scan path)"
                    if n >= 4096 else "fp32 scan (index below the tensor path's 4096-row threshold)"}


def caller_api_probe(dev, siglip_sd, n_frames: int = 128) -> dict:
    """`pipeline.run_siglip_encoder([(timestamp, PIL 1080p frame)], device)` — the signature, data model and output of
    scripts/extract_features.py:502-610 (fallback branch: every frame encoded whole, 1152-d embeddings returned on the CPU
    in the reference's list of dicts; no projector: the reference applies it later).  Wall clock around the call."""
    import torch
    from PIL import Image

    from gameplay_vision_llm_b200 import synth
    from gameplay_vision_llm_b200.pipeline import run_siglip_encoder
    from gameplay_vision_llm_b200.siglip_semantic_encoder import NaFlexConfig, SigLIPSemanticEncoder
    enc = SigLIPSemanticEncoder(NaFlexConfig(device=str(dev), state_dict=siglip_sd, fold_layernorm=True))
    frames = [(float(i), Image.fromarray(f)) for i, f in enumerate(synth.scene_frames_np(0, n_frames, FRAME_H, FRAME_W))]
    run_siglip_encoder(frames[:32], str(dev), encoder=enc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = run_siglip_encoder(frames, str(dev), encoder=enc)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert len(out) == n_frames and out[0]["embedding"].shape[-1] == 1152
    return {"call": "run_siglip_encoder([(timestamp, PIL.Image 1080p)], device) -> [{timestamp, embedding (CPU), ...}]",
            "frames": n_frames, "batch_size": enc.config.batch_size, "frames_per_s": round(n_frames / dt, 1),
            "note": "host wall clock incl. PIL -> pinned memory (Pillow's zero-copy Arrow export), H2D, tower, D2H and the "
                    "reference's list of dicts; the 1152-d SigLIP embedding only, as in the reference function"}


def retrieval_probe(dev, n_index: int = 72000, dim: int = 4096, n_queries: int = 128, k: int = 16) -> dict:
    """BASELINE.json configs[4] retrieval step on one GPU: cosine top-16 of 128 queries over a (72 000, 4096) bf16
    timeline index (10 h at 2 fps) through TimelineEmbeddingIndex.search — the tensor-core scoring path with exact
    re-scoring.  Random clustered rows (the index content does not change the work).  HBM roofline: the index read once."""
    import torch

    from gameplay_vision_llm_b200.timeline import TimelineEmbeddingIndex

    peaks = load_peaks()
    idx = TimelineEmbeddingIndex(n_index, dim, fps=2.0, device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    centers = torch.randn(n_index // 60, dim, device=dev, generator=g)
    rows = idx.local_rows()
    for i0 in range(0, n_index, 6000):
        c = centers[i0 // 60:(i0 + 6000) // 60].repeat_interleave(60, 0)
        rows[i0:i0 + 6000] = (c + 0.35 * torch.randn(c.shape, device=dev, generator=g)).to(torch.bfloat16)
    queries = (centers[torch.randint(0, n_index // 60, (n_queries,), device=dev, generator=g)]
               + 0.35 * torch.randn(n_queries, dim, device=dev, generator=g)).to(torch.bfloat16)
    del centers
    idx.search(queries, top_k=k)  # also fills the cached 1/|e_n|
    # warm-up: a second of searches — the probe runs right after the embedding steps held the GPU at its power cap, and
    # the governor takes a few hundred ms to release the SM clock (a retrieval service does not share that state)
    t0 = time.time()
    while time.time() - t0 < 1.0:
        for _ in range(50):
            idx.search(queries, top_k=k)
        torch.cuda.synchronize()
    sampler = ClockSampler(dev.index if hasattr(dev, "index") and dev.index is not None else 0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2000
    e0.record()
    for _ in range(reps):
        idx.search(queries, top_k=k)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / reps
    nbytes = n_index * dim * 2
    return {"workload": f"cosine top-{k} of {n_queries} queries over a ({n_index}, {dim}) bf16 timeline index",
            "ms_per_query_batch": round(ms, 4), "queries_per_s": round(n_queries / (ms * 1e-3), 1),
            "index_gbs": round(nbytes / (ms * 1e-3) / 1e9, 1), "hbm_frac": round(nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
            "reps": reps, "sm_mhz": clocks.get("sm_mhz"),
            "path": "fused tcgen05 scoring with per-CTA candidate lists + float64 re-score of everything within the margin (bit-identical to the scan path)"}


# ------------------------------------------------------------------------------ VideoMAE workload (configs[3])
def main_videomae(args) -> None:
    """BASELINE.json configs[3]: VideoMAE-base 16-frame tubelet clips through the same ViT kernels + the 768 -> 4096
    projector, batch 32 clips per GPU.  A step = 32 clips (512 synthetic 1080p frames)."""
    import torch
    import torch.distributed as dist

    from gameplay_vision_llm_b200 import _lib, ops, synth
    from gameplay_vision_llm_b200.videomae_encoder import VideoMAEClipEncoder
    from gameplay_vision_llm_b200.weights import (ProjectorPack, VideoMAESpec, synth_projector_state_dict,
                                                    synth_videomae_state_dict)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().gvl_check_device(local_rank), "gvl_check_device")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    spec = VideoMAESpec.base()
    C = 32 if args.batch == 64 else args.batch  # clips per step
    K, W = args.steps, args.warmup
    enc = VideoMAEClipEncoder(synth_videomae_state_dict(spec, seed=2), spec, dev, clips_per_batch=C)
    pp = ProjectorPack(synth_projector_state_dict(spec.hidden, 4096, seed=3), dev)
    nf = C * spec.frames
    # a ring of 4 distinct clip batches stays resident (4 x 3.2 GB); every step reads 3.2 GB of frames, > L2
    ring = []
    for r in range(min(4, max(K, 1))):
        buf = torch.empty((nf, FRAME_H, FRAME_W, 3), dtype=torch.uint8, device=dev)
        for i0 in range(0, nf, 16):
            buf[i0:i0 + 16] = synth.scene_frames((rank * 4 + r) * nf + i0, 16, FRAME_H, FRAME_W, device=dev)
        ring.append(buf)
    n_local = K * C
    index_local = torch.empty((n_local, 4096), dtype=torch.bfloat16, device=dev)
    index_full = torch.empty((world * n_local, 4096), dtype=torch.bfloat16, device=dev) if world > 1 else index_local
    hidden = torch.empty((C, 4096), dtype=torch.bfloat16, device=dev)

    def step(s, frames):
        emb = enc.encode_clips(frames, out_dtype=torch.bfloat16)
        ops.project(pp, emb, hidden=hidden, out=index_local[s * C:(s + 1) * C])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_region(steps):
        for s in range(steps):
            step(s, ring[s % len(ring)])
        if world > 1:
            dist.all_gather_into_tensor(index_full, index_local)

    for s in range(W):
        step(s % K, ring[s % len(ring)])
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    barrier()
    e0.record()
    run_region(K)
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * n_local / (ms * 1e-3)

    _lib.prof_enable(True)
    run_region(K)
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_summary()
    gemm = prof.get("gemm", {"ms": 0.0, "launches": 0, "work": 0.0})
    gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else 0.0
    prof_total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_step": round(v["ms"] / K, 4), "launches_per_step": round(v["launches"] / K, 2),
                   "share": round(v["ms"] / prof_total_ms, 4)} for k, v in prof.items()}
    if "preprocess" in prof and prof["preprocess"]["ms"]:
        gbs = prof["preprocess"]["work"] / (prof["preprocess"]["ms"] * 1e-3) / 1e9
        kernels["preprocess"].update({"achieved_gbs": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4)})
    if "attention" in prof and prof["attention"]["ms"]:
        kernels["attention"]["achieved_tflops"] = round(prof["attention"]["work"] / (prof["attention"]["ms"] * 1e-3) / 1e12, 2)

    # end to end: host frames (pinned) -> H2D of the column band the center crop reads -> clips -> D2H of the projected
    # rows, every step (whole frames stay on the host: the crop of a 16:9 frame never touches ~43 % of each row)
    host_ring = [r.cpu().pin_memory() for r in ring[:2]]
    host_out = torch.empty((n_local, 4096), dtype=torch.bfloat16).pin_memory()
    band_x0, band_w = enc.source_band(FRAME_H, FRAME_W)
    band = (band_x0, FRAME_W)
    stage = [torch.empty((nf, FRAME_H, band_w, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]

    def step_band(s, frames):
        emb = enc.encode_clips(frames, out_dtype=torch.bfloat16, band=band)
        ops.project(pp, emb, hidden=hidden, out=index_local[s * C:(s + 1) * C])

    def e2e_region(steps):
        main = torch.cuda.current_stream()
        for s in range(steps + 1):
            if s < steps:
                with torch.cuda.stream(copy_stream):
                    if s >= 2:
                        copy_stream.wait_event(freed[s % 2])
                    ops.copy_band_h2d(stage[s % 2], host_ring[s % len(host_ring)], band_x0)
                    ready[s % 2].record(copy_stream)
            if s >= 1:
                t = s - 1
                main.wait_event(ready[t % 2])
                step_band(t, stage[t % 2])
                freed[t % 2].record(main)
                host_out[t * C:(t + 1) * C].copy_(index_local[t * C:(t + 1) * C], non_blocking=True)
        if world > 1:
            dist.all_gather_into_tensor(index_full, index_local)

    e2e_region(min(2, K))
    barrier()
    e0.record()
    e2e_region(K)
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * n_local / (float(ms_e2e.item()) * 1e-3)

    if rank == 0:
        model_tflops = value / world * spec.flops_per_clip() / 1e12
        line = {
            "metric": "clips/s VideoMAE-base+projector", "value": round(value, 2), "unit": "clips/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "dtype_note": "bf16 operands with fp32 accumulation; the reference runs VideoMAE in fp32 "
                          "(scripts/extract_features.py:351) — parity is pinned at cosine >= 0.999 against the fp32 HF model",
            "config": {"workload": "BASELINE.json configs[3]: VideoMAE-base 16-frame tubelet clips of synthetic 1080p "
                                   "frames through the ViT kernels + 768->4096 projector", "clips_per_step": C,
                       "frames_per_clip": spec.frames, "frame": [FRAME_H, FRAME_W, 3], "weights": "random init, seeds 2/3",
                       "l2": "inputs larger than L2 (3.2 GB of frames per step)"},
            "model_tflops_per_gpu": round(model_tflops, 1),
            "model_frac_of_sustained_peak": round(model_tflops / peaks["bf16_tflops_sustained"], 4),
            "roofline": {"kernel": "gemm_bf16_cg2_kernel", "bound": "tensor", "achieved": round(gemm_tflops, 2),
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": round(gemm_tflops / peaks["bf16_tflops_sustained"], 4), "traffic": None,
                         "peak_source": f"{peaks['source']} (sustained; burst {peaks['bf16_tflops']})"},
            "kernels": kernels, "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "clips/s", "h2d_bytes_per_step": nf * FRAME_H * band_w * 3,
                    "d2h_bytes_per_step": C * 4096 * 2,
                    "h2d_note": f"source columns [{band_x0}, {band_x0 + band_w}) of every {FRAME_W}-wide row: what the "
                                "processor's center crop reads (cudaMemcpy2DAsync from pinned whole frames)"},
            "gpu_launches": int(launches),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()

# ------------------------------------------------------------------------------ masked-region workload
def main_regions(args) -> None:
    """The masked-region route (scripts/extract_features.py:551-585 -> encode_masked_regions): a step = one synthetic
    1080p frame with 16 seeded SAM-style detections, each encoded exactly as the reference's one-call-per-detection loop
    would (own canvas, no padding to a neighbour), through `SigLIPSemanticEncoder.encode_regions_individually` = one
    ragged tower pass per frame.
    value: the frame already resident on the device, detections as `BoxMask`es (`encode_regions_individually`); e2e: the
    reference's caller API, `pipeline.run_siglip_encoder([(ts, host frame)], sam_results=detections, ...)`, frame copied
    from host memory and the reference's list of dicts built.  Both include the D2H of the fp32 embeddings, which the
    API returns on the CPU like the reference.  Replicas only: no collective on this route."""
    import numpy as np
    import torch

    from gameplay_vision_llm_b200 import _lib, synth
    from gameplay_vision_llm_b200.pipeline import run_siglip_encoder
    from gameplay_vision_llm_b200.siglip_semantic_encoder import BoxMask, NaFlexConfig, SigLIPSemanticEncoder
    from gameplay_vision_llm_b200.weights import (SiglipVisionSpec, synth_ren_projection_state_dict,
                                                    synth_siglip_state_dict)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().gvl_check_device(local_rank), "gvl_check_device")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    spec = SiglipVisionSpec.so400m()
    R = 16
    K, W = args.steps, args.warmup
    enc = SigLIPSemanticEncoder(NaFlexConfig(device=str(dev), state_dict=synth_siglip_state_dict(spec, seed=0), batch_size=R,
                                             fold_layernorm=not args.no_fold_ln))
    enc.projection.load_state_dict(synth_ren_projection_state_dict(spec.hidden, seed=3))
    frames_host = [synth.scene_frames_np((rank * 4 + i) * 30, 1, FRAME_H, FRAME_W)[0] for i in range(4)]
    frames_dev = [torch.from_numpy(f).to(dev) for f in frames_host]
    boxes = synth.region_boxes(R, FRAME_H, FRAME_W)
    masks = [(f"det{i}", BoxMask((FRAME_H, FRAME_W), y1, y2, x1, x2)) for i, (x1, y1, x2, y2) in enumerate(boxes)]
    dets = [{"timestamp": 0.0, "bbox": list(b), "entity_type": "ui", "entity_id": f"det{i}"} for i, b in enumerate(boxes)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_region(frames, steps):
        out = None
        for s in range(steps):
            out = enc.encode_regions_individually(frames[s % len(frames)], masks)
        return out

    def run_caller(frames, steps):  # the reference's caller-level API, one frame per call
        out = None
        for s in range(steps):
            out = run_siglip_encoder([(0.0, frames[s % len(frames)])], str(dev), sam_results=dets, entity_tracker=True,
                                     encoder=enc)
        assert len(out) == R
        return out

    run_region(frames_dev, max(W, 3))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    barrier()
    e0.record()
    out = run_region(frames_dev, K)
    e1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * K * R / (ms * 1e-3)
    tokens = sum((e_h // 14) * (e_w // 14) for e_h, e_w in
                 [enc.region_extractor.resizer.target_size(b[3] - b[1], b[2] - b[0]) for b in
                  [e.original_bbox for e in out]])

    _lib.prof_enable(True)
    run_region(frames_dev, K)
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    prof = _lib.prof_summary()
    gemm = prof.get("gemm", {"ms": 0.0, "launches": 0, "work": 0.0})
    gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else 0.0
    prof_total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    kernels = {k: {"ms_per_step": round(v["ms"] / K, 4), "launches_per_step": round(v["launches"] / K, 2),
                   "share": round(v["ms"] / prof_total_ms, 4)} for k, v in prof.items()}
    if "preprocess" in prof and prof["preprocess"]["ms"]:
        gbs = prof["preprocess"]["work"] / (prof["preprocess"]["ms"] * 1e-3) / 1e9
        kernels["preprocess"].update({"achieved_gbs": round(gbs, 1), "hbm_frac": round(gbs / peaks["hbm_gbs"], 4)})

    run_caller(frames_host, 2)
    barrier()
    e0.record()
    run_caller(frames_host, K)
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * K * R / (float(ms_e2e.item()) * 1e-3)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import hf_baseline  # the cpu_baseline leg: the one place the ours-arm may run oracle/ code
            res = hf_baseline.run_regions(n_regions=4, warmup_regions=1, frame_hw=(FRAME_H, FRAME_W))
            cpu = {"value": res["regions_per_s"], "unit": "regions/s", "cores": res["cores"], "kind": res["kind"],
                   "sample": res["sample"]}
        line = {
            "metric": "regions/s SigLIP2 masked-region route", "value": round(value, 2), "unit": "regions/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"SURVEY 8 f.4: {R} seeded detections per synthetic 1080p frame, one frame per step, each "
                                   "detection on its own canvas like scripts/extract_features.py:552-583 (no padding to a "
                                   "neighbour), all of a frame's detections in ONE ragged tower pass "
                                   "(gvl_siglip_forward_ragged), so400m tower + mean pool + REN projection",
                       "regions_per_step": R, "tokens_per_step": int(tokens), "frame": [FRAME_H, FRAME_W, 3],
                       "weights": "random init, seeds 0/3", "sharding": "replicas only (no collective on this route)",
                       "layernorm": "separate kernels" if args.no_fold_ln else "folded into the consuming GEMM epilogues",
                       "l2": "a ring of 4 distinct frames; activations of a step exceed L2 only for the larger groups"},
            "roofline": {"kernel": "gemm_bf16_cg2_kernel", "bound": "tensor", "achieved": round(gemm_tflops, 2),
                         "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": round(gemm_tflops / peaks["bf16_tflops"], 4),
                         "traffic": None, "peak_source": f"{peaks['source']} (burst: short launches, not power-capped)",
                         "note": "one GEMM launch per layer op over the ~7 k concatenated token rows of a frame's detections "
                                 "(27 row pairs: wave quantisation costs more here than at the headline's M = 46 656)"},
            "kernels": kernels, "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "regions/s", "h2d_bytes_per_step": FRAME_BYTES,
                    "d2h_bytes_per_step": R * spec.hidden * 4},
            "gpu_launches": int(launches),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=57)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak (default): --steps batches per GPU; strong: --frames in total, split across the GPUs "
                         "(BASELINE.json configs[2])")
    ap.add_argument("--frames", type=int, default=3600, help="total frames of the timeline with --scaling strong")
    ap.add_argument("--cpu-frames", type=int, default=32, help="frames of the bounded CPU-baseline sample (batch 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true", help="skip the configs[4] retrieval side measurement")
    ap.add_argument("--no-caller-api", action="store_true",
                    help="skip the side measurement of run_siglip_encoder on a list of PIL frames")
    ap.add_argument("--no-fold-ln", action="store_true",
                    help="A/B: run the 56 LayerNorm kernels per batch instead of folding them into the GEMM epilogues")
    ap.add_argument("--workload", choices=["siglip", "videomae", "regions"], default="siglip",
                    help="siglip = the headline metric (default); videomae = BASELINE.json configs[3]; regions = the "
                         "masked-region route (SURVEY 8 f.4)")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    elif a.workload == "videomae":
        main_videomae(a)
    elif a.workload == "regions":
        main_regions(a)
    else:
        main_ours(a)
