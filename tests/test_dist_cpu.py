"""The N > 1 path on CPU: world_size-2 gloo ranks each fill their contiguous chunk of the timeline index and
all-gather it; the result must equal the single-rank index, in timestamp order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gameplay_vision_llm_b200.timeline import TimelineEmbeddingIndex


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        idx = TimelineEmbeddingIndex(n, dim, fps=2.0, device="cpu", rank=rank, world=world)
        rows = idx.local_rows()
        # row i of the global index = i + 1 in every column (what this rank's frames would produce)
        rows.copy_((torch.arange(idx.lo, idx.hi, dtype=torch.float32)[:, None] + 1).expand(-1, dim).to(torch.bfloat16))
        full = idx.all_gather()
        q.put((rank, full.float()[:, 0].tolist(), idx.timestamps.tolist(), (idx.lo, idx.hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("n", [10, 7])
def test_two_rank_gloo_all_gather_assembles_timeline(n):
    world, dim = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dim, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, col, ts, span in results:
        assert col == [float(i + 1) for i in range(n)], f"rank {rank} sees {col}"
        assert ts == [i / 2.0 for i in range(n)]
    assert results[0][3][1] == results[1][3][0]  # chunks are contiguous
