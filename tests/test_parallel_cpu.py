"""world_size-2 gloo test of the data-parallel plumbing (host logic; the per-image results are stand-ins produced by
the oracle NMS so that the test needs no GPU): ONE all-gather of the fixed-size per-shard detection records == the single-process
result of the same global batch, ordered by global image index."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_detections(global_batch, seed=0):
    rs = np.random.RandomState(seed)
    out = []
    for i in range(global_batch):
        n = int(rs.randint(0, 6))
        out.append(rs.uniform(0, 100, (n, 5)) if n else None)
    return out


def _worker(rank, world, port, global_batch, q):
    sys.path.insert(0, ROOT)
    from kg_instance_segmentation_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _fake_detections(global_batch)
        lo, hi = parallel.shard_range(global_batch, rank, world)
        local_max = max(parallel.shard_range(global_batch, r, world)[1] - parallel.shard_range(global_batch, r, world)[0]
                        for r in range(world))
        rec = parallel.pack_detections(full[lo:hi], kmax=8)
        got = parallel.unpack_records(parallel.all_gather_records(rec, local_max))
        ok = len(got) == global_batch and all((a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b))
                                              for a, b in zip(got, full))
        q.put((rank, ok, len(got)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("global_batch", [8, 7])
def test_all_gather_of_sharded_detections_matches_single_process(global_batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + global_batch
    procs = [ctx.Process(target=_worker, args=(r, 2, port, global_batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_shard_ranges_cover_the_batch():
    from kg_instance_segmentation_b200 import parallel
    for B in (1, 7, 32, 33):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(B, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
