"""CPU, world_size 2, gloo: the N>1 host logic (chunk sharding + the single all-gather before clustering)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyannote_audio_b200.parallel import shard_files, shard_range


def test_shard_arithmetic():
    for n in (0, 1, 5, 591, 3591):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sorted(sum((shard_files(n, r, world) for r in range(world)), [])) == list(range(n))


def _worker(rank, world, port, C, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyannote_audio_b200.parallel import sharded_forward

    g = torch.Generator().manual_seed(0)
    full_cls = torch.randint(0, 7, (C, 589), generator=g, dtype=torch.uint8)
    full_emb = torch.randn(C, 3, 256, generator=g)
    calls = []

    def seg_fn(a, b):
        calls.append((a, b))
        return full_cls[a:b].clone()

    def emb_fn(a, b, cls):
        assert torch.equal(cls, full_cls[a:b])
        return full_emb[a:b].clone()

    cls, emb = sharded_forward(C, seg_fn, emb_fn)
    ok = torch.equal(cls, full_cls) and torch.equal(emb, full_emb) and calls == [shard_range(C, rank, world)]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("C", [21, 592])
def test_sharded_forward_all_gather_world2(C):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + C) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, C, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=10) for _ in range(2))
    assert results == [(0, True), (1, True)]


def _pool_worker(rank, world, port, q):
    """ChunkPool host logic on gloo: ragged files per rank, packed in-place all-gather, round-robin ownership."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyannote_audio_b200.parallel import ChunkPool

    per_rank = {0: [5, 1, 17], 1: [9, 4]}                   # chunks per file; rank 1 has fewer chunks in total
    g = torch.Generator().manual_seed(1)
    truth = {}
    for r in range(world):
        for i, c in enumerate(per_rank[r]):
            truth[(r, i)] = (torch.randn(c, 3, 256, generator=g), torch.randint(0, 7, (c, 589), generator=g,
                                                                                dtype=torch.uint8))
    pool = ChunkPool(pipeline=None)
    layouts = [(0, np.zeros(c), None, 160000 + 16000 * (c - 1)) for c in per_rank[rank]]
    plan = pool.plan(layouts, [f"r{rank}_f{i}" for i in range(len(layouts))])
    ok = plan["counts"] == [23, 13] and plan["cmax"] == 23 and plan["blk"] % 16 == 0
    ok = ok and [f["uri"] for f in plan["files"]] == ["r0_f0", "r0_f1", "r0_f2", "r1_f0", "r1_f1"]
    buf = torch.zeros(world * plan["blk"], dtype=torch.uint8)
    emb, cls = pool.views(buf, plan, rank)
    pos = 0
    for i, c in enumerate(per_rank[rank]):                  # "the kernels write into this rank's slice"
        emb[pos: pos + c] = truth[(rank, i)][0]
        cls[pos: pos + c] = truth[(rank, i)][1]
        pos += c
    pool.exchange(buf, plan)
    e, c, bounds, owned = pool.owned_inputs(buf, plan)
    ok = ok and owned == [gi for gi in range(5) if gi % world == rank]
    keys = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)]
    for j, gi in enumerate(owned):
        te, tc = truth[keys[gi]]
        ok = ok and torch.equal(e[bounds[j]: bounds[j + 1]], te) and torch.equal(c[bounds[j]: bounds[j + 1]], tc)
    ok = ok and pool.last_collective["bytes_sent"] == plan["blk"]
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_chunk_pool_all_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_pool_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=10) for _ in range(2)) == [(0, True), (1, True)]
