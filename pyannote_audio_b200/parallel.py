"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

The reference has no inference-time parallelism at all; the path shards naturally:
  * many files  -> file-level round-robin, no data-path collective (bench.py, weak scaling);
  * one long file -> contiguous chunk ranges per rank through segmentation and embedding (chunk c only needs
    samples [c*step, c*step+160000)), then ONE all-gather of the per-chunk results (powerset classes (C,589) u8 and
    embeddings (C,3,256) f32) before the per-file clustering barrier.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n chunks: first (n % world) ranks get one extra."""
    base, extra = divmod(n, world)
    a = rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def shard_files(n: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n, world))


def all_gather_rows(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """All-gather of row blocks with per-rank row counts known to everyone (derived from shard_range)."""
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def sharded_forward(num_chunks: int, seg_fn: Callable[[int, int], torch.Tensor],
                    emb_fn: Callable[[int, int, torch.Tensor], torch.Tensor], group=None):
    """Runs seg_fn / emb_fn on this rank's chunk range and all-gathers the results.

    seg_fn(a, b) -> (b-a, 589) uint8 powerset classes; emb_fn(a, b, classes) -> (b-a, 3, 256) float32.
    Returns (classes (C,589) u8, embeddings (C,3,256) f32) identical on every rank.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [shard_range(num_chunks, r, world)[1] - shard_range(num_chunks, r, world)[0] for r in range(world)]
    a, b = shard_range(num_chunks, rank, world)
    cls = seg_fn(a, b)
    emb = emb_fn(a, b, cls)
    return all_gather_rows(cls.contiguous(), counts, group), all_gather_rows(emb.contiguous(), counts, group)


def apply_sharded(pipeline, file, group=None, **kwargs):
    """SpeakerDiarization over ONE long file with its chunks sharded across the ranks of `group`.

    Every rank holds the (host) waveform, uploads only the sample range its chunks touch, runs PyanNet + WeSpeaker on
    them, all-gathers (NCCL over NVLink) classes + embeddings, and rank-locally finishes clustering / reconstruction
    (cheap, replicated) so that every rank returns the same DiarizeOutput.
    """
    from . import ops
    from .inference import chunk_layout
    from .models import get_context

    ctx = get_context(pipeline.device)
    file = pipeline._audio.validate_file(file)
    wav, sr = pipeline._audio(file)
    off, valid, _, _ = chunk_layout(wav.shape[1], ops.CHUNK, round(pipeline._segmentation.step * sr))
    C = len(off)
    state = {}

    def seg_fn(a, b):
        lo, hi = int(off[a]), int(off[b - 1]) + ops.CHUNK
        buf = torch.zeros(hi - lo, dtype=torch.float32, device=ctx.device)
        n = max(0, min(hi, wav.shape[1]) - lo)
        buf[:n].copy_(wav[0, lo: lo + n])
        state.update(buf=buf, off=off[a:b] - lo, valid=valid[a:b])
        return pipeline._segmentation.model.forward_chunks(buf, state["off"], state["valid"])

    def emb_fn(a, b, cls):
        seg = ctx.powerset_to_multilabel(cls)
        return pipeline.embedding.forward_chunks(state["buf"], state["off"], state["valid"], pipeline._masks(seg))

    cls, emb = sharded_forward(C, seg_fn, emb_fn, group)
    seg = ctx.powerset_to_multilabel(cls)
    from .pipeline import set_num_speakers

    ns, mn, mx = set_num_speakers(kwargs.get("num_speakers"), kwargs.get("min_speakers"), kwargs.get("max_speakers"))
    return pipeline._finish_file(ctx, file, seg, emb, ns, mn, mx, kwargs.get("hook"), False)
