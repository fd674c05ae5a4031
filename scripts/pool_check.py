"""N-GPU check of parallel.ChunkPool (torchrun --nproc-per-node N): every rank holds 3 files of different lengths, the
pool runs the networks on them, exchanges embeddings + classes (p2p: pushed from the GEMM epilogue over symmetric
memory; nccl: one all-gather) and finishes file g on rank g mod N.  Every output must equal the plain single-GPU
pipeline on the same waveform (regenerated from its seed), for both collectives."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pyannote_audio_b200 import synthetic as syn  # noqa: E402
from pyannote_audio_b200.models import PyanNet, WeSpeakerResNet34  # noqa: E402
from pyannote_audio_b200.parallel import ChunkPool  # noqa: E402
from pyannote_audio_b200.pipeline import SpeakerDiarization  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
seg, emb = PyanNet(), WeSpeakerResNet34()
seg.load_state_dict(syn.make_segmentation_state_dict(0), strict=False)
emb.load_state_dict(syn.make_embedding_state_dict(1), strict=False)
pipe = SpeakerDiarization(segmentation=seg, embedding=emb, plda=syn.make_plda(2), device=dev)


def make(r, i):
    return syn.make_conversation(40.0 + 13.0 * i + 7.0 * r, seed=500 + 10 * r + i)


files = [{"waveform": make(rank, i), "sample_rate": 16000, "uri": f"r{rank}_f{i}"} for i in range(3)]
ok = True
for mode in ("p2p", "nccl"):
    pool = ChunkPool(pipe, collective=mode)
    for rep in range(2):                                   # second pass reuses the symmetric buffer
        got = {meta["uri"]: out for meta, out in pool.apply_batch(files)}
    torch.cuda.synchronize()
    owned = [g for g in range(3 * world) if g % world == rank]
    assert len(got) == len(owned), (mode, rank, sorted(got))
    for g in owned:
        r, i = divmod(g, 3)
        uri = f"r{r}_f{i}"
        ref = pipe.apply({"waveform": make(r, i), "sample_rate": 16000, "uri": uri})
        a = [(s.start, s.end, lab) for s, _, lab in got[uri].speaker_diarization.itertracks(yield_label=True)]
        b = [(s.start, s.end, lab) for s, _, lab in ref.speaker_diarization.itertracks(yield_label=True)]
        same = a == b and bool((got[uri].speaker_embeddings == ref.speaker_embeddings).all())
        ok = ok and same
        print(f"[pool_check] rank {rank} mode {pool.collective} (asked {mode}) file {uri} computed on rank {r}: "
              f"{'identical' if same else 'DIFFERENT'} ({len(a)} segments), collective {pool.collective_ms():.3f} ms, "
              f"{pool.last_collective}", flush=True)
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("[pool_check] ALL IDENTICAL" if int(t.item()) else "[pool_check] MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
