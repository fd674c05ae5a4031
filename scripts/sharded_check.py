"""2-GPU check of parallel.apply_sharded: one long file, chunks sharded over the ranks, single NCCL all-gather of
classes + embeddings, result must equal the single-GPU pipeline bit-for-bit (torchrun --nproc-per-node 2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from pyannote_audio_b200 import synthetic as syn  # noqa: E402
from pyannote_audio_b200.models import PyanNet, WeSpeakerResNet34  # noqa: E402
from pyannote_audio_b200.parallel import apply_sharded  # noqa: E402
from pyannote_audio_b200.pipeline import SpeakerDiarization  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
seg, emb = PyanNet(), WeSpeakerResNet34()
seg.load_state_dict(syn.make_segmentation_state_dict(0), strict=False)
emb.load_state_dict(syn.make_embedding_state_dict(1), strict=False)
pipe = SpeakerDiarization(segmentation=seg, embedding=emb, plda=syn.make_plda(2), device=dev)
wav = syn.make_conversation(600.0, seed=31)
file = {"waveform": wav, "sample_rate": 16000, "uri": "long"}
out = apply_sharded(pipe, file)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
dist.barrier(); t0.record()
out = apply_sharded(pipe, file)
t1.record(); torch.cuda.synchronize()
ms_sharded = t0.elapsed_time(t1)
if rank == 0:
    ref = pipe.apply(file)
    t0.record(); ref = pipe.apply(file); t1.record(); torch.cuda.synchronize()
    a = [(s.start, s.end, l) for s, _, l in out.speaker_diarization.itertracks(yield_label=True)]
    b = [(s.start, s.end, l) for s, _, l in ref.speaker_diarization.itertracks(yield_label=True)]
    print(f"sharded == single-GPU: {a == b}  ({len(a)} segments); world={dist.get_world_size()} "
          f"sharded {ms_sharded:.1f} ms vs single {t0.elapsed_time(t1):.1f} ms", flush=True)
    assert a == b
dist.barrier()
dist.destroy_process_group()
