"""Debug helper: run the 75 s e2e test file through apply_batch and save the artifacts (not a test)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
from pyannote_audio_b200 import synthetic as syn
from pyannote_audio_b200.models import PyanNet, WeSpeakerResNet34
from pyannote_audio_b200.pipeline import SpeakerDiarization
dev = torch.device("cuda:0")
seg, emb = PyanNet(), WeSpeakerResNet34()
seg.load_state_dict(syn.make_segmentation_state_dict(0), strict=False)
emb.load_state_dict(syn.make_embedding_state_dict(1), strict=False)
pipe = SpeakerDiarization(segmentation=seg, embedding=emb, plda=syn.make_plda(2), device=dev)
wav = syn.make_conversation(75.0, seed=1234)
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
(_, (out, art)), = list(pipe.apply_batch([{"waveform": wav, "sample_rate": 16000, "uri": "x"}], return_artifacts=True))
os.makedirs("gpurun_out/dump", exist_ok=True)
np.savez(f"gpurun_out/dump/e2e_{tag}.npz", emb=art["embeddings"].cpu().numpy(), hard=np.asarray(art["hard_clusters"]),
         seg=art["segmentations"].cpu().numpy())
print("saved", tag, art["embeddings"].shape)
