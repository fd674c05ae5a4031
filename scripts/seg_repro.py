"""Faithful re-run of the start of tests/test_gpu_parity.py::test_segmentation_parity in a fresh process, with
diagnostics instead of an assertion: which of {tensor-core path, fp32 CUDA-core twin} deviates from the CPU oracle,
on which chunks, and whether a second call repeats it."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import nets, pipeline as P  # noqa: E402
from pyannote_audio_b200 import synthetic as syn  # noqa: E402
from pyannote_audio_b200.inference import chunk_layout  # noqa: E402
from pyannote_audio_b200.models import get_context  # noqa: E402

dev = torch.device("cuda:0")
ctx = get_context(dev)
ctx.load_segmentation(syn.make_segmentation_state_dict(0))
ctx.load_embedding(syn.make_embedding_state_dict(1))
seg_model = nets.PyanNet()
seg_model.load_state_dict(syn.make_segmentation_state_dict(0))
seg_model.eval()
x = torch.tensor([[[2.0, 4.0], [2.0, 4.0]], [[1.0, 1.0], [1.0, 1.0]]], device=dev)
ctx.stats_pool(x)                                          # what the two earlier tests of the module do
wav = syn.make_conversation(37.3, seed=11)
chunks = P.chunk_waveform(wav)
T = wav.shape[1]
off, valid, _, _ = chunk_layout(T, 160000, 16000)
buf = torch.zeros(int(off[-1]) + 160000, dtype=torch.float32, device=dev)
buf[:T] = wav[0].to(dev)
t0 = time.time()
with torch.inference_mode():
    ref = seg_model.sincnet(chunks).transpose(1, 2).numpy()
    ref_logp = seg_model(chunks).numpy()
idle = time.time() - t0
res = {}
for name, mode in (("tc-cold", 1), ("twin", 0), ("tc-warm", 1), ("sinc-tc-only", 2), ("conv-tc-only", 3)):
    ctx.set_option("seg_conv_impl", mode)
    res[name] = ctx.sincnet_forward(buf, off, valid).cpu().numpy()
ctx.set_option("seg_conv_impl", 1)
line = [f"idle {idle:.1f}s"]
bad = False
for name, out in res.items():
    d = np.abs(out - ref)
    per_chunk = d.reshape(d.shape[0], -1).max(axis=1)
    line.append(f"{name}: max {d.max():.2e}")
    if d.max() > 2e-4:
        bad = True
        line.append(f"  <-- chunks over 2e-4: {np.nonzero(per_chunk > 2e-4)[0].tolist()}, "
                    f"{int((d > 2e-4).sum())} elements")
print(("MISMATCH " if bad else "ok ") + "; ".join(line), flush=True)
cls, logp = ctx.seg_forward(buf, off, valid, return_logp=True)
print(f"   logp vs oracle: max {np.abs(logp.cpu().numpy() - ref_logp).max():.2e}", flush=True)
