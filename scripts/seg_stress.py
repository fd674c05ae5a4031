"""Race hunt for the SincNet front end (not a benchmark): the tensor-core path is run cold (first call of a fresh
process) and then `reps` more times; every result is compared bitwise with the first warm result of the same mode
(the kernels are deterministic, so any difference is a race) and with the fp32 CUDA-core twin (seg_conv_impl = 0).
usage: python scripts/seg_stress.py <mode: 1 all tensor-core | 2 sinc only | 3 conv1d only> [reps] [seconds]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyannote_audio_b200 import ops, synthetic as syn  # noqa: E402
from pyannote_audio_b200.inference import chunk_layout  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 37.3
dev = torch.device("cuda:0")
ctx = ops.Context(dev)
ctx.load_segmentation(syn.make_segmentation_state_dict(0))
wav = syn.make_conversation(seconds, seed=11)
T = wav.shape[1]
off, valid, _, _ = chunk_layout(T, 160000, 16000)
buf = torch.zeros(int(off[-1]) + 160000, dtype=torch.float32, device=dev)
buf[:T] = wav[0].to(dev)
ctx.set_option("seg_conv_impl", mode)
cold = ctx.sincnet_forward(buf, off, valid).cpu().numpy()          # the very first launch of these kernels
ctx.set_option("seg_conv_impl", 0)
ref = ctx.sincnet_forward(buf, off, valid).cpu().numpy()
ctx.set_option("seg_conv_impl", mode)
first = ctx.sincnet_forward(buf, off, valid).cpu().numpy()


def describe(a, b):
    d = np.abs(a - b)
    bad = np.argwhere(d.reshape(d.shape[0], -1).max(axis=1) > 0).ravel()
    return f"max {d.max():.3e}, {int((d > 0).sum())} elements in chunks {bad[:12].tolist()}{'...' if len(bad) > 12 else ''}"


print(f"[mode {mode}] {len(off)} chunks; cold vs twin: max {np.abs(cold - ref).max():.3e}; warm vs twin: max "
      f"{np.abs(first - ref).max():.3e}; cold == warm: {np.array_equal(cold, first)}"
      + ("" if np.array_equal(cold, first) else "  <-- " + describe(cold, first)), flush=True)
nbad = 0
for r in range(reps):
    out = ctx.sincnet_forward(buf, off, valid).cpu().numpy()
    if not np.array_equal(out, first):
        nbad += 1
        if nbad <= 5:
            print(f"[mode {mode}] rep {r}: differs from the first warm result: {describe(out, first)}; vs twin max "
                  f"{np.abs(out - ref).max():.3e}", flush=True)
print(f"[mode {mode}] {nbad} of {reps} warm repetitions differ", flush=True)
