"""Small embedding / segmentation workload for ncu captures (not a benchmark)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyannote_audio_b200 import ops, synthetic as syn  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "emb"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda:0")
ctx = ops.Context(dev)
wav = syn.make_conversation(10.0 + n, seed=3)
off = np.arange(n, dtype=np.int64) * 16000
valid = np.full(n, 160000, dtype=np.int32)
buf = torch.zeros(int(off[-1]) + 160000, device=dev)
buf[: wav.shape[1]] = wav[0].to(dev)[: buf.numel()]
if what == "emb":
    ctx.load_embedding(syn.make_embedding_state_dict(1))
    masks = torch.ones((n, 3, 589), dtype=torch.uint8, device=dev)
    for _ in range(2):
        ctx.emb_forward(buf, off, valid, masks)
else:
    ctx.load_segmentation(syn.make_segmentation_state_dict(0))
    for _ in range(2):
        ctx.seg_forward(buf, off, valid)
torch.cuda.synchronize()
print("done")
