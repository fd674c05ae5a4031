"""segmentation perf probe: time seg_forward on N chunks (not a benchmark)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from pyannote_audio_b200 import ops, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
dev = torch.device("cuda:0")
ctx = ops.Context(dev)
wav = syn.make_conversation(60.0, seed=3)
off = (np.arange(n, dtype=np.int64) % 50) * 16000
valid = np.full(n, 160000, dtype=np.int32)
buf = wav[0].to(dev).contiguous()
ctx.load_segmentation(syn.make_segmentation_state_dict(0))
import os
if len(sys.argv) > 2 and sys.argv[2] == "tc":
    # A/B over the tensor-core recurrence knobs (read per call by the library)
    for np_, pf, asy in ((0, 0, 0), (0, 0, 1), (-1, 0, 1), (0, 2, 1), (0, 0, 0), (0, 0, 1)):
        os.environ["B200_LSTM_NP"], os.environ["B200_LSTM_PF"], os.environ["B200_LSTM_ASYNC"] = str(np_), str(pf), str(asy)
        ctx.seg_forward(buf, off, valid); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = ctx.seg_forward(buf, off, valid); e1.record(); torch.cuda.synchronize()
        cls = out[0] if isinstance(out, tuple) else out
        print(f"np={np_} pf={pf} async={asy}: {n} chunks seg_forward {e0.elapsed_time(e1):.1f} ms, checksum {int(cls.sum())}", flush=True)
    sys.exit(0)
for impl in (1, 0):
    ctx.set_option("seg_rec_impl", impl)
    ctx.seg_forward(buf, off, valid); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = ctx.seg_forward(buf, off, valid); e1.record(); torch.cuda.synchronize()
    cls = out[0] if isinstance(out, tuple) else out
    print(f"rec_impl={impl}: {n} chunks seg_forward {e0.elapsed_time(e1):.1f} ms, checksum {int(cls.sum())}", flush=True)
