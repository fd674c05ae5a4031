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
if len(sys.argv) > 2 and sys.argv[2] == "env":
    # A/B inside one process: SEG_PERF_ENVS="A=1;A=2,B=3" runs one timing per ';'-separated setting (knobs are read per call)
    for setting in os.environ.get("SEG_PERF_ENVS", "").split(";"):
        for kv in filter(None, setting.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v
        ctx.seg_forward(buf, off, valid); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); out = ctx.seg_forward(buf, off, valid); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        cls = out[0] if isinstance(out, tuple) else out
        print(f"[{setting}] {n} chunks seg_forward min {min(ts):.1f} ms, checksum {int(cls.sum())}", flush=True)
    sys.exit(0)
for impl in (1, 0):
    ctx.set_option("seg_rec_impl", impl)
    ctx.seg_forward(buf, off, valid); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = ctx.seg_forward(buf, off, valid); e1.record(); torch.cuda.synchronize()
    cls = out[0] if isinstance(out, tuple) else out
    print(f"rec_impl={impl}: {n} chunks seg_forward {e0.elapsed_time(e1):.1f} ms, checksum {int(cls.sum())}", flush=True)
