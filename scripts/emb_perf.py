"""embedding perf probe: time emb_forward on N ten-second segments (not a benchmark)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pyannote_audio_b200 import ops, synthetic as syn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ctx = ops.Context(dev)
wav = syn.make_conversation(60.0, seed=3)
off = (np.arange(n, dtype=np.int64) % 50) * 16000
valid = np.full(n, 160000, dtype=np.int32)
buf = wav[0].to(dev).contiguous()
ctx.load_embedding(syn.make_embedding_state_dict(1))
masks = torch.ones((n, 3, 589), dtype=torch.uint8, device=dev)
import os, subprocess
# A/B inside one process (same box, same clocks): EMB_PERF_ENVS="A=1;A=2,B=3" runs one timing per ';'-separated setting.
# EMB_PERF_ITERS=N times N back-to-back passes with ONE pair of events (sustained load: the power cap shows) and reports
# the SM clock / power draw sampled at the end of the run.
iters = int(os.environ.get("EMB_PERF_ITERS", "0"))
def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                              capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception:
        return "n/a"
for setting in os.environ.get("EMB_PERF_ENVS", "").split(";"):
    for kv in filter(None, setting.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    for _ in range(3):
        out = ctx.emb_forward(buf, off, valid, masks)
    torch.cuda.synchronize()
    if iters > 0:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            out = ctx.emb_forward(buf, off, valid, masks)
            if i == iters - 8: mid = smi()                  # sampled while the queue is still full
        e1.record(); torch.cuda.synchronize()
        print(f"[{setting}] sustained: {e0.elapsed_time(e1) / iters:.3f} ms per pass over {iters} passes; sm MHz, W near the end: {mid}; "
              f"checksum {float(out.double().sum()):.6f}", flush=True)
        continue
    ts = []
    for _ in range(7):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = ctx.emb_forward(buf, off, valid, masks); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"[{setting}] emb_forward {n} segments: min {min(ts):.3f} ms  median {sorted(ts)[3]:.3f} ms  checksum {float(out.double().sum()):.6f}", flush=True)
