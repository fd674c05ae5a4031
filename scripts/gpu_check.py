"""Stage-by-stage CUDA-vs-oracle diagnostics (development aid; the assertions live in tests/).
Usage: python scripts/gpu_check.py --stage seg|emb_simt|emb_tc|emb_tc_s1|post|perf
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import nets, pipeline as P  # noqa: E402
from pyannote_audio_b200 import ops, synthetic as syn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stage", default="seg")
ap.add_argument("--seconds", type=float, default=25.0)
ap.add_argument("--impl", type=int, default=1)
args = ap.parse_args()
torch.set_num_threads(max(1, torch.get_num_threads()))
dev = torch.device("cuda:0")
ctx = ops.Context(dev)
seg_sd, emb_sd = syn.make_segmentation_state_dict(0), syn.make_embedding_state_dict(1)
wav = syn.make_conversation(args.seconds, seed=1234)
chunks = P.chunk_waveform(wav)
C = chunks.shape[0]
T = wav.shape[1]
off = np.arange(C, dtype=np.int64) * 16000
valid = np.minimum(160000, T - off).astype(np.int32)
# device buffer padded so that every chunk window is addressable
wav_dev = torch.zeros(int(off[-1]) + 160000, dtype=torch.float32, device=dev)
wav_dev[:T] = wav[0].to(dev)
print(f"[{args.stage}] chunks={C} T={T}", flush=True)


def report(name, a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    print(f"  {name}: max|d|={d.max():.3e} mean|d|={d.mean():.3e} ref_rms={np.sqrt((b**2).mean()):.3e} "
          f"nan={np.isnan(a).sum()}", flush=True)


if args.stage == "seg":
    ctx.load_segmentation(seg_sd)
    om = nets.PyanNet(); om.load_state_dict(seg_sd); om.eval()
    with torch.inference_mode():
        ref_sinc = om.sincnet(chunks).transpose(1, 2).numpy()
        ref_logp = om(chunks).numpy()
    got_sinc = ctx.sincnet_forward(wav_dev, off, valid).cpu().numpy()
    report("sincnet (C,589,60)", got_sinc, ref_sinc)
    cls, logp = ctx.seg_forward(wav_dev, off, valid, return_logp=True)
    torch.cuda.synchronize()
    report("logp (C,589,7)", logp.cpu().numpy(), ref_logp)
    ref_cls = ref_logp.argmax(-1)
    mism = (cls.cpu().numpy() != ref_cls)
    top2 = np.sort(ref_logp, -1)
    margin = top2[..., -1] - top2[..., -2]
    print(f"  class mismatches: {mism.sum()} / {mism.size}; margins at mismatches: {margin[mism][:10]}")
    ml = ctx.powerset_to_multilabel(cls).cpu().numpy()
    ref_ml = nets.powerset_to_multilabel(torch.from_numpy(ref_logp), nets.powerset_mapping()).numpy()
    print("  multilabel equal (where classes equal):", np.array_equal(ml[~mism], ref_ml[~mism].astype(np.uint8)))

if args.stage.startswith("emb"):
    impl = {"emb_simt": 0, "emb_tc": 1, "emb_tc_s1": 2}[args.stage]
    ctx.load_embedding(emb_sd)
    ctx.set_option("conv_impl", impl)
    om = nets.WeSpeakerResNet34(); om.load_state_dict(emb_sd); om.eval()
    n = min(C, 4)
    with torch.inference_mode():
        ref_fb = om.compute_fbank(chunks[:n])
        ref_frames = om.resnet.forward_frames(ref_fb)
    fb = ctx.emb_fbank(wav_dev, off[:n], valid[:n])
    report("fbank (n,998,80)", fb.cpu().numpy(), ref_fb.numpy())
    t0 = time.time()
    frames = ctx.emb_trunk(ref_fb.to(dev))
    torch.cuda.synchronize()
    print(f"  trunk impl={impl} took {time.time()-t0:.3f}s", flush=True)
    report("trunk frames (n,256,10,125)", frames.cpu().numpy(), ref_frames.numpy())
    rng = np.random.default_rng(0)
    masks = (rng.uniform(size=(n, 3, 589)) < 0.5).astype(np.uint8)
    masks[0, 2] = 0
    with torch.inference_mode():
        ref_emb = om.forward_embedding(ref_frames, weights=torch.from_numpy(masks.astype(np.float32))).numpy()
    emb = ctx.emb_forward(wav_dev, off[:n], valid[:n], torch.from_numpy(masks).to(dev)).cpu().numpy()
    report("embeddings (n,3,256)", emb, ref_emb)
    cos = (emb * ref_emb).sum(-1) / (np.linalg.norm(emb, axis=-1) * np.linalg.norm(ref_emb, axis=-1))
    print("  cosine distance to oracle:", np.round(1 - cos, 6).ravel())

if args.stage == "conv_ab":
    # A/B of a tensor-core conv variant against the CUDA-core reference conv (same fp16 inputs/weights)
    ctx.load_embedding(emb_sd)
    n = min(C, 6)
    fb = ctx.emb_fbank(wav_dev, off[:n], valid[:n])
    ctx.set_option("conv_impl", 0)
    ref = ctx.emb_trunk(fb).cpu().numpy()
    ctx.set_option("conv_impl", args.impl)
    got = ctx.emb_trunk(fb)
    torch.cuda.synchronize()
    report(f"trunk impl={args.impl} vs SIMT", got.cpu().numpy(), ref)
    big = fb.repeat(43, 1, 1)[:256].contiguous()
    ctx.set_option("conv_fuse", 0)
    got_u = ctx.emb_trunk(fb)
    torch.cuda.synchronize()
    report("fused layer1 blocks vs unfused", got.cpu().numpy(), got_u.cpu().numpy())
    for impl, fuse in ((1, 0), (args.impl, 0), (args.impl, 1)):
        ctx.set_option("conv_fuse", fuse)
        ctx.set_option("conv_impl", impl)
        ctx.emb_trunk(big); torch.cuda.synchronize()
        t0 = time.time(); ctx.emb_trunk(big); torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"  impl={impl} fuse={fuse}: 256 segments trunk {dt*1e3:.1f} ms -> {256*45.18e9/dt/1e12:.0f} TFLOP/s", flush=True)

if args.stage == "post":
    x = torch.tensor([[[2.0, 4.0], [2.0, 4.0]], [[1.0, 1.0], [1.0, 1.0]]], device=dev)
    print("  stats_pool weightless:", ctx.stats_pool(x).cpu().numpy().round(4))
    w = torch.tensor([[0.5, 0.01], [0.2, 0.1]], device=dev)
    print("  stats_pool one speaker:", ctx.stats_pool(x, w).cpu().numpy().round(4))
    rng = np.random.default_rng(1)
    Cn = 40
    seg = (rng.uniform(size=(Cn, 589, 3)) < 0.3).astype(np.float32)
    swf = P.SWF(seg, P.SW(0.0, 10.0, 1.0))
    frames = P.SW(*nets.sincnet_receptive_field())
    ref_count = P.speaker_count(swf, frames, (0.0, 0.0))
    sf = P.chunk_start_frames(Cn, frames)
    F = len(ref_count.data)
    seg_dev = torch.from_numpy(seg.astype(np.uint8)).to(dev)
    count = ctx.speaker_count(seg_dev, sf, F)
    print("  speaker_count equal:", np.array_equal(count.cpu().numpy(), ref_count.data[:, 0]), F)
    hard = rng.integers(-1, 4, size=(Cn, 3)).astype(np.int8)
    hard[hard == -1] = -2
    cnt = SWF = P.SWF(np.minimum(ref_count.data, 3).astype(np.int8), ref_count.sw)
    ref_d = P.reconstruct(swf, hard, cnt)
    K = int(hard.max()) + 1
    Kout = max(K, int(cnt.data.max()))
    d = ctx.reconstruct(seg_dev, hard, sf, F, torch.from_numpy(cnt.data[:, 0].astype(np.uint8)).to(dev), Kout)
    print("  reconstruct equal:", np.array_equal(d.cpu().numpy(), ref_d.data.astype(np.uint8)), ref_d.data.shape)
    clean, active = ctx.clean_frames(seg_dev)
    single = (seg.sum(2, keepdims=True) == 1)
    print("  clean_frames equal:", np.array_equal(clean.cpu().numpy(), (seg * single).sum(1).astype(np.int32)),
          np.array_equal(active.cpu().numpy() > 0, seg.sum(1) > 0))

if args.stage == "perf":
    ctx.load_segmentation(seg_sd)
    ctx.load_embedding(emb_sd)
    n = 592
    wav_big = syn.make_conversation(600.0, seed=7)
    Tb = wav_big.shape[1]
    offb = np.arange(n, dtype=np.int64)[: (Tb - 160000) // 16000 + 1] * 16000
    validb = np.full(len(offb), 160000, dtype=np.int32)
    wb = wav_big[0].to(dev).contiguous()
    for name, fn in (("seg", lambda: ctx.seg_forward(wb, offb, validb)),):
        fn(); torch.cuda.synchronize()
        t0 = time.time(); fn(); torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"  {name}: {len(offb)} chunks in {dt*1e3:.1f} ms -> {len(offb)/dt:.0f} chunks/s", flush=True)
    masks = torch.ones((len(offb), 3, 589), dtype=torch.uint8, device=dev)
    for impl in (1, 0):
        ctx.set_option("conv_impl", impl)
        m = len(offb) if impl else 64
        ctx.emb_forward(wb, offb[:m], validb[:m], masks[:m]); torch.cuda.synchronize()
        t0 = time.time(); ctx.emb_forward(wb, offb[:m], validb[:m], masks[:m]); torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"  emb impl={impl}: {m} chunks in {dt*1e3:.1f} ms -> {m/dt:.0f} chunks/s "
              f"({m*45.23e9/dt/1e12:.1f} TFLOP/s)", flush=True)
print("done", flush=True)
