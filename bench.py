#!/usr/bin/env python
"""Benchmark of the community-1 diarization hot path (BASELINE.json metric: audio-hours/sec).

A "step" = one pass of the whole pipeline (segmentation -> embeddings -> clustering -> reconstruction -> annotations)
over a batch of synthetic 10-minute files (12 per GPU by default = BASELINE.json configs[4], 100 x 10 min over
8 GPUs, scaled to one GPU).  Weak scaling: every rank processes its own files, there is no data-path collective.

  value : audio-hours/sec, waveforms already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through SpeakerDiarization.apply_batch with HOST waveforms (H2D + D2H inside the timed region)
  roofline     : ResNet34 trunk conv kernels (~98 % of the FLOPs) measured live with CUDA events
  cpu_baseline : the CPU oracle (reference-equivalent: 3 trunk passes per chunk) on a bounded sample, rank 0, N=1

`--impl reference` times the CPU oracle arm (the reference package itself cannot be imported in this image:
lightning / pyannote.core / asteroid_filterbanks are absent).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-hours/sec (RTF) community-1 diarization, 16kHz mono, 1/2/4/8 B200"
TRUNK_FLOP_PER_SEGMENT = 45.18e9   # 33 conv3x3 + 3 conv1x1 of ResNet34 at (80 x 998), SURVEY.md section 8(d)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files-per-gpu", type=int, default=12)
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--cpu-sample-seconds", type=float, default=12.0)
    ap.add_argument("--min-warmup", type=int, default=3, help="lower only when profiling under ncu")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_models():
    from oracle import nets, pipeline as P
    from pyannote_audio_b200 import synthetic as syn

    seg, emb = nets.PyanNet(), nets.WeSpeakerResNet34()
    seg.load_state_dict(syn.make_segmentation_state_dict(0))
    emb.load_state_dict(syn.make_embedding_state_dict(1))
    return seg.eval(), emb.eval(), P.PLDA(**syn.make_plda(2))


def cpu_pass(seconds, models, seed=4242):
    """One reference-equivalent CPU pass (3 trunk forwards per chunk like the reference) -> wall seconds."""
    from oracle import pipeline as P
    from pyannote_audio_b200 import synthetic as syn

    seg, emb, plda = models
    wav = syn.make_conversation(seconds, seed=seed)
    t0 = time.perf_counter()
    P.apply(seg, emb, plda, wav, seg_batch=32, emb_batch=8, share_trunk=False)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    torch.set_num_threads(min(32, os.cpu_count() or 1))   # conv on >32 threads oversubscribes and gets slower
    models = oracle_models()
    secs = args.cpu_sample_seconds
    for _ in range(min(args.warmup, 1)):
        cpu_pass(min(secs, 12.0), models)
    times = [cpu_pass(secs, models, seed=4242 + i) for i in range(max(1, args.steps))]
    t = float(np.mean(times))
    value = (secs / 3600.0) / t
    sample = (f"CPU oracle (PyTorch CPU fp32, 3 trunk passes per chunk as the reference), one {secs:g} s synthetic "
              f"file end-to-end per step")
    line = {"metric": METRIC, "value": value, "unit": "audio-hours/sec", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"community-1 pipeline, {args.files_per_gpu} x {args.minutes:g} min synthetic files "
                                   f"per GPU (CPU arm: bounded sample, see cpu_baseline.sample)"},
            "rtf": t / secs,
            "cpu_baseline": {"value": value, "unit": "audio-hours/sec", "cores": torch.get_num_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "audio-hours/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch.distributed as dist

    from pyannote_audio_b200 import synthetic as syn
    from pyannote_audio_b200.models import PyanNet, WeSpeakerResNet34, get_context
    from pyannote_audio_b200.pipeline import SpeakerDiarization

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    seg, emb = PyanNet(), WeSpeakerResNet34()
    seg.load_state_dict(syn.make_segmentation_state_dict(0), strict=False)
    emb.load_state_dict(syn.make_embedding_state_dict(1), strict=False)
    pipe = SpeakerDiarization(segmentation=seg, embedding=emb, plda=syn.make_plda(2), device=dev)
    ctx = get_context(dev)
    nfiles = args.files_per_gpu
    files = []
    for i in range(nfiles):
        wav = syn.make_conversation(args.minutes * 60.0, seed=1000 + rank * 1000 + i)
        files.append({"waveform": wav.pin_memory(), "sample_rate": 16000, "uri": f"r{rank}_f{i}"})
    audio_hours = nfiles * args.minutes / 60.0
    h2d = sum(f["waveform"].numel() * 4 for f in files)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    resident = pipe.upload(files)

    def step_resident():
        for _ in pipe.run_resident(resident):
            pass

    d2h = [0]

    def step_e2e():
        pipe.d2h_bytes = 0
        for _ in pipe.apply_batch(files):
            pass
        d2h[0] = pipe.d2h_bytes

    for _ in range(max(args.min_warmup, args.warmup)):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.set_option("profile", 1)
    ctx.timer("trunk"); ctx.timer("seg")
    l0 = ctx.launch_count
    ms_resident = timed(step_resident, args.steps)
    launches = (ctx.launch_count - l0) // max(1, args.steps)
    trunk_ms, trunk_segments = ctx.timer("trunk")
    seg_ms, seg_chunks = ctx.timer("seg")
    ctx.set_option("profile", 0)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, max(1, args.steps))
    value = world * audio_hours / (ms_resident / 1e3)
    e2e = world * audio_hours / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    peak = pk.get("bf16_tflops_sustained", 1400.0)
    achieved = trunk_segments * TRUNK_FLOP_PER_SEGMENT / (trunk_ms / 1e3) / 1e12 if trunk_ms > 0 else 0.0
    # DRAM traffic of one 256-segment trunk pass (stem + 35 tcgen05 conv launches) from ncu
    # (profiles/r01_trunk_traffic_256.csv: 29.66 GB read + 18.70 GB written; algorithmic minimum 50.9 GB counting every
    # activation tensor once per read/write, i.e. no re-read waste; small layers hit L2)
    traffic_256 = 48.354e9
    roofline = {"bound": "tensor",
                "kernel": "ResNet34 trunk = conv_tc4/conv_tc3/conv_tc kernels, 36 dependent launches per 256-segment sub-batch",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": traffic_256, "traffic_unit": "bytes per 256-segment trunk pass (ncu dram read+write)",
                "algorithmic_flop_per_launch_unit": 256 * TRUNK_FLOP_PER_SEGMENT,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (fp16 = same tensor rate)"
                if pk else "fallback 1.4 PFLOP/s sustained",
                "trunk_ms_per_step": trunk_ms / args.steps, "seg_ms_per_step": seg_ms / args.steps}
    cpu = None
    if args.gpus == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        models = oracle_models()
        t = cpu_pass(args.cpu_sample_seconds, models)
        cpu = {"value": (args.cpu_sample_seconds / 3600.0) / t, "unit": "audio-hours/sec",
               "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"CPU oracle (3 trunk passes per chunk, as the reference), one {args.cpu_sample_seconds:g} s "
                         f"synthetic file end-to-end ({t:.1f} s wall)"}
    line = {"metric": METRIC, "value": value, "unit": "audio-hours/sec", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.min_warmup, args.warmup), "ms_per_step": ms_resident, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 tensor-core trunk (f32 accumulate) + split-f16x3 tensor-core segmentation (f32-level accuracy) + f64 clustering",
            "data": "synthetic",
            "config": {"workload": f"community-1 diarization pipeline end-to-end, {nfiles} x {args.minutes:g} min "
                                   f"synthetic 16 kHz mono files per GPU (BASELINE.json configs[4] scaled per GPU)",
                       "files_per_gpu": nfiles, "chunks_per_gpu": int(sum(len(r[1]) for r in resident["layouts"])),
                       "audio_hours_per_step_per_gpu": audio_hours, "parallelism": f"file-sharded x{world}",
                       "l2": f"inputs larger than L2: {h2d / 1e6:.0f} MB of waveform per step"},
            "rtf": (ms_resident / 1e3) / (audio_hours * 3600.0) / world,
            "e2e": {"value": e2e, "unit": "audio-hours/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h[0],
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
