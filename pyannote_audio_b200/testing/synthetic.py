"""Deterministic synthetic weights, PLDA and audio for parity tests and benchmarks.

No pretrained community-1 checkpoint exists offline, so every test/bench uses seeded random weights
with the *reference's state-dict key names and shapes* (real checkpoints drop in unchanged):

* segmentation: ``PyanNet`` (SincNet + 4-layer BiLSTM(128) + 2xLinear(128) + Linear(7));
  keys as in /root/reference/src/pyannote/audio/models/segmentation/PyanNet.py:92-161 and
  models/blocks/sincnet.py:41-79 (module tree: tutorials/training_a_model.ipynb:1001-1016)
* embedding: ``WeSpeakerResNet34``; keys as in models/embedding/wespeaker/resnet.py:233-252
* PLDA: ``xvec_transform.npz{mean1,mean2,lda}`` + ``plda.npz{mu,tr,psi}`` (utils/vbx.py:195-199)

Audio: a synthetic multi-speaker "conversation" (harmonic sources, 3-6 Hz amplitude modulation,
Markov turn-taking with some overlap, -30 dB noise floor), float32 mono 16 kHz in [-1, 1].
"""

from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

SAMPLE_RATE = 16000

# ----------------------------------------------------------------------------------------
# segmentation state dict
# ----------------------------------------------------------------------------------------


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * bound


from ..models import _mel_sinc_init, sinc_buffers  # noqa: E402,F401  (ParamSincFB default initialisation)


def make_segmentation_state_dict(seed: int = 0, lstm_layers: int = 4, num_classes: int = 7,
                                 logit_scale: float = 6.0, ih_gain: float = 4.0, hh_gain: float = 1.5,
                                 lin_gain: float = 3.0,
                                 fitted_classifier: bool = True) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    sd["sincnet.wav_norm1d.weight"] = 1.0 + 0.1 * torch.randn(1, generator=g)
    sd["sincnet.wav_norm1d.bias"] = 0.05 * torch.randn(1, generator=g)
    low, band = _mel_sinc_init()
    # jitter the learnable cut-offs a little so they are not the textbook initialisation
    sd["sincnet.conv1d.0.filterbank.low_hz_"] = low * (1.0 + 0.05 * torch.randn(low.shape, generator=g))
    sd["sincnet.conv1d.0.filterbank.band_hz_"] = band * (1.0 + 0.05 * torch.randn(band.shape, generator=g))
    window_, n_ = sinc_buffers()
    sd["sincnet.conv1d.0.filterbank.window_"] = window_
    sd["sincnet.conv1d.0.filterbank.n_"] = n_
    for i, (cin, cout) in zip((1, 2), ((80, 60), (60, 60))):
        bound = 1.0 / math.sqrt(cin * 5)
        sd[f"sincnet.conv1d.{i}.weight"] = _uniform(g, (cout, cin, 5), bound)
        sd[f"sincnet.conv1d.{i}.bias"] = _uniform(g, (cout,), bound)
    for i, c in enumerate((80, 60, 60)):
        sd[f"sincnet.norm1d.{i}.weight"] = 1.0 + 0.2 * torch.randn(c, generator=g)
        sd[f"sincnet.norm1d.{i}.bias"] = 0.2 * torch.randn(c, generator=g)
    H = 128
    bound = 1.0 / math.sqrt(H)
    for layer in range(lstm_layers):
        isz = 60 if layer == 0 else 2 * H
        for suffix in ("", "_reverse"):
            sd[f"lstm.weight_ih_l{layer}{suffix}"] = _uniform(g, (4 * H, isz), bound) * ih_gain
            sd[f"lstm.weight_hh_l{layer}{suffix}"] = _uniform(g, (4 * H, H), bound) * hh_gain
            sd[f"lstm.bias_ih_l{layer}{suffix}"] = _uniform(g, (4 * H,), bound)
            sd[f"lstm.bias_hh_l{layer}{suffix}"] = _uniform(g, (4 * H,), bound)
    sd["linear.0.weight"] = _uniform(g, (128, 256), 1.0 / math.sqrt(256)) * lin_gain
    sd["linear.0.bias"] = _uniform(g, (128,), 1.0 / math.sqrt(256))
    sd["linear.1.weight"] = _uniform(g, (128, 128), 1.0 / math.sqrt(128)) * lin_gain
    sd["linear.1.bias"] = _uniform(g, (128,), 1.0 / math.sqrt(128))
    sd["classifier.weight"] = _uniform(g, (num_classes, 128), 1.0 / math.sqrt(128)) * logit_scale
    sd["classifier.bias"] = _uniform(g, (num_classes,), 1.0 / math.sqrt(128))
    if fitted_classifier and seed == 0 and lstm_layers == 4 and num_classes == 7:
        # last layer fitted in closed form on synthetic conversations so that segmentations are not
        # degenerate (generator: tests/golden/make_synthetic_classifier.py)
        import os

        path = os.path.join(os.path.dirname(__file__), "data", "synthetic_classifier_seed0.npz")
        if os.path.exists(path):
            z = np.load(path)
            sd["classifier.weight"] = torch.from_numpy(z["weight"]).clone()
            sd["classifier.bias"] = torch.from_numpy(z["bias"]).clone()
    return sd


# ----------------------------------------------------------------------------------------
# embedding state dict
# ----------------------------------------------------------------------------------------


def _bn(sd, prefix, c, g):
    sd[prefix + ".weight"] = 0.8 + 0.4 * torch.rand(c, generator=g)
    sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(1000, dtype=torch.long)


def _conv(sd, name, cout, cin, k, g, gain=1.1):
    fan_in = cin * k * k
    sd[name] = torch.randn(cout, cin, k, k, generator=g) * (gain / math.sqrt(fan_in))


def make_embedding_state_dict(seed: int = 1, centered: bool = True) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    _conv(sd, "resnet.conv1.weight", 32, 1, 3, g)
    _bn(sd, "resnet.bn1", 32, g)
    in_planes = 32
    for li, (planes, n, stride) in enumerate(((32, 3, 1), (64, 4, 2), (128, 6, 2), (256, 3, 2)), start=1):
        for bi in range(n):
            s = stride if bi == 0 else 1
            p = f"resnet.layer{li}.{bi}"
            _conv(sd, p + ".conv1.weight", planes, in_planes, 3, g)
            _bn(sd, p + ".bn1", planes, g)
            _conv(sd, p + ".conv2.weight", planes, planes, 3, g, gain=0.5)
            _bn(sd, p + ".bn2", planes, g)
            if s != 1 or in_planes != planes:
                _conv(sd, p + ".shortcut.0.weight", planes, in_planes, 1, g, gain=0.8)
                _bn(sd, p + ".shortcut.1", planes, g)
            in_planes = planes
    sd["resnet.seg_1.weight"] = torch.randn(256, 5120, generator=g) / math.sqrt(5120)
    sd["resnet.seg_1.bias"] = 0.01 * torch.randn(256, generator=g)
    if centered and seed == 1:
        # bias calibrated so that embeddings of synthetic conversations are roughly zero-mean (otherwise all
        # cosines are > 0.95 and clustering is trivial); generator: tests/golden/make_synthetic_classifier.py
        import os

        path = os.path.join(os.path.dirname(__file__), "data", "synthetic_embedding_bias_seed1.npz")
        if os.path.exists(path):
            sd["resnet.seg_1.bias"] = torch.from_numpy(np.load(path)["bias"]).clone()
    return sd


# ----------------------------------------------------------------------------------------
# PLDA
# ----------------------------------------------------------------------------------------


def make_plda(seed: int = 2, dim: int = 256, lda_dim: int = 128):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
    lda = q[:, :lda_dim] * (1.0 + 0.1 * rng.standard_normal((1, lda_dim)))
    mean1 = 0.05 * rng.standard_normal(dim)
    mean2 = 0.05 * rng.standard_normal(lda_dim)
    mu = 0.05 * rng.standard_normal(lda_dim)
    tr = np.eye(lda_dim) + 0.05 * rng.standard_normal((lda_dim, lda_dim))
    psi = np.sort(np.exp(rng.uniform(np.log(0.05), np.log(20.0), lda_dim)))[::-1].copy()
    return dict(mean1=mean1, mean2=mean2, lda=lda, mu=mu, tr=tr, psi=psi)


# ----------------------------------------------------------------------------------------
# audio
# ----------------------------------------------------------------------------------------


def make_conversation(duration_s: float, seed: int = 1234, num_speakers: int = 3,
                      sample_rate: int = SAMPLE_RATE, return_turns: bool = False):
    """(1, T) float32 synthetic conversation in [-1, 1] (+ list of (start_s, end_s, speaker) turns)."""
    rng = np.random.default_rng(seed)
    T = int(round(duration_s * sample_rate))
    out = np.zeros(T, dtype=np.float32)
    t_all = np.arange(T, dtype=np.float64) / sample_rate
    spk = []
    for _ in range(num_speakers):
        f0 = rng.uniform(80, 250)
        nh = int(rng.integers(10, 21))
        env = np.exp(-0.5 * ((np.arange(1, nh + 1) * f0 - rng.uniform(300, 2500)) / rng.uniform(400, 1500)) ** 2)
        env = env / env.sum() + 0.02
        spk.append(dict(f0=f0, amps=env, phases=rng.uniform(0, 2 * np.pi, nh), am=rng.uniform(3, 6),
                        vib=rng.uniform(0.002, 0.01)))
    # turn taking: alternate speakers with 0.5-5 s turns, ~8% overlap, some silences
    t = 0.0
    turns = []
    cur = int(rng.integers(num_speakers))
    while t < duration_s:
        turn = float(rng.uniform(0.5, 5.0))
        if rng.uniform() < 0.15:
            t += float(rng.uniform(0.2, 1.5))      # silence
        a, b = t, min(duration_s, t + turn)
        i0, i1 = int(a * sample_rate), int(b * sample_rate)
        if i1 > i0:
            turns.append((a, b, cur))
            s = spk[cur]
            tt = t_all[i0:i1]
            sig = np.zeros(i1 - i0)
            f0 = s["f0"] * (1.0 + s["vib"] * np.sin(2 * np.pi * 5.0 * tt))
            ph = 2 * np.pi * np.cumsum(f0) / sample_rate
            for h, (amp, p0) in enumerate(zip(s["amps"], s["phases"]), start=1):
                if h * s["f0"] < sample_rate / 2 - 200:
                    sig += amp * np.sin(h * ph + p0)
            am = 0.6 + 0.4 * np.sin(2 * np.pi * s["am"] * tt + rng.uniform(0, 2 * np.pi))
            ramp = np.minimum(1.0, np.minimum(np.arange(i1 - i0), np.arange(i1 - i0)[::-1]) / (0.02 * sample_rate))
            out[i0:i1] += (0.35 * sig * am * ramp).astype(np.float32)
        if b >= duration_s:
            break
        overlap = turn * (rng.uniform(0.0, 0.16))
        t = b - overlap
        nxt = int(rng.integers(num_speakers - 1))
        cur = nxt if nxt < cur else nxt + 1
    out += (10 ** (-30 / 20)) * rng.standard_normal(T).astype(np.float32) * 0.3
    peak = np.max(np.abs(out)) + 1e-9
    out = np.clip(out / max(1.0, peak / 0.95), -1.0, 1.0)
    wav = torch.from_numpy(out.astype(np.float32))[None]
    return (wav, turns) if return_turns else wav
