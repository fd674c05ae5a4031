"""Model-level seam: ``PyanNet`` and ``WeSpeakerResNet34`` with the reference's state-dict keys and ``forward``
contract, computing through libb200diar.so (no torch ops on the forward path, no CPU fallback).

Reference interfaces mirrored (paths relative to /root/reference/src/pyannote/audio):
  core/model.py:69-183 (Model: specifications, audio, receptive_field, device)
  models/segmentation/PyanNet.py:92-240 (ctor hyper-parameters, num_frames, receptive field, forward)
  models/embedding/wespeaker/__init__.py:41-372 (forward / forward_frames / forward_embedding / dimension)
"""
from __future__ import annotations

from functools import cached_property
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .audio import Audio
from .core import Problem, Resolution, SlidingWindow, Specifications

_CONTEXTS: Dict[int, "ops.Context"] = {}


def get_context(device) -> "ops.Context":
    """One library context per CUDA device, shared by all models placed on it."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(
            f"pyannote_audio_b200 models only run on CUDA (B200 / sm_100a) devices, not on '{device}': move the "
            f"model with .to(torch.device('cuda'))")
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index not in _CONTEXTS:
        _CONTEXTS[index] = ops.Context(torch.device("cuda", index))
    return _CONTEXTS[index]


def _conv1d_num_frames(n, k, s, p=0, d=1):
    return 1 + (n + 2 * p - d * (k - 1) - 1) // s


class Model(nn.Module):
    """Subset of pyannote.audio.core.model.Model that inference relies on."""

    def __init__(self, sample_rate: int = 16000, num_channels: int = 1):
        super().__init__()
        self.hparams = type("HParams", (), {})()
        self.hparams.sample_rate = sample_rate
        self.hparams.num_channels = num_channels
        self.audio = Audio(sample_rate=sample_rate, mono="downmix")
        self._dummy = nn.Parameter(torch.zeros(0), requires_grad=False)
        self._uploaded_to: Optional[int] = None

    @property
    def device(self) -> torch.device:
        return self._dummy.device

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._uploaded_to = None          # weights may have moved: re-upload lazily
        return out

    def _ctx(self) -> "ops.Context":
        ctx = get_context(self.device)
        if self._uploaded_to != ctx.device.index:
            self._upload(ctx)
            self._uploaded_to = ctx.device.index
        return ctx

    def _upload(self, ctx):
        raise NotImplementedError

    @cached_property
    def receptive_field(self) -> SlidingWindow:
        size = self.receptive_field_size(num_frames=1)
        step = self.receptive_field_size(num_frames=2) - size
        start = self.receptive_field_center(frame=0) - (size - 1) / 2
        sr = self.hparams.sample_rate
        return SlidingWindow(start=start / sr, duration=size / sr, step=step / sr)


class _ParamSincFB(nn.Module):
    def __init__(self):
        super().__init__()
        from .synthetic import _mel_sinc_init, sinc_buffers

        low, band = _mel_sinc_init()
        self.low_hz_ = nn.Parameter(low, requires_grad=False)
        self.band_hz_ = nn.Parameter(band, requires_grad=False)
        window_, n_ = sinc_buffers()
        self.register_buffer("window_", window_)
        self.register_buffer("n_", n_)


class _Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.filterbank = _ParamSincFB()


class _SincNetParams(nn.Module):
    """Parameter container with the key names of models/blocks/sincnet.py:41-79."""

    def __init__(self):
        super().__init__()
        self.wav_norm1d = nn.InstanceNorm1d(1, affine=True)
        self.conv1d = nn.ModuleList([_Encoder(), nn.Conv1d(80, 60, 5), nn.Conv1d(60, 60, 5)])
        self.norm1d = nn.ModuleList([nn.InstanceNorm1d(80, affine=True), nn.InstanceNorm1d(60, affine=True),
                                     nn.InstanceNorm1d(60, affine=True)])


class PyanNet(Model):
    """SincNet > LSTM > Feed forward > Classifier, community-1 shape (4 BiLSTM layers of 128, 2x128 linear)."""

    KERNEL = [251, 3, 5, 3, 5, 3]
    STRIDE = [10, 3, 1, 3, 1, 3]

    def __init__(self, sincnet: Optional[dict] = None, lstm: Optional[dict] = None, linear: Optional[dict] = None,
                 sample_rate: int = 16000, num_channels: int = 1, duration: float = 10.0):
        super().__init__(sample_rate=sample_rate, num_channels=num_channels)
        if sample_rate != 16000:
            raise NotImplementedError("SincNet only supports 16kHz audio for now.")
        lstm_hp = {"hidden_size": 128, "num_layers": 4, "bidirectional": True, "monolithic": True, "dropout": 0.0}
        lstm_hp.update(lstm or {})
        linear_hp = {"hidden_size": 128, "num_layers": 2}
        linear_hp.update(linear or {})
        sinc_hp = {"stride": 10}
        sinc_hp.update(sincnet or {})
        if (lstm_hp["hidden_size"], lstm_hp["bidirectional"], lstm_hp["monolithic"]) != (128, True, True) or \
                not (1 <= lstm_hp["num_layers"] <= 4) or (linear_hp["hidden_size"], linear_hp["num_layers"]) != (128, 2) \
                or sinc_hp["stride"] != 10:
            raise NotImplementedError("the sm_100a kernels implement the community-1 PyanNet shape only: "
                                      "SincNet stride 10, 1-4 bidirectional LSTM layers of 128, 2 linear layers of 128")
        self.hparams.sincnet, self.hparams.lstm, self.hparams.linear = sinc_hp, lstm_hp, linear_hp
        self.sincnet = _SincNetParams()
        self.lstm = nn.LSTM(60, hidden_size=128, num_layers=lstm_hp["num_layers"], bidirectional=True,
                            batch_first=True)
        self.linear = nn.ModuleList([nn.Linear(256, 128), nn.Linear(128, 128)])
        self.classifier = nn.Linear(128, 7)
        self.specifications = Specifications(problem=Problem.MONO_LABEL_CLASSIFICATION, resolution=Resolution.FRAME,
                                             duration=duration, warm_up=(0.0, 0.0),
                                             classes=["speaker#1", "speaker#2", "speaker#3"], powerset_max_classes=2,
                                             permutation_invariant=True)
        for p in self.parameters():
            p.requires_grad_(False)

    @property
    def dimension(self) -> int:
        return self.specifications.num_powerset_classes

    def num_frames(self, num_samples: int) -> int:
        n = num_samples
        for k, s in zip(self.KERNEL, self.STRIDE):
            n = _conv1d_num_frames(n, k, s)
        return n

    def receptive_field_size(self, num_frames: int = 1) -> int:
        rf = num_frames
        for k, s in reversed(list(zip(self.KERNEL, self.STRIDE))):
            rf = 1 + (k - 1) + (rf - 1) * s
        return rf

    def receptive_field_center(self, frame: int = 0) -> int:
        c = frame
        for k, s in reversed(list(zip(self.KERNEL, self.STRIDE))):
            c = c * s + (k - 1) // 2
        return c

    def _upload(self, ctx):
        ctx.load_segmentation(self.state_dict())

    def forward_chunks(self, wav: torch.Tensor, chunk_off, chunk_valid, return_logp: bool = False):
        """Hot-path entry: chunks addressed inside one resident device waveform (no unfold copy)."""
        return self._ctx().seg_forward(wav, chunk_off, chunk_valid, return_logp=return_logp)

    def forward(self, waveforms: torch.Tensor) -> torch.Tensor:
        """waveforms (batch, channel, sample) -> log-probabilities (batch, 589, 7)."""
        b, c, s = waveforms.shape
        if c != 1 or s != ops.CHUNK:
            raise ValueError(f"PyanNet kernels expect mono {ops.CHUNK}-sample (10 s @ 16 kHz) chunks, got {c}x{s}")
        ctx = self._ctx()
        flat = waveforms.to(device=ctx.device, dtype=torch.float32).reshape(-1).contiguous()
        off = np.arange(b, dtype=np.int64) * s
        valid = np.full(b, s, dtype=np.int32)
        _, logp = ctx.seg_forward(flat, off, valid, return_logp=True)
        return logp


class _BasicBlockParams(nn.Module):
    def __init__(self, in_planes, planes, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes:
            self.shortcut = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride, bias=False),
                                          nn.BatchNorm2d(planes))


class _ResNet34Params(nn.Module):
    """Parameter container with the key names of models/embedding/wespeaker/resnet.py:233-252."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 32, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        in_planes = 32
        for li, (planes, n, stride) in enumerate(((32, 3, 1), (64, 4, 2), (128, 6, 2), (256, 3, 2)), start=1):
            blocks = []
            for s in [stride] + [1] * (n - 1):
                blocks.append(_BasicBlockParams(in_planes, planes, s))
                in_planes = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.seg_1 = nn.Linear(5120, 256)


class WeSpeakerResNet34(Model):
    def __init__(self, sample_rate: int = 16000, num_channels: int = 1, num_mel_bins: int = 80,
                 frame_length: int = 25, frame_shift: int = 10, dither: float = 0.0, window_type: str = "hamming",
                 use_energy: bool = False):
        super().__init__(sample_rate=sample_rate, num_channels=num_channels)
        if (sample_rate, num_mel_bins, frame_length, frame_shift, dither, window_type, use_energy) != \
                (16000, 80, 25, 10, 0.0, "hamming", False):
            raise NotImplementedError("the sm_100a fbank kernel implements the community-1 configuration only "
                                      "(16 kHz, 80 mel bins, 25/10 ms hamming frames, no dither, no energy)")
        self.resnet = _ResNet34Params()
        self.specifications = Specifications(problem=Problem.REPRESENTATION, resolution=Resolution.CHUNK, duration=10.0)
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    @property
    def dimension(self) -> int:
        return 256

    def _upload(self, ctx):
        ctx.load_embedding(self.state_dict())

    def num_frames(self, num_samples: int) -> int:
        n = _conv1d_num_frames(num_samples, 400, 160)
        for s in (1, 2, 2, 2):
            n = _conv1d_num_frames(n, 3, s, p=1)
        return n

    def forward_chunks(self, wav: torch.Tensor, chunk_off, chunk_valid, masks: torch.Tensor) -> torch.Tensor:
        """Hot-path entry: (num_chunks, 3, 589) uint8 masks -> (num_chunks, 3, 256) embeddings, one trunk pass."""
        return self._ctx().emb_forward(wav, chunk_off, chunk_valid, masks)

    def _flat(self, waveforms):
        b, c, s = waveforms.shape
        if c != 1 or s != ops.CHUNK:
            raise ValueError(f"WeSpeaker kernels expect mono {ops.CHUNK}-sample (10 s @ 16 kHz) chunks, got {c}x{s}")
        ctx = self._ctx()
        flat = waveforms.to(device=ctx.device, dtype=torch.float32).reshape(-1).contiguous()
        return ctx, flat, np.arange(b, dtype=np.int64) * s, np.full(b, s, dtype=np.int32)

    def compute_fbank(self, waveforms: torch.Tensor) -> torch.Tensor:
        ctx, flat, off, valid = self._flat(waveforms)
        return ctx.emb_fbank(flat, off, valid)

    def forward_frames(self, waveforms: torch.Tensor) -> torch.Tensor:
        return self._ctx().emb_trunk(self.compute_fbank(waveforms))

    def forward(self, waveforms: torch.Tensor, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        """waveforms (batch, 1, 160000), weights (batch, 589) or (batch, speakers<=3, 589) in {0,1}."""
        ctx, flat, off, valid = self._flat(waveforms)
        b = len(off)
        if weights is None:
            w = torch.ones((b, 1, ops.FRAMES), dtype=torch.float32)
            squeeze = True
        else:
            squeeze = weights.dim() == 2
            w = weights.unsqueeze(1) if squeeze else weights
        if w.shape[-1] != ops.FRAMES or w.shape[1] > 3:
            raise ValueError("weights must have 589 frames and at most 3 speakers")
        if not bool(((w == 0) | (w == 1)).all()):
            raise ValueError("the masked statistics pooling kernel takes binary (0/1) weights")
        masks = torch.zeros((b, 3, ops.FRAMES), dtype=torch.uint8, device=ctx.device)
        masks[:, : w.shape[1]] = w.to(ctx.device).to(torch.uint8)
        emb = ctx.emb_forward(flat, off, valid, masks)[:, : w.shape[1]]
        return emb[:, 0] if squeeze else emb
