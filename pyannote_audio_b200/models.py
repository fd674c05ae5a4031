"""Model-level seam: ``PyanNet`` and ``WeSpeakerResNet34`` with the reference's state-dict keys and ``forward``
contract, computing through libb200diar.so (no torch ops on the forward path, no CPU fallback).

Reference interfaces mirrored (paths relative to /root/reference/src/pyannote/audio):
  core/model.py:69-183 (Model: specifications, audio, receptive_field, device)
  models/segmentation/PyanNet.py:92-240 (ctor hyper-parameters, num_frames, receptive field, forward)
  models/embedding/wespeaker/__init__.py:41-372 (forward / forward_frames / forward_embedding / dimension)
"""
from __future__ import annotations

import itertools
from functools import cached_property
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .audio import Audio
from .core import Problem, Resolution, SlidingWindow, Specifications

_CONTEXTS: Dict[int, "ops.Context"] = {}
_MODEL_IDS = itertools.count(1)


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
        self.register_buffer("_dummy", torch.zeros(0), persistent=False)      # device tracker, not in state_dict
        # The library context of a device holds ONE set of segmentation and ONE set of embedding weights.  Each model
        # stamps the slot it uploads into with (its unique id, its weight version); a forward re-uploads whenever the
        # slot carries another stamp, so several models on one GPU and load_state_dict() after a forward stay correct.
        self._model_id = next(_MODEL_IDS)
        self._weights_version = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._bump_weights())

    _SLOT = ""          # "seg" | "emb": the context slot this model family uploads into

    def _bump_weights(self):
        self._weights_version += 1

    @property
    def device(self) -> torch.device:
        return self._dummy.device

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._bump_weights()              # weights may have moved / changed dtype: re-upload lazily
        return out

    def _ctx(self) -> "ops.Context":
        ctx = get_context(self.device)
        stamp = (self._model_id, self._weights_version)
        if ctx.owners.get(self._SLOT) != stamp:
            self._upload(ctx)
            ctx.owners[self._SLOT] = stamp
        return ctx

    def _upload(self, ctx):
        raise NotImplementedError

    # ---- checkpoints (core/model.py:497-655 without lightning / the HF hub) -----------------------------------
    @classmethod
    def from_pretrained(cls, checkpoint, map_location=None, strict: bool = True, subfolder: Optional[str] = None,
                        revision: Optional[str] = None, token=None, cache_dir=None, **kwargs) -> "Model":
        """Load a pyannote.audio (Lightning-format) checkpoint: a local ``pytorch_model.bin``, a directory holding
        one (optionally under ``subfolder``), or an ``io.BytesIO``.  The file is read with plain ``torch.load``;
        the pickled ``pyannote.audio.core.task`` objects (Specifications / Problem / Resolution) are mapped onto this
        package's mirrors, so neither lightning nor pyannote.audio needs to be importable.  The checkpoint names its
        own architecture (``checkpoint["pyannote.audio"]["architecture"]``); ``kwargs`` override saved
        hyper-parameters like the reference.  Hub identifiers cannot be resolved offline and raise."""
        import io
        import os
        from pathlib import Path

        if isinstance(checkpoint, io.BytesIO) or os.path.isfile(checkpoint):
            if revision is not None:
                raise ValueError("Revisions cannot be used with local checkpoints.")
            path = checkpoint
        elif os.path.isdir(checkpoint):
            if revision is not None:
                raise ValueError("Revisions cannot be used with local checkpoints.")
            path = Path(checkpoint) / subfolder / "pytorch_model.bin" if subfolder else \
                Path(checkpoint) / "pytorch_model.bin"
        else:
            if "@" in str(checkpoint):
                raise ValueError("Revisions must be passed with `revision` keyword argument.")
            raise ValueError(f"'{checkpoint}' is not a local checkpoint; Hugging Face hub identifiers cannot be "
                             f"downloaded here (no network): pass the path of a downloaded pytorch_model.bin")
        if map_location is None:
            map_location = "cpu"
        loaded = torch.load(path, map_location=map_location, weights_only=False, pickle_module=_checkpoint_pickle)
        meta = loaded["pyannote.audio"]
        class_name = meta["architecture"]["class"]
        klass = {"PyanNet": PyanNet, "WeSpeakerResNet34": WeSpeakerResNet34}.get(class_name)
        if klass is None:
            raise NotImplementedError(f"architecture {meta['architecture']['module']}.{class_name} has no sm_100a "
                                      f"implementation (community-1 uses PyanNet and WeSpeakerResNet34)")
        if cls not in (Model, klass) and not issubclass(klass, cls):
            raise ValueError(f"checkpoint holds a {class_name}, not a {cls.__name__}")
        hparams = dict(loaded.get("hyper_parameters", {}))
        hparams.update(kwargs)
        hparams = {k: v for k, v in hparams.items() if k in klass._HPARAMS}
        model = klass(**hparams)
        specs = meta.get("specifications", None)
        if specs is not None:
            if isinstance(specs, (tuple, list)):
                raise NotImplementedError("multi-task checkpoints are not supported")
            model.specifications = specs
        sd = loaded["state_dict"]
        own = set(model.state_dict().keys())
        if not strict:
            sd = {k: v for k, v in sd.items() if k in own}
        model.load_state_dict(sd, strict=strict)
        model.eval()
        return model

    @cached_property
    def receptive_field(self) -> SlidingWindow:
        size = self.receptive_field_size(num_frames=1)
        step = self.receptive_field_size(num_frames=2) - size
        start = self.receptive_field_center(frame=0) - (size - 1) / 2
        sr = self.hparams.sample_rate
        return SlidingWindow(start=start / sr, duration=size / sr, step=step / sr)


class _CheckpointUnpickler(__import__("pickle").Unpickler):
    """Resolves the reference's pickled task types to this package's mirrors (core.py)."""

    _MAP = {("pyannote.audio.core.task", "Specifications"): Specifications,
            ("pyannote.audio.core.task", "Problem"): Problem,
            ("pyannote.audio.core.task", "Resolution"): Resolution}

    def find_class(self, module, name):
        hit = self._MAP.get((module, name))
        if hit is not None:
            return hit
        if module.split(".")[0] in ("lightning", "pytorch_lightning", "lightning_fabric"):
            # e.g. AttributeDict for hyper_parameters: a plain dict subclass is enough
            return dict
        return super().find_class(module, name)


class _checkpoint_pickle:
    """``pickle_module`` for torch.load: the stdlib pickle with the class mapping above."""

    import pickle as _p

    Unpickler = _CheckpointUnpickler
    Pickler = _p.Pickler
    load = staticmethod(lambda f, **kw: _CheckpointUnpickler(f, **kw).load())
    loads = staticmethod(_p.loads)
    dump = staticmethod(_p.dump)
    dumps = staticmethod(_p.dumps)
    HIGHEST_PROTOCOL = _p.HIGHEST_PROTOCOL
    __name__ = "pickle"


def _mel_sinc_init(n_filters=80, sample_rate=16000.0, min_low_hz=50, min_band_hz=50):
    def to_mel(hz):
        return 2595 * np.log10(1 + hz / 700)

    def to_hz(mel):
        return 700 * (10 ** (mel / 2595) - 1)

    mel = np.linspace(to_mel(30), to_mel(sample_rate / 2 - (min_low_hz + min_band_hz)),
                      n_filters // 2 + 1, dtype="float32")
    hz = to_hz(mel)
    return torch.from_numpy(hz[:-1]).view(-1, 1), torch.from_numpy(np.diff(hz)).view(-1, 1)


def sinc_buffers(kernel_size=251, sample_rate=16000.0):
    half = kernel_size // 2
    window_ = torch.from_numpy(np.hamming(kernel_size)[:half]).float()
    n_ = 2 * np.pi * (torch.arange(-half, 0.0).view(1, -1) / sample_rate)
    return window_, n_


class _ParamSincFB(nn.Module):
    def __init__(self):
        super().__init__()
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

    _SLOT = "seg"
    _HPARAMS = ("sincnet", "lstm", "linear", "sample_rate", "num_channels")
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

    def forward_chunks(self, wav: torch.Tensor, chunk_off, chunk_valid, return_logp: bool = False, out=None):
        """Hot-path entry: chunks addressed inside one resident device waveform (no unfold copy)."""
        return self._ctx().seg_forward(wav, chunk_off, chunk_valid, return_logp=return_logp, out=out)

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
    _SLOT = "emb"
    _HPARAMS = ("sample_rate", "num_channels", "num_mel_bins", "frame_length", "frame_shift", "dither",
                "window_type", "use_energy")

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

    def forward_chunks(self, wav: torch.Tensor, chunk_off, chunk_valid, masks: torch.Tensor, out=None,
                       peers=None) -> torch.Tensor:
        """Hot-path entry: (num_chunks, 3, 589) uint8 masks -> (num_chunks, 3, 256) embeddings, one trunk pass."""
        return self._ctx().emb_forward(wav, chunk_off, chunk_valid, masks, out=out, peers=peers)

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
