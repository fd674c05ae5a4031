"""Audio I/O boundary (mirror of /root/reference/src/pyannote/audio/core/io.py:110-484, in-memory branch first).

The hot path works on float32 mono 16 kHz waveforms.  ``AudioFile`` may be a mapping
``{"waveform": (channel, time) tensor, "sample_rate": int}`` (io.py:59-71, 328-332, 384-414), or a path to a PCM WAV
file (read with the stdlib / scipy -- torchcodec is not part of this image).
"""
from __future__ import annotations

from io import IOBase
from pathlib import Path
from typing import Mapping, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .core import Segment

AudioFile = Union[str, Path, IOBase, Mapping]


class Audio:
    def __init__(self, sample_rate: Optional[int] = None, mono: Optional[str] = None):
        self.sample_rate = sample_rate
        self.mono = mono

    @staticmethod
    def validate_file(file: AudioFile) -> Mapping:
        """io.py:151-214: mapping / path / file-like object -> validated mapping (same messages)."""
        if isinstance(file, Mapping):
            pass
        elif isinstance(file, (str, Path)):
            file = {"audio": str(file), "uri": Path(file).stem}
        elif isinstance(file, IOBase):
            return {"audio": file, "uri": "stream"}
        else:
            raise ValueError("AudioFile must be a path, a file-like object, or a mapping with an 'audio' or "
                             "'waveform' (+ 'sample_rate') key.")
        if "waveform" in file:
            waveform = file["waveform"]
            if len(waveform.shape) != 2 or waveform.shape[0] > waveform.shape[1]:
                raise ValueError("'waveform' must be provided as a (channel, time) torch Tensor.")
            if file.get("sample_rate", None) is None:
                raise ValueError("'waveform' must be provided with their 'sample_rate'.")
            file.setdefault("uri", "waveform")           # in place like the reference (io.py:193): hooks store
        elif "audio" in file:                            # their artifacts in the caller's mapping
            if isinstance(file["audio"], IOBase):
                return file
            path = Path(file["audio"])
            if not path.is_file():
                raise ValueError(f"File {path} does not exist")
            file.setdefault("uri", path.stem)
        else:
            raise ValueError("Neither 'waveform' nor 'audio' is available for this file.")
        return file

    def get_num_samples(self, duration: float, sample_rate: Optional[int] = None) -> int:
        sample_rate = sample_rate or self.sample_rate
        if sample_rate is None:
            raise ValueError("`sample_rate` must be provided to compute number of samples.")
        return round(duration * sample_rate)

    def downmix_and_resample(self, waveform: torch.Tensor, sample_rate: int, channel: Optional[int] = None):
        if channel is not None:
            waveform = waveform[channel: channel + 1]
        num_channels = waveform.shape[0]
        if num_channels > 1:
            if self.mono == "random":
                channel = np.random.randint(num_channels)
                waveform = waveform[channel: channel + 1]
            elif self.mono == "downmix":
                waveform = waveform.mean(dim=0, keepdim=True)
        if (self.sample_rate is not None) and (self.sample_rate != sample_rate):
            import torchaudio.functional as AF

            waveform = AF.resample(waveform, sample_rate, self.sample_rate)
            sample_rate = self.sample_rate
        return waveform, sample_rate

    @staticmethod
    def _read_wav(path) -> Tuple[torch.Tensor, int]:
        from scipy.io import wavfile

        if isinstance(path, IOBase):                     # file-like objects are read from their start (io.py:337-338)
            path.seek(0)
        sr, data = wavfile.read(path)
        if data.ndim == 1:
            data = data[:, None]
        if data.dtype == np.int16:
            data = data.astype(np.float32) / 32768.0
        elif data.dtype == np.int32:
            data = data.astype(np.float32) / 2147483648.0
        elif data.dtype == np.uint8:
            data = (data.astype(np.float32) - 128.0) / 128.0
        else:
            data = data.astype(np.float32)
        return torch.from_numpy(np.ascontiguousarray(data.T)), int(sr)

    # ---- device ingest (b200_audio_ingest): raw PCM crosses PCIe once, downmix + resample run on the GPU ---------
    @staticmethod
    def _read_pcm(path: str):
        """Raw samples of a WAV file without conversion: int16 (frames, channels) when the file is 16-bit PCM (the
        common case: half the bytes of float32), otherwise float32 (channels, frames) converted on the host."""
        from scipy.io import wavfile

        if isinstance(path, IOBase):
            path.seek(0)
        sr, data = wavfile.read(path)
        if data.ndim == 1:
            data = data[:, None]
        if data.dtype == np.int16:
            return torch.from_numpy(np.ascontiguousarray(data)), int(sr)
        w, sr = Audio._read_wav(path)
        return w.contiguous(), sr

    def raw(self, file: AudioFile):
        """(raw tensor, sample_rate, channel): int16 (frames, channels) or float32 (channels, frames), on the host."""
        file = self.validate_file(file)
        if "waveform" in file:
            w = file["waveform"]
            return (w if w.dtype == torch.float32 else w.float()), int(file["sample_rate"]), file.get("channel", None)
        pcm, sr = self._read_pcm(file["audio"])
        return pcm, sr, file.get("channel", None)

    def num_samples_out(self, raw: torch.Tensor, sample_rate: int) -> int:
        """Length of the mono waveform the device ingest will produce (ceil(new * len / orig), like torchaudio)."""
        from . import _lib

        frames = raw.shape[0] if raw.dtype == torch.int16 else raw.shape[1]
        target = self.sample_rate or sample_rate
        return int(_lib.load().b200_audio_num_frames(int(frames), int(sample_rate), int(target)))

    def needs_ingest(self, raw: torch.Tensor, sample_rate: int, channel=None) -> bool:
        mono = raw.dtype == torch.float32 and raw.shape[0] == 1
        return not (mono and (self.sample_rate in (None, sample_rate)))

    def ingest(self, ctx, raw: torch.Tensor, sample_rate: int, channel=None, out: Optional[torch.Tensor] = None):
        """H2D of the raw samples + downmix / resample on the device -> float32 mono (samples,) device tensor."""
        if self.mono == "random" and channel is None and (raw.shape[1] if raw.dtype == torch.int16 else raw.shape[0]) > 1:
            nch = raw.shape[1] if raw.dtype == torch.int16 else raw.shape[0]
            channel = int(np.random.randint(nch))
        dev = raw.contiguous().to(ctx.device, non_blocking=True)
        return ctx.audio_ingest(dev, sample_rate, self.sample_rate or sample_rate, channel=channel,
                                downmix=(self.mono == "downmix") or channel is None, out=out)

    def get_duration(self, file: AudioFile) -> float:
        file = self.validate_file(file)
        if "waveform" in file:
            return file["waveform"].shape[1] / file["sample_rate"]
        w, sr = self._read_wav(file["audio"])
        return w.shape[1] / sr

    def __call__(self, file: AudioFile) -> Tuple[torch.Tensor, int]:
        file = self.validate_file(file)
        if "waveform" in file:
            waveform, sample_rate = file["waveform"], file["sample_rate"]
        else:
            waveform, sample_rate = self._read_wav(file["audio"])
        return self.downmix_and_resample(waveform, sample_rate, channel=file.get("channel", None))

    def crop(self, file: AudioFile, segment: Segment, mode: str = "raise") -> Tuple[torch.Tensor, int]:
        file = self.validate_file(file)
        if "waveform" in file:
            waveform, sample_rate = file["waveform"], file["sample_rate"]
        else:
            waveform, sample_rate = self._read_wav(file["audio"])
        _, num_samples = waveform.shape
        duration = num_samples / sample_rate
        start_sample = self.get_num_samples(segment.start, sample_rate)
        pad_start = max(0, -start_sample)
        if start_sample < 0:
            if mode == "raise":
                raise ValueError(f"requested chunk with negative start time (t={segment.start:.3f}s)")
            start_sample = 0
        end_sample = self.get_num_samples(segment.end, sample_rate)
        pad_end = max(end_sample, num_samples) - num_samples
        if end_sample >= num_samples:
            if mode == "raise":
                raise ValueError(f"requested chunk with end time (t={segment.end:.3f}s) greater than "
                                 f"{file.get('uri', 'in-memory')} file duration ({duration:.3f}s).")
            end_sample = num_samples
        data = F.pad(waveform[:, start_sample:end_sample], (pad_start, pad_end))
        return self.downmix_and_resample(data, sample_rate, channel=file.get("channel", None))
