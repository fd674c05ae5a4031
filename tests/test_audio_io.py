"""CPU: host mirror of Audio (pyannote_audio_b200/audio.py) against the reference's own I/O tests
(/root/reference/tests/io_test.py:10-100), on a synthetic WAV written by the test (the reference's dev00.wav is not
copied): resampling on load, defaults, channel selection, in-memory waveforms, crop by Segment, file-like objects."""
import io

import numpy as np
import pytest
import torch
from scipy.io import wavfile

from pyannote_audio_b200.audio import Audio
from pyannote_audio_b200.core import Segment


@pytest.fixture(scope="module")
def wav_file(tmp_path_factory):
    path = tmp_path_factory.mktemp("audio") / "dev00.wav"
    rng = np.random.default_rng(0)
    t = np.arange(3 * 16000) / 16000.0
    stereo = np.stack([0.4 * np.sin(2 * np.pi * 220 * t), 0.3 * np.sin(2 * np.pi * 330 * t)], axis=1)
    stereo += 0.01 * rng.standard_normal(stereo.shape)
    wavfile.write(str(path), 16000, (stereo * 32767).astype(np.int16))
    return str(path)


def test_audio_resample(wav_file):
    # io_test.py:10-18
    loader = Audio(sample_rate=8000, mono="downmix")
    wav, sr = loader(wav_file)
    assert isinstance(wav, torch.Tensor) and sr == 8000
    assert wav.shape == (1, 3 * 8000)


def test_basic_load_with_defaults(wav_file):
    # io_test.py:21-25
    wav, sr = Audio(mono="downmix")(wav_file)
    assert isinstance(wav, torch.Tensor) and wav.shape == (1, 3 * 16000) and sr == 16000
    ref = wavfile.read(wav_file)[1].astype(np.float32) / 32768.0
    np.testing.assert_allclose(wav[0].numpy(), ref.mean(axis=1), atol=1e-7)


def test_correct_audio_channel():
    # io_test.py:28-34
    waveform = torch.rand(2, 16000 * 2)
    wav, sr = Audio(mono="downmix")({"waveform": waveform, "sample_rate": 16000, "channel": 1})
    assert torch.equal(wav, waveform[1:2]) and sr == 16000


def test_can_load_with_waveform():
    # io_test.py:37-43
    waveform = torch.rand(2, 16000 * 2)
    wav, sr = Audio(mono="downmix")({"waveform": waveform, "sample_rate": 16000})
    assert isinstance(wav, torch.Tensor) and wav.shape == (1, 32000) and sr == 16000
    assert torch.allclose(wav, waveform.mean(dim=0, keepdim=True))


def test_can_crop(wav_file):
    # io_test.py:46-52
    wav, sr = Audio(mono="downmix").crop(wav_file, Segment(0.2, 0.7))
    assert wav.shape[1] / sr == 0.5


def test_can_crop_waveform():
    # io_test.py:55-63
    waveform = torch.rand(1, 16000 * 2)
    wav, sr = Audio(mono="downmix").crop({"waveform": waveform, "sample_rate": 16000}, Segment(0.2, 0.7))
    assert isinstance(wav, torch.Tensor) and sr == 16000
    assert torch.equal(wav, waveform[:, 3200:11200])


def test_can_load_from_file_like(wav_file):
    # io_test.py:66-74
    loader = Audio(mono="downmix")
    with open(wav_file, "rb") as f:
        wav, sr = loader(f)
        again, _ = loader(f)                               # file-like objects are rewound before every read
    assert isinstance(wav, torch.Tensor) and sr == 16000 and torch.equal(wav, again)
    assert torch.equal(wav, loader(wav_file)[0])



def test_can_crop_from_file_like(wav_file):
    # io_test.py:77-88
    loader = Audio(mono="downmix")
    with open(wav_file, "rb") as f:
        wav, sr = loader.crop(f, Segment(0.2, 0.7))
    assert isinstance(wav, torch.Tensor) and sr == 16000 and wav.shape[1] == 0.5 * 16000
    with open(wav_file, "rb") as f:
        assert loader.get_duration({"audio": f}) == 3.0


def test_validate_file_errors(tmp_path):
    # io.py:151-214: the error behaviour of the boundary
    with pytest.raises(ValueError, match="does not exist"):
        Audio.validate_file(str(tmp_path / "missing.wav"))
    with pytest.raises(ValueError, match="channel, time"):
        Audio.validate_file({"waveform": torch.zeros(100, 2), "sample_rate": 16000})
    with pytest.raises(ValueError, match="sample_rate"):
        Audio.validate_file({"waveform": torch.zeros(1, 100)})
    with pytest.raises(ValueError, match="Neither"):
        Audio.validate_file({"uri": "x"})
    with pytest.raises(ValueError):
        Audio.validate_file(42)
    assert Audio.validate_file(io.BytesIO(b""))["uri"] == "stream"
    with pytest.raises(ValueError, match="negative start"):
        Audio().crop({"waveform": torch.zeros(1, 16000), "sample_rate": 16000}, Segment(-0.5, 0.5))
    wav, _ = Audio().crop({"waveform": torch.ones(1, 16000), "sample_rate": 16000}, Segment(-0.5, 1.5), mode="pad")
    assert wav.shape == (1, 32000) and float(wav[0, :8000].abs().sum()) == 0.0 and float(wav[0, 8000:24000].sum()) == 16000.0
