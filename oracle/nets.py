"""Oracle (TEST INFRASTRUCTURE): CPU fp32 restatement of the two networks on the hot path.

Follows (paths relative to /root/reference/src/pyannote/audio):

* ``SincNet``            models/blocks/sincnet.py:41-79 (ctor), :163-184 (forward)
* ``ParamSincFB``        asteroid-filterbanks 0.4.0 (NOT in the tree, call site
                         models/blocks/sincnet.py:31,58-69) -- restated, UNPINNED
* ``PyanNet``            models/segmentation/PyanNet.py:64-72, 92-139, 152-161, 211-240
                         + ``default_activation`` core/model.py:284-300 (LogSoftmax)
* ``WeSpeakerResNet34``  models/embedding/wespeaker/__init__.py:113-139, 324-372
* ``ResNet`` / ``BasicBlock`` / ``TSTP``  models/embedding/wespeaker/resnet.py:36-66,84-145,214-252,399-430
* ``StatsPool``          models/blocks/pooling.py:30-61, 76-130
* ``Powerset``           utils/powerset.py:80-109, 115-140

State-dict key names equal the reference's so that real checkpoints would load.

Pinning: ResNet / StatsPool / Powerset against the reference's leaf files (tests/golden/make_golden.py); PyanNet and
WeSpeakerResNet34 (compute_fbank, forward with weights) against the reference's own PyanNet.py / sincnet.py /
wespeaker/__init__.py executed by path with these very state dicts loaded strictly (tests/golden/make_golden_apply.py,
tests/test_oracle_apply_golden.py).  ParamSincFB / Encoder stay unpinned (asteroid-filterbanks is not available).
"""

from __future__ import annotations

from itertools import combinations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# SincNet front-end
# ----------------------------------------------------------------------------------------


def _to_mel(hz):
    return 2595 * np.log10(1 + hz / 700)


def _to_hz(mel):
    return 700 * (10 ** (mel / 2595) - 1)


class ParamSincFB(nn.Module):
    """Restatement of asteroid_filterbanks.ParamSincFB (0.4.0) -- parity unpinned."""

    def __init__(self, n_filters=80, kernel_size=251, stride=10, sample_rate=16000.0,
                 min_low_hz=50, min_band_hz=50):
        super().__init__()
        self.n_filters, self.kernel_size, self.stride = n_filters, kernel_size, stride
        self.sample_rate = sample_rate
        self.min_low_hz, self.min_band_hz = min_low_hz, min_band_hz
        self.half_kernel = kernel_size // 2
        # mel-spaced initialisation
        low_hz = 30
        high_hz = sample_rate / 2 - (min_low_hz + min_band_hz)
        mel = np.linspace(_to_mel(low_hz), _to_mel(high_hz), n_filters // 2 + 1, dtype="float32")
        hz = _to_hz(mel)
        self.low_hz_ = nn.Parameter(torch.from_numpy(hz[:-1]).view(-1, 1))
        self.band_hz_ = nn.Parameter(torch.from_numpy(np.diff(hz)).view(-1, 1))
        window_ = np.hamming(kernel_size)[: self.half_kernel]
        n_ = 2 * np.pi * (torch.arange(-self.half_kernel, 0.0).view(1, -1) / sample_rate)
        self.register_buffer("window_", torch.from_numpy(window_).float())
        self.register_buffer("n_", n_)

    def _make(self, low, high, kind):
        band = (high - low)[:, 0]
        ft_low = torch.matmul(low, self.n_)
        ft_high = torch.matmul(high, self.n_)
        if kind == "cos":
            left = ((torch.sin(ft_high) - torch.sin(ft_low)) / (self.n_ / 2)) * self.window_
            center = 2 * band.view(-1, 1)
            right = torch.flip(left, dims=[1])
        else:
            left = ((torch.cos(ft_low) - torch.cos(ft_high)) / (self.n_ / 2)) * self.window_
            center = torch.zeros_like(band.view(-1, 1))
            right = -torch.flip(left, dims=[1])
        bp = torch.cat([left, center, right], dim=1)
        bp = bp / (2 * band[:, None])
        return bp.view(self.n_filters // 2, 1, self.kernel_size)

    def filters(self):
        low = self.min_low_hz + torch.abs(self.low_hz_)
        high = torch.clamp(low + self.min_band_hz + torch.abs(self.band_hz_),
                           self.min_low_hz, self.sample_rate / 2)
        return torch.cat([self._make(low, high, "cos"), self._make(low, high, "sin")], dim=0)


class Encoder(nn.Module):
    """asteroid_filterbanks.Encoder: conv1d with the filter bank, no bias, no padding."""

    def __init__(self, filterbank):
        super().__init__()
        self.filterbank = filterbank

    def forward(self, x):
        return F.conv1d(x, self.filterbank.filters(), stride=self.filterbank.stride, padding=0)


class SincNet(nn.Module):
    def __init__(self, sample_rate=16000, stride=10):
        super().__init__()
        self.stride = stride
        self.wav_norm1d = nn.InstanceNorm1d(1, affine=True)
        self.conv1d = nn.ModuleList()
        self.pool1d = nn.ModuleList()
        self.norm1d = nn.ModuleList()
        self.conv1d.append(Encoder(ParamSincFB(80, 251, stride=stride, sample_rate=sample_rate,
                                               min_low_hz=50, min_band_hz=50)))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(80, affine=True))
        self.conv1d.append(nn.Conv1d(80, 60, 5, stride=1))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(60, affine=True))
        self.conv1d.append(nn.Conv1d(60, 60, 5, stride=1))
        self.pool1d.append(nn.MaxPool1d(3, stride=3, padding=0, dilation=1))
        self.norm1d.append(nn.InstanceNorm1d(60, affine=True))

    def forward(self, waveforms):
        outputs = self.wav_norm1d(waveforms)
        for c, (conv1d, pool1d, norm1d) in enumerate(zip(self.conv1d, self.pool1d, self.norm1d)):
            outputs = conv1d(outputs)
            if c == 0:
                outputs = torch.abs(outputs)
            outputs = F.leaky_relu(norm1d(pool1d(outputs)))
        return outputs


# receptive-field arithmetic: utils/receptive_field.py:26-165 (restated; pinned by ref_loader test)
def conv1d_num_frames(n, kernel_size=5, stride=1, padding=0, dilation=1):
    return 1 + (n + 2 * padding - dilation * (kernel_size - 1) - 1) // stride


def multi_conv_num_frames(n, kernel_size, stride, padding, dilation):
    for k, s, p, d in zip(kernel_size, stride, padding, dilation):
        n = conv1d_num_frames(n, k, s, p, d)
    return n


def conv1d_receptive_field_size(num_frames=1, kernel_size=5, stride=1, padding=0, dilation=1):
    effective = 1 + (kernel_size - 1) * dilation
    return effective + (num_frames - 1) * stride - 2 * padding


def multi_conv_receptive_field_size(num_frames, kernel_size, stride, padding, dilation):
    rf = num_frames
    for k, s, p, d in reversed(list(zip(kernel_size, stride, padding, dilation))):
        rf = conv1d_receptive_field_size(rf, k, s, p, d)
    return rf


def conv1d_receptive_field_center(frame=0, kernel_size=5, stride=1, padding=0, dilation=1):
    effective = 1 + (kernel_size - 1) * dilation
    return frame * stride + (effective - 1) // 2 - padding


def multi_conv_receptive_field_center(frame, kernel_size, stride, padding, dilation):
    c = frame
    for k, s, p, d in reversed(list(zip(kernel_size, stride, padding, dilation))):
        c = conv1d_receptive_field_center(c, k, s, p, d)
    return c


SINCNET_K = [251, 3, 5, 3, 5, 3]
SINCNET_S = [10, 3, 1, 3, 1, 3]
SINCNET_P = [0] * 6
SINCNET_D = [1] * 6


def sincnet_num_frames(num_samples):
    return multi_conv_num_frames(num_samples, SINCNET_K, SINCNET_S, SINCNET_P, SINCNET_D)


def sincnet_receptive_field(sample_rate=16000):
    """(start, duration, step) in seconds -- core/model.py:168-183."""
    size = multi_conv_receptive_field_size(1, SINCNET_K, SINCNET_S, SINCNET_P, SINCNET_D)
    step = multi_conv_receptive_field_size(2, SINCNET_K, SINCNET_S, SINCNET_P, SINCNET_D) - size
    center = multi_conv_receptive_field_center(0, SINCNET_K, SINCNET_S, SINCNET_P, SINCNET_D)
    start = center - (size - 1) / 2
    return start / sample_rate, size / sample_rate, step / sample_rate


class PyanNet(nn.Module):
    """Pretrained community-1 segmentation shape: 4-layer BiLSTM(128), 2xLinear(128), 7 classes."""

    def __init__(self, lstm_layers=4, num_classes=7):
        super().__init__()
        self.sincnet = SincNet(stride=10)
        self.lstm = nn.LSTM(60, hidden_size=128, num_layers=lstm_layers, bidirectional=True,
                            batch_first=True, dropout=0.0)
        self.linear = nn.ModuleList([nn.Linear(256, 128), nn.Linear(128, 128)])
        self.classifier = nn.Linear(128, num_classes)
        self.activation = nn.LogSoftmax(dim=-1)

    def forward(self, waveforms):
        outputs = self.sincnet(waveforms)
        outputs, _ = self.lstm(outputs.transpose(1, 2))
        for linear in self.linear:
            outputs = F.leaky_relu(linear(outputs))
        return self.activation(self.classifier(outputs))


# ----------------------------------------------------------------------------------------
# Powerset
# ----------------------------------------------------------------------------------------


def powerset_mapping(num_classes=3, max_set_size=2):
    rows = []
    for set_size in range(0, max_set_size + 1):
        for current_set in combinations(range(num_classes), set_size):
            row = torch.zeros(num_classes)
            row[list(current_set)] = 1
            rows.append(row)
    return torch.stack(rows)


def powerset_to_multilabel(powerset, mapping):
    """Hard conversion: one_hot(argmax) @ mapping (utils/powerset.py:135-140)."""
    probs = F.one_hot(torch.argmax(powerset, dim=-1), mapping.shape[0]).float()
    return torch.matmul(probs, mapping)


# ----------------------------------------------------------------------------------------
# WeSpeaker ResNet34
# ----------------------------------------------------------------------------------------


def stats_pool(sequences, weights=None):
    """models/blocks/pooling.py:76-130 (+ _pool :30-61)."""
    if weights is None:
        mean = sequences.mean(dim=-1)
        std = sequences.std(dim=-1, correction=1)
        return torch.cat([mean, std], dim=-1)
    if weights.dim() == 2:
        has_spk = False
        weights = weights.unsqueeze(dim=1)
    else:
        has_spk = True
    _, _, num_frames = sequences.size()
    _, num_speakers, num_weights = weights.size()
    if num_frames != num_weights:
        weights = F.interpolate(weights, size=num_frames, mode="nearest")

    def _pool(seq, w):
        w = w.unsqueeze(dim=1)
        v1 = w.sum(dim=2) + 1e-8
        mean = torch.sum(seq * w, dim=2) / v1
        dx2 = torch.square(seq - mean.unsqueeze(2))
        v2 = torch.square(w).sum(dim=2)
        var = torch.sum(dx2 * w, dim=2) / (v1 - v2 / v1 + 1e-8)
        return torch.cat([mean, torch.sqrt(var)], dim=1)

    out = torch.stack([_pool(sequences, weights[:, s, :]) for s in range(num_speakers)], dim=1)
    return out if has_spk else out.squeeze(dim=1)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != planes:
            self.shortcut = nn.Sequential(
                nn.Conv2d(in_planes, planes, 1, stride=stride, bias=False), nn.BatchNorm2d(planes))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        out = out + self.shortcut(x)
        return F.relu(out)


class ResNet34(nn.Module):
    def __init__(self, feat_dim=80, embed_dim=256, m_channels=32, num_blocks=(3, 4, 6, 3)):
        super().__init__()
        self.in_planes = m_channels
        self.stats_dim = int(feat_dim / 8) * m_channels * 8
        self.conv1 = nn.Conv2d(1, m_channels, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(m_channels)
        self.layer1 = self._make_layer(m_channels, num_blocks[0], 1)
        self.layer2 = self._make_layer(m_channels * 2, num_blocks[1], 2)
        self.layer3 = self._make_layer(m_channels * 4, num_blocks[2], 2)
        self.layer4 = self._make_layer(m_channels * 8, num_blocks[3], 2)
        self.seg_1 = nn.Linear(self.stats_dim * 2, embed_dim)

    def _make_layer(self, planes, n, stride):
        layers = []
        for s in [stride] + [1] * (n - 1):
            layers.append(BasicBlock(self.in_planes, planes, s))
            self.in_planes = planes
        return nn.Sequential(*layers)

    def forward_frames(self, fbank):
        x = fbank.permute(0, 2, 1).unsqueeze(1)  # (B,T,F) -> (B,1,F,T)
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.layer4(self.layer3(self.layer2(self.layer1(out))))
        return out  # (B,256,10,T/8)

    def forward_embedding(self, frames, weights=None):
        b, c, f, t = frames.shape
        stats = stats_pool(frames.reshape(b, c * f, t), weights=weights)
        return self.seg_1(stats)

    def forward(self, fbank, weights=None):
        return self.forward_embedding(self.forward_frames(fbank), weights=weights)


class WeSpeakerResNet34(nn.Module):
    def __init__(self):
        super().__init__()
        self.resnet = ResNet34(80, 256)

    @staticmethod
    def compute_fbank(waveforms):
        """models/embedding/wespeaker/__init__.py:113-139 (global-mean branch)."""
        import torchaudio.compliance.kaldi as kaldi

        waveforms = waveforms * (1 << 15)
        feats = torch.stack([
            kaldi.fbank(w, num_mel_bins=80, frame_length=25, frame_shift=10,
                        round_to_power_of_two=True, snip_edges=True, dither=0.0,
                        sample_frequency=16000, window_type="hamming", use_energy=False)
            for w in waveforms])
        return feats - torch.mean(feats, dim=1, keepdim=True)

    def forward(self, waveforms, weights=None):
        return self.resnet(self.compute_fbank(waveforms), weights=weights)

    def forward_frames(self, waveforms):
        return self.resnet.forward_frames(self.compute_fbank(waveforms))

    def forward_embedding(self, frames, weights=None):
        return self.resnet.forward_embedding(frames, weights=weights)
