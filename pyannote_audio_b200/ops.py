"""Thin Python layer over the C ABI: one ``Context`` per device, torch tensors for device memory and streams.

Nothing here computes on the CPU: every op forwards raw device pointers to libb200diar.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Optional, Sequence

import numpy as np
import torch

from . import _lib

CHUNK = 160000
FRAMES = 589
SPEAKERS = 3
CLASSES = 7
EMB_DIM = 256


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _fp(t: torch.Tensor):
    return C.cast(C.c_void_p(t.data_ptr()), _lib.c_float_p)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def sinc_filter_bank(low_hz_: torch.Tensor, band_hz_: torch.Tensor, window_: Optional[torch.Tensor] = None,
                     n_: Optional[torch.Tensor] = None, sample_rate: float = 16000.0, min_low_hz: float = 50.0,
                     min_band_hz: float = 50.0, kernel_size: int = 251) -> torch.Tensor:
    """Realise the (80,251) ParamSincFB filter bank from its parameters (done once at load, the weights are frozen).

    Same arithmetic, same op order as asteroid_filterbanks.ParamSincFB.filters() (call site
    /root/reference/src/pyannote/audio/models/blocks/sincnet.py:58-69), in torch fp32 on the CPU.
    """
    half = kernel_size // 2
    low_hz_ = low_hz_.detach().float().cpu()
    band_hz_ = band_hz_.detach().float().cpu()
    if window_ is None:
        window_ = torch.from_numpy(np.hamming(kernel_size)[:half]).float()
    if n_ is None:
        n_ = 2 * np.pi * (torch.arange(-half, 0.0).view(1, -1) / sample_rate)
    window_, n_ = window_.float().cpu(), n_.float().cpu()
    low = min_low_hz + torch.abs(low_hz_)
    high = torch.clamp(low + min_band_hz + torch.abs(band_hz_), min_low_hz, sample_rate / 2)
    band = (high - low)[:, 0]
    ft_low, ft_high = torch.matmul(low, n_), torch.matmul(high, n_)
    cos_left = ((torch.sin(ft_high) - torch.sin(ft_low)) / (n_ / 2)) * window_
    cos = torch.cat([cos_left, 2 * band.view(-1, 1), torch.flip(cos_left, dims=[1])], dim=1)
    sin_left = ((torch.cos(ft_low) - torch.cos(ft_high)) / (n_ / 2)) * window_
    sin = torch.cat([sin_left, torch.zeros_like(band.view(-1, 1)), -torch.flip(sin_left, dims=[1])], dim=1)
    bank = torch.cat([cos / (2 * band[:, None]), sin / (2 * band[:, None])], dim=0)
    return bank.contiguous()


class Context:
    """Owns a ``b200_ctx`` (weights + workspaces) on one CUDA device."""

    def __init__(self, device: torch.device | int | str = "cuda:0"):
        self.lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.B200Error(f"pyannote_audio_b200 runs on CUDA (sm_100a) devices only, got '{device}'")
        if not torch.cuda.is_available():
            raise _lib.B200Error("no CUDA device is visible: pyannote_audio_b200 has no CPU fallback")
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        torch.cuda.init()
        h = C.c_void_p()
        _lib.check(self.lib.b200_ctx_create(C.byref(h), self.device.index))
        self._h = h
        self.seg_loaded = False
        self.emb_loaded = False
        self.owners = {}          # slot ("seg" | "emb") -> stamp of the model whose weights are resident (models.py)
        # A/B knob for scripts (same role as the B200_* variables the library reads): B200_OPTIONS="key=value,..."
        # is applied through b200_ctx_set_option, so unknown keys / bad values fail loudly
        import os

        for kv in filter(None, os.environ.get("B200_OPTIONS", "").split(",")):
            key, value = kv.split("=")
            self.set_option(key.strip(), int(value))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        _lib.check(self.lib.b200_ctx_set_option(self._h, key.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.b200_ctx_launch_count(self._h))

    def timer(self, name: str):
        """(accumulated device ms, units) of a profiled region ("trunk" | "seg"); needs set_option("profile", 1)."""
        ms, units = C.c_double(0.0), C.c_int64(0)
        _lib.check(self.lib.b200_ctx_timer(self._h, name.encode(), C.byref(ms), C.byref(units)))
        return float(ms.value), int(units.value)

    # ---- weights ---------------------------------------------------------------------------------
    def load_segmentation(self, sd: Mapping[str, torch.Tensor]):
        keep = []

        def f(name):
            t = sd[name].detach().to(torch.float32).cpu().contiguous()
            keep.append(t)
            return _fp(t)

        w = _lib.SegWeights()
        w.wav_norm_weight = float(sd["sincnet.wav_norm1d.weight"].reshape(-1)[0])
        w.wav_norm_bias = float(sd["sincnet.wav_norm1d.bias"].reshape(-1)[0])
        p = "sincnet.conv1d.0.filterbank."
        bank = sinc_filter_bank(sd[p + "low_hz_"], sd[p + "band_hz_"], sd.get(p + "window_"), sd.get(p + "n_"))
        keep.append(bank)
        w.sinc_filters = _fp(bank)
        for i in range(3):
            w.norm_weight[i] = f(f"sincnet.norm1d.{i}.weight")
            w.norm_bias[i] = f(f"sincnet.norm1d.{i}.bias")
        for i in range(2):
            w.conv_weight[i] = f(f"sincnet.conv1d.{i + 1}.weight")
            w.conv_bias[i] = f(f"sincnet.conv1d.{i + 1}.bias")
        layers = 0
        while f"lstm.weight_ih_l{layers}" in sd:
            layers += 1
        w.lstm_layers = layers
        for layer in range(layers):
            for d, suffix in enumerate(("", "_reverse")):
                w.lstm_w_ih[layer * 2 + d] = f(f"lstm.weight_ih_l{layer}{suffix}")
                w.lstm_w_hh[layer * 2 + d] = f(f"lstm.weight_hh_l{layer}{suffix}")
                w.lstm_b_ih[layer * 2 + d] = f(f"lstm.bias_ih_l{layer}{suffix}")
                w.lstm_b_hh[layer * 2 + d] = f(f"lstm.bias_hh_l{layer}{suffix}")
        for i in range(2):
            w.linear_weight[i] = f(f"linear.{i}.weight")
            w.linear_bias[i] = f(f"linear.{i}.bias")
        w.classifier_weight = f("classifier.weight")
        w.classifier_bias = f("classifier.bias")
        if sd["classifier.weight"].shape[0] != CLASSES:
            raise ValueError("only the 7-class powerset (3 speakers, max 2 simultaneous) head is supported")
        self.owners.pop("seg", None)          # whoever uploaded before no longer owns the slot
        _lib.check(self.lib.b200_seg_load(self._h, C.byref(w)))
        self.seg_loaded = True

    def load_embedding(self, sd: Mapping[str, torch.Tensor]):
        keep = []

        def f(name):
            t = sd[name].detach().to(torch.float32).cpu().contiguous()
            keep.append(t)
            return _fp(t)

        def conv_bn(dst, conv, bn):
            dst.conv_weight = f(conv + ".weight")
            dst.bn_weight = f(bn + ".weight")
            dst.bn_bias = f(bn + ".bias")
            dst.bn_mean = f(bn + ".running_mean")
            dst.bn_var = f(bn + ".running_var")

        w = _lib.EmbWeights()
        conv_bn(w.stem, "resnet.conv1", "resnet.bn1")
        bi = 0
        for li, n in enumerate((3, 4, 6, 3), start=1):
            for i in range(n):
                p = f"resnet.layer{li}.{i}"
                conv_bn(w.block_conv1[bi], p + ".conv1", p + ".bn1")
                conv_bn(w.block_conv2[bi], p + ".conv2", p + ".bn2")
                if p + ".shortcut.0.weight" in sd:
                    conv_bn(w.block_shortcut[bi], p + ".shortcut.0", p + ".shortcut.1")
                bi += 1
        w.seg1_weight = f("resnet.seg_1.weight")
        w.seg1_bias = f("resnet.seg_1.bias")
        self.owners.pop("emb", None)
        _lib.check(self.lib.b200_emb_load(self._h, C.byref(w)))
        self.emb_loaded = True

    # ---- helpers ---------------------------------------------------------------------------------
    def _chunks(self, wav: torch.Tensor, chunk_off, chunk_valid):
        if wav.device != self.device or wav.dtype != torch.float32 or not wav.is_contiguous():
            raise ValueError(f"waveform must be a contiguous float32 tensor on {self.device}")
        off = np.ascontiguousarray(chunk_off, dtype=np.int64)
        valid = np.ascontiguousarray(chunk_valid, dtype=np.int32)
        if off.shape != valid.shape or off.ndim != 1:
            raise ValueError("chunk_off / chunk_valid must be 1-D and of equal length")
        if len(off) and int((off + valid).max()) > wav.numel():
            raise ValueError("a chunk reads past the end of the waveform buffer")
        return off, valid

    # ---- segmentation ----------------------------------------------------------------------------
    def _out(self, out: Optional[torch.Tensor], shape, dtype) -> torch.Tensor:
        """Caller-provided output (e.g. this rank's slice of a collective's buffer) or a fresh tensor."""
        if out is None:
            return torch.empty(shape, dtype=dtype, device=self.device)
        if tuple(out.shape) != tuple(shape) or out.dtype != dtype or out.device != self.device or \
                not out.is_contiguous():
            raise ValueError(f"`out` must be a contiguous {dtype} tensor of shape {tuple(shape)} on {self.device}")
        return out

    def seg_forward(self, wav, chunk_off, chunk_valid, return_logp=False, out: Optional[torch.Tensor] = None):
        off, valid = self._chunks(wav, chunk_off, chunk_valid)
        n = len(off)
        cls = self._out(out, (n, FRAMES), torch.uint8)
        logp = torch.empty((n, FRAMES, CLASSES), dtype=torch.float32, device=self.device) if return_logp else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_seg_forward(self._h, _ptr(wav), off.ctypes.data, valid.ctypes.data, n, _ptr(cls),
                                                 _ptr(logp), _stream(self.device)))
        return (cls, logp) if return_logp else cls

    def sincnet_forward(self, wav, chunk_off, chunk_valid):
        off, valid = self._chunks(wav, chunk_off, chunk_valid)
        n = len(off)
        out = torch.empty((n, FRAMES, 60), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_sincnet_forward(self._h, _ptr(wav), off.ctypes.data, valid.ctypes.data, n,
                                                     _ptr(out), _stream(self.device)))
        return out

    def powerset_to_multilabel(self, cls: torch.Tensor):
        out = torch.empty(tuple(cls.shape) + (SPEAKERS,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_powerset_to_multilabel(self._h, _ptr(cls), cls.numel(), _ptr(out),
                                                            _stream(self.device)))
        return out

    # ---- embeddings ------------------------------------------------------------------------------
    def emb_forward(self, wav, chunk_off, chunk_valid, masks: torch.Tensor, out: Optional[torch.Tensor] = None,
                    peers: Optional[Sequence[int]] = None):
        """``peers``: raw device addresses of this rank's (n,3,256) float32 slot inside other GPUs' gather buffers
        (peer-mapped memory); the final GEMM's epilogue then pushes the embeddings there (fused all-gather)."""
        off, valid = self._chunks(wav, chunk_off, chunk_valid)
        n = len(off)
        if tuple(masks.shape) != (n, SPEAKERS, FRAMES) or masks.dtype != torch.uint8 or not masks.is_contiguous():
            raise ValueError(f"masks must be a contiguous uint8 tensor of shape ({n}, 3, 589)")
        emb = self._out(out, (n, SPEAKERS, EMB_DIM), torch.float32)
        with torch.cuda.device(self.device):
            if peers:
                arr = (C.c_void_p * len(peers))(*[C.c_void_p(int(a)) for a in peers])
                _lib.check(self.lib.b200_emb_forward_push(self._h, _ptr(wav), off.ctypes.data, valid.ctypes.data, n,
                                                          _ptr(masks), _ptr(emb), arr, len(peers),
                                                          _stream(self.device)))
            else:
                _lib.check(self.lib.b200_emb_forward(self._h, _ptr(wav), off.ctypes.data, valid.ctypes.data, n,
                                                     _ptr(masks), _ptr(emb), _stream(self.device)))
        return emb

    def push(self, src: torch.Tensor, peers: Sequence[int]):
        """P2P copy of a contiguous device tensor to raw peer addresses (same layout), on the current stream."""
        if not peers or src.numel() == 0:
            return
        arr = (C.c_void_p * len(peers))(*[C.c_void_p(int(a)) for a in peers])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_push(self._h, _ptr(src), src.numel() * src.element_size(), arr, len(peers),
                                          _stream(self.device)))

    def emb_fbank(self, wav, chunk_off, chunk_valid):
        off, valid = self._chunks(wav, chunk_off, chunk_valid)
        n = len(off)
        fb = torch.empty((n, 998, 80), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_emb_fbank(self._h, _ptr(wav), off.ctypes.data, valid.ctypes.data, n, _ptr(fb),
                                               _stream(self.device)))
        return fb

    def emb_trunk(self, fbank: torch.Tensor):
        n = fbank.shape[0]
        fbank = fbank.contiguous()
        out = torch.empty((n, 256, 10, 125), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_emb_trunk(self._h, _ptr(fbank), n, _ptr(out), _stream(self.device)))
        return out

    def stats_pool(self, seq: torch.Tensor, weights: Optional[torch.Tensor] = None):
        B, F, T = seq.shape
        seq = seq.contiguous().float()
        squeeze = False
        if weights is None:
            S, Tw = 1, T
            squeeze = True
        else:
            if weights.dim() == 2:
                weights = weights.unsqueeze(1)
                squeeze = True
            weights = weights.contiguous().float()
            S, Tw = weights.shape[1], weights.shape[2]
        out = torch.empty((B, S, 2 * F), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_stats_pool(self._h, _ptr(seq), _ptr(weights), _ptr(out), B, F, T, S, Tw,
                                                _stream(self.device)))
        return out.squeeze(1) if squeeze else out

    # ---- overlap-add / reconstruction ----------------------------------------------------------------
    def start_frames(self, start_frame) -> torch.Tensor:
        """Device copy of the per-chunk start frames (cached per distinct array)."""
        if isinstance(start_frame, torch.Tensor):
            return start_frame
        sf = np.ascontiguousarray(start_frame, dtype=np.int32)
        if len(sf) > 1 and not bool((np.diff(sf) >= 0).all()):
            raise ValueError("start_frame must be non-decreasing")
        key = (len(sf), int(sf[0]) if len(sf) else 0, int(sf[-1]) if len(sf) else 0)
        cache = self.__dict__.setdefault("_sf_cache", {})
        hit = cache.get(key)
        if hit is None or not np.array_equal(hit[0], sf):
            hit = (sf, torch.from_numpy(sf).to(self.device))
            if len(cache) > 64:
                cache.clear()
            cache[key] = hit
        return hit[1]

    def speaker_count(self, seg: torch.Tensor, start_frame, num_frames: int):
        sf = self.start_frames(start_frame)
        count = torch.empty((num_frames,), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_speaker_count(self._h, _ptr(seg), _ptr(sf), sf.numel(), num_frames, _ptr(count),
                                                   _stream(self.device)))
        return count

    def reconstruct(self, seg: torch.Tensor, hard_clusters, start_frame, num_frames: int, count: torch.Tensor,
                    num_clusters_out: int):
        sf = self.start_frames(start_frame)
        if not isinstance(hard_clusters, torch.Tensor):
            hard_clusters = torch.from_numpy(np.ascontiguousarray(hard_clusters, dtype=np.int8)).to(self.device)
        hc = hard_clusters.to(torch.int8).contiguous()
        out = torch.empty((num_frames, num_clusters_out), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_reconstruct(self._h, _ptr(seg), _ptr(hc), _ptr(sf), sf.numel(), num_frames,
                                                 _ptr(count), num_clusters_out, _ptr(out), _stream(self.device)))
        return out

    def frame_transitions(self, discrete: torch.Tensor, cap: int = 4096):
        """discrete (F, K) u8 on the device -> sorted flat event indices k * (F + 1) + f of onsets and offsets
        (host int64 arrays).  One kernel + one 32 KB D2H instead of shipping and scanning the whole matrix."""
        return self.frame_transitions_many([discrete], cap)[0]

    def frame_transitions_many(self, matrices: Sequence[torch.Tensor], cap: int = 4096):
        """frame_transitions of several (F_i, K_i) matrices with ONE device -> host copy (and one synchronisation)
        for all of them: every matrix gets a row [n_on, n_off, on[cap], off[cap]] of one int32 buffer.  A row whose
        event count exceeds ``cap`` is redone alone with a larger buffer."""
        m = len(matrices)
        self.last_transfer_bytes = 0
        if m == 0:
            return []
        mats = [d.contiguous() for d in matrices]
        buf = torch.empty((m, 2 + 2 * cap), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            for i, d in enumerate(mats):
                _lib.check(self.lib.b200_frame_transitions(self._h, _ptr(d), int(d.shape[0]), int(d.shape[1]), cap,
                                                           _ptr(buf[i]), _stream(self.device)))
        host = buf.cpu().numpy()
        self.last_transfer_bytes = host.nbytes
        out = []
        for i, d in enumerate(mats):
            row, c = host[i], cap
            n_on, n_off = int(row[0]), int(row[1])
            while max(n_on, n_off) > c:                       # rare: more events than slots -> this matrix again
                c = 1 << int(max(n_on, n_off) - 1).bit_length()
                big = torch.empty((2 + 2 * c,), dtype=torch.int32, device=self.device)
                with torch.cuda.device(self.device):
                    _lib.check(self.lib.b200_frame_transitions(self._h, _ptr(d), int(d.shape[0]), int(d.shape[1]), c,
                                                               _ptr(big), _stream(self.device)))
                row = big.cpu().numpy()
                self.last_transfer_bytes += row.nbytes
                n_on, n_off = int(row[0]), int(row[1])
            on = np.sort(row[2: 2 + n_on].astype(np.int64))
            off = np.sort(row[2 + c: 2 + c + n_off].astype(np.int64))
            out.append((on, off))
        return out

    # ---- audio ingest ------------------------------------------------------------------------------
    def audio_ingest(self, pcm: torch.Tensor, sample_rate: int, target_rate: Optional[int] = None,
                     channel: Optional[int] = None, downmix: bool = True, out: Optional[torch.Tensor] = None):
        """Raw decoded audio on the device -> float32 mono waveform (frames_out,) at ``target_rate``.

        ``pcm``: int16 (frames, channels) interleaved PCM, or float32 (channels, frames) like the reference's
        in-memory files.  ``channel`` selects one channel, otherwise channels are averaged (mono="downmix").
        (core/io.py:223-265: downmix, then torchaudio.functional.resample.)"""
        if pcm.device != self.device or not pcm.is_contiguous() or pcm.dim() != 2:
            raise ValueError(f"pcm must be a contiguous 2-D tensor on {self.device}")
        if pcm.dtype == torch.int16:
            fmt, (frames, channels) = 0, pcm.shape
        elif pcm.dtype == torch.float32:
            fmt, (channels, frames) = 1, pcm.shape
        else:
            raise ValueError("pcm must be int16 (frames, channels) or float32 (channels, frames)")
        if channel is None and not downmix and channels > 1:
            raise ValueError("multi-channel audio needs `channel` or downmix=True")
        target_rate = int(target_rate or sample_rate)
        n = int(self.lib.b200_audio_num_frames(int(frames), int(sample_rate), target_rate))
        if out is None:
            out = torch.empty((n,), dtype=torch.float32, device=self.device)
        elif out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous() or out.numel() < n:
            raise ValueError(f"`out` must be a contiguous float32 tensor with at least {n} elements on {self.device}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_audio_ingest(self._h, _ptr(pcm), fmt, int(channels), int(frames),
                                                  int(sample_rate), target_rate, -1 if channel is None else int(channel),
                                                  _ptr(out), out.numel(), _stream(self.device)))
        return out[:n]

    # ---- generic overlap-add (Inference.aggregate) ---------------------------------------------------
    def _window(self, key, build):
        cache = self.__dict__.setdefault("_win_cache", {})
        if key not in cache:
            cache[key] = torch.from_numpy(np.ascontiguousarray(build(), dtype=np.float64)).to(self.device)
        return cache[key]

    def aggregate(self, scores: torch.Tensor, start_frame, num_frames: int, hamming: bool = False,
                  warm_up=(0.0, 0.0), chunk_duration: float = 10.0, epsilon: float = 1e-12,
                  missing: float = float("nan"), skip_average: bool = False) -> torch.Tensor:
        """scores (C,589,K) float32 device (NaN = missing) -> (num_frames, K) float32 device, bit-identical to
        Inference.aggregate's numpy arithmetic (core/inference.py:498-620)."""
        if scores.dtype != torch.float32 or scores.device != self.device or scores.dim() != 3 or \
                scores.shape[1] != FRAMES:
            raise ValueError(f"scores must be a float32 (chunks, {FRAMES}, classes) tensor on {self.device}")
        scores = scores.contiguous()
        sf = self.start_frames(start_frame)
        ham = self._window("hamming", lambda: np.hamming(FRAMES)) if hamming else None
        wl = round(warm_up[0] / chunk_duration * FRAMES)
        wr = round(warm_up[1] / chunk_duration * FRAMES)
        warm = None
        if wl or wr:
            def build():
                w = np.ones(FRAMES)
                w[:wl] = epsilon
                w[FRAMES - wr:] = epsilon
                return w
            warm = self._window(("warm", wl, wr, epsilon), build)
        out = torch.empty((num_frames, scores.shape[2]), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_aggregate(self._h, _ptr(scores), _ptr(sf), sf.numel(), int(num_frames),
                                               int(scores.shape[2]), _ptr(ham), _ptr(warm), int(skip_average),
                                               float(missing), float(np.float32(epsilon)), _ptr(out),
                                               _stream(self.device)))
        return out

    def powerset_speech(self, cls: torch.Tensor) -> torch.Tensor:
        """(…) uint8 powerset classes -> (…, 1) float32 speech indicator (max over the speakers of the multilabel)."""
        out = torch.empty(tuple(cls.shape) + (1,), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_powerset_speech(self._h, _ptr(cls.contiguous()), cls.numel(), _ptr(out),
                                                     _stream(self.device)))
        return out

    def clean_frames(self, seg: torch.Tensor):
        n = seg.shape[0]
        clean = torch.empty((n, SPEAKERS), dtype=torch.int32, device=self.device)
        active = torch.empty((n, SPEAKERS), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.b200_clean_frames(self._h, _ptr(seg), n, _ptr(clean), _ptr(active),
                                                  _stream(self.device)))
        return clean, active


# ---- clustering (fp64) -------------------------------------------------------------------------------
def _ctx_method(fn):
    setattr(Context, fn.__name__, fn)
    return fn


def _norm_mode(normalize) -> int:
    """False/0: rows as given; True/1: fp64 L2 normalisation; "float32"/2: numpy's float32 normalisation of rows that
    hold float32 values (what the reference does to float32 embeddings before scipy's linkage)."""
    if normalize in ("float32", 2):
        return 2
    return int(bool(normalize))


@_ctx_method
def plda_transform(self, x: torch.Tensor, mean1, mean2, lda, mu, trT) -> torch.Tensor:
    """PLDA.__call__ on the device: x (n, Din) float64 -> (n, L) float64 (core/plda.py:50-63)."""
    x = x.contiguous()
    n, din = x.shape
    dout, L = trT.shape
    fea = torch.empty((n, L), dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_plda_transform(self._h, _ptr(x), n, din, dout, L, _ptr(mean1), _ptr(mean2),
                                                _ptr(lda.contiguous()), _ptr(mu), _ptr(trT.contiguous()), _ptr(fea),
                                                _stream(self.device)))
    return fea


@_ctx_method
def weighted_centroids(self, q: torch.Tensor, kept: torch.Tensor, train: torch.Tensor) -> torch.Tensor:
    """(W.T @ train) / W.sum(0).T with W = q[:, kept] (clustering.py:620-621): q (n,S), train (n,dim) float64."""
    q, train = q.contiguous(), train.contiguous()
    kept = kept.to(device=self.device, dtype=torch.int32).contiguous()
    out = torch.empty((kept.numel(), train.shape[1]), dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_weighted_centroids(self._h, _ptr(q), q.shape[0], q.shape[1], _ptr(kept), kept.numel(),
                                                    _ptr(train), train.shape[1], _ptr(out), _stream(self.device)))
    return out


@_ctx_method
def linkage_centroid(self, x: torch.Tensor, normalize=True) -> torch.Tensor:
    """x (n, dim) float64 on device -> Z (n-1, 4) float64 on device (scipy linkage format)."""
    n, dim = x.shape
    x = x.contiguous()
    Z = torch.empty((n - 1, 4), dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_linkage_centroid(self._h, _ptr(x), n, dim, _norm_mode(normalize), _ptr(Z),
                                                  _stream(self.device)))
    return Z


def fcluster_distance(Z: np.ndarray, t: float) -> np.ndarray:
    """Host tree cut, 1-based labels like scipy.cluster.hierarchy.fcluster(Z, t, 'distance')."""
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    n = Z.shape[0] + 1
    T = np.zeros(n, dtype=np.int32)
    _lib.check(_lib.load().b200_fcluster_distance(Z.ctypes.data, n, C.c_double(float(t)), T.ctypes.data))
    return T


@_ctx_method
def cdist_cosine(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.contiguous(), b.contiguous()
    m, dim = a.shape
    k = b.shape[0]
    d = torch.empty((m, k), dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_cdist_cosine(self._h, _ptr(a), m, _ptr(b), k, dim, _ptr(d), _stream(self.device)))
    return d


@_ctx_method
def vbx(self, fea: torch.Tensor, phi: torch.Tensor, gamma0: torch.Tensor, Fa: float, Fb: float, max_iters: int = 20,
        epsilon: float = 1e-4):
    fea, phi = fea.contiguous(), phi.contiguous()
    gamma = gamma0.contiguous().clone()
    n, D = fea.shape
    S = gamma.shape[1]
    pi = torch.empty((S,), dtype=torch.float64, device=self.device)
    iters = C.c_int32(0)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_vbx(self._h, _ptr(fea), _ptr(phi), n, D, S, C.c_double(Fa), C.c_double(Fb), max_iters,
                                     C.c_double(epsilon), _ptr(gamma), _ptr(pi), C.byref(iters),
                                     _stream(self.device)))
    return gamma, pi, int(iters.value)


@_ctx_method
def assign(self, soft: torch.Tensor, constrained: bool = True) -> torch.Tensor:
    soft = soft.contiguous()
    c, s, k = soft.shape
    assert s == SPEAKERS
    hard = torch.empty((c, s), dtype=torch.int8, device=self.device)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_assign(self._h, _ptr(soft), c, k, int(constrained), _ptr(hard), _stream(self.device)))
    return hard


@_ctx_method
def linkage_centroid_batched(self, x: torch.Tensor, row_offsets, normalize=True) -> torch.Tensor:
    """x (sum n_f, dim) f64; row_offsets host int array (F+1,) -> concatenated Z ((sum max(n_f-1,0)), 4) f64."""
    x = x.contiguous()
    ro = np.ascontiguousarray(row_offsets, dtype=np.int32)
    nz = int(np.maximum(np.diff(ro) - 1, 0).sum())
    Z = torch.empty((nz, 4), dtype=torch.float64, device=self.device)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_linkage_centroid_batched(self._h, _ptr(x), ro.ctypes.data, len(ro) - 1, x.shape[1],
                                                          _norm_mode(normalize), _ptr(Z), _stream(self.device)))
    return Z


@_ctx_method
def vbx_batched(self, fea: torch.Tensor, phi: torch.Tensor, gamma0: torch.Tensor, n, S, Fa: float, Fb: float,
                max_iters: int = 20, epsilon: float = 1e-4, want_iters: bool = False):
    """fea (sum n_f, D); gamma0 flat concatenation of the per-problem (n_f, S_f) initial responsibilities."""
    fea, phi = fea.contiguous(), phi.contiguous()
    gamma = gamma0.contiguous().clone()
    n = np.ascontiguousarray(n, dtype=np.int32)
    S = np.ascontiguousarray(S, dtype=np.int32)
    pi = torch.empty((int(S.sum()),), dtype=torch.float64, device=self.device)
    iters = np.zeros(len(n), dtype=np.int32)
    with torch.cuda.device(self.device):
        _lib.check(self.lib.b200_vbx_batched(self._h, _ptr(fea), _ptr(phi), n.ctypes.data, S.ctypes.data, len(n),
                                             fea.shape[1], C.c_double(Fa), C.c_double(Fb), max_iters,
                                             C.c_double(epsilon), _ptr(gamma), _ptr(pi),
                                             iters.ctypes.data if want_iters else None, _stream(self.device)))
    return gamma, pi, iters
