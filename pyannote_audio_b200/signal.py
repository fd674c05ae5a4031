"""Binarize (mirror of /root/reference/src/pyannote/audio/utils/signal.py:181-318): hysteresis thresholding of
frame-level scores into an Annotation, vectorised (the reference runs a Python loop per frame and per class)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .core import Annotation, SlidingWindowFeature


def hysteresis(y: np.ndarray, onset: float, offset: float) -> np.ndarray:
    """State of the reference's loop (signal.py:276-300) after each sample: starts as y[0] > onset; an active region
    ends at the first sample < offset, an inactive one ends at the first sample > onset."""
    n = len(y)
    if n == 0:
        return np.zeros(0, dtype=bool)
    up, down = y > onset, y < offset
    if bool((up & down).any()):                 # offset > onset: a sample can satisfy both -> plain sequential scan
        active = np.empty(n, dtype=bool)
        a = bool(y[0] > onset)
        active[0] = a
        for t in range(1, n):
            if a:
                if down[t]:
                    a = False
            elif up[t]:
                a = True
            active[t] = a
        return active
    ev = np.zeros(n, dtype=np.int8)
    ev[up] = 1
    ev[down] = -1
    ev[0] = 1 if y[0] > onset else -1           # initial state
    idx = np.where(ev != 0, np.arange(n), 0)
    np.maximum.accumulate(idx, out=idx)         # index of the last event at or before t
    return ev[idx] > 0


class Binarize:
    def __init__(self, onset: float = 0.5, offset: Optional[float] = None, min_duration_on: float = 0.0,
                 min_duration_off: float = 0.0, pad_onset: float = 0.0, pad_offset: float = 0.0):
        self.onset = onset
        self.offset = offset or onset
        self.pad_onset, self.pad_offset = pad_onset, pad_offset
        self.min_duration_on, self.min_duration_off = min_duration_on, min_duration_off

    def __call__(self, scores: SlidingWindowFeature) -> Annotation:
        from .pipeline import binarize_frames

        if self.pad_onset or self.pad_offset:
            raise NotImplementedError("pad_onset / pad_offset are not used by the diarization / VAD pipelines")
        data = np.asarray(scores.data)
        active = np.stack([hysteresis(data[:, k], self.onset, self.offset) for k in range(data.shape[1])], axis=1)
        ann, _ = binarize_frames(active.astype(np.uint8), scores.sliding_window, self.min_duration_off)
        if scores.labels is not None:
            ann = ann.rename_labels({k: lab for k, lab in enumerate(scores.labels)})
        if self.min_duration_on > 0:
            kept = Annotation(uri=ann.uri)
            for segment, track, label in ann.itertracks(yield_label=True):
                if segment.duration >= self.min_duration_on:
                    kept.add(segment, track, label)
            ann = kept
        return ann
