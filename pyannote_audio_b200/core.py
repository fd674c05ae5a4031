"""Minimal host-side mirror of the pyannote.core / pyannote.audio.core.task types the hot path exchanges.

pyannote.core (6.0.1, not in the reference tree) provides ``Segment``, ``SlidingWindow``,
``SlidingWindowFeature`` and ``Annotation``; pyannote.audio.core.task provides ``Specifications`` /
``Problem`` / ``Resolution`` (/root/reference/src/pyannote/audio/core/task.py:59-137).  Only the members used by
``Inference`` / ``SpeakerDiarization`` are provided, with the same names and semantics.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Iterator, List, Optional, Tuple

import numpy as np


class Problem(Enum):
    BINARY_CLASSIFICATION = 0
    MONO_LABEL_CLASSIFICATION = 1
    MULTI_LABEL_CLASSIFICATION = 2
    REPRESENTATION = 3
    REGRESSION = 4


class Resolution(Enum):
    FRAME = 1
    CHUNK = 2


@dataclass
class Specifications:
    problem: Problem
    resolution: Resolution
    duration: float
    min_duration: Optional[float] = None
    warm_up: Tuple[float, float] = (0.0, 0.0)
    classes: Optional[List[str]] = None
    powerset_max_classes: Optional[int] = None
    permutation_invariant: bool = False

    @property
    def powerset(self) -> bool:
        if self.powerset_max_classes is None:
            return False
        if self.problem != Problem.MONO_LABEL_CLASSIFICATION:
            raise ValueError("`powerset_max_classes` only makes sense with multi-class classification problems.")
        return True

    @property
    def num_powerset_classes(self) -> int:
        from math import comb

        return int(sum(comb(len(self.classes), i) for i in range(0, self.powerset_max_classes + 1)))

    def __len__(self):
        return 1

    def __iter__(self):
        yield self


SEGMENT_PRECISION = 1e-6      # pyannote.core.segment.SEGMENT_PRECISION: shorter segments are "empty"


@dataclass(frozen=True, order=True)
class Segment:
    start: float = 0.0
    end: float = 0.0

    @property
    def duration(self) -> float:
        return max(0.0, self.end - self.start)

    @property
    def middle(self) -> float:
        return 0.5 * (self.start + self.end)

    def __and__(self, other: "Segment") -> "Segment":
        return Segment(max(self.start, other.start), min(self.end, other.end))

    def __bool__(self):
        return (self.end - self.start) > SEGMENT_PRECISION

    def __iter__(self):
        yield self.start
        yield self.end


class SlidingWindow:
    def __init__(self, duration: float = 0.030, step: float = 0.010, start: float = 0.0, end: Optional[float] = None):
        if duration <= 0:
            raise ValueError("'duration' must be a float > 0.")
        if step <= 0:
            raise ValueError("'step' must be a float > 0.")
        self.duration, self.step, self.start = float(duration), float(step), float(start)
        self.end = np.inf if end is None else end

    def closest_frame(self, t: float) -> int:
        return int(np.rint((t - self.start - 0.5 * self.duration) / self.step))

    def closest_frames(self, t: np.ndarray) -> np.ndarray:
        return np.rint((np.asarray(t) - self.start - 0.5 * self.duration) / self.step).astype(np.int64)

    def __getitem__(self, i: int) -> Segment:
        start = self.start + i * self.step
        return Segment(start, start + self.duration)

    def range_to_segment(self, i0: int, n: int) -> Segment:
        start = self.start + (i0 - 0.5) * self.step + 0.5 * self.duration
        end = start + n * self.step
        if i0 == 0:
            start = self.start
        return Segment(start, end)

    def crop(self, focus: Segment, mode: str = "loose") -> Tuple[int, int]:
        if mode == "loose":
            i = int(np.ceil((focus.start - self.duration - self.start) / self.step))
            j = int(np.floor((focus.end - self.start) / self.step))
            return i, j + 1
        if mode == "strict":
            i = int(np.ceil((focus.start - self.start) / self.step))
            j = int(np.floor((focus.end - self.duration - self.start) / self.step))
            return i, j + 1
        if mode == "center":
            return self.closest_frame(focus.start), self.closest_frame(focus.end) + 1
        raise ValueError("'mode' must be one of {'loose', 'strict', 'center'}.")

    def __eq__(self, other):
        return (isinstance(other, SlidingWindow) and self.duration == other.duration and self.step == other.step
                and self.start == other.start)

    def __repr__(self):
        return f"SlidingWindow(start={self.start:g}, duration={self.duration:g}, step={self.step:g})"


class SlidingWindowFeature:
    def __init__(self, data: np.ndarray, sliding_window: SlidingWindow, labels: Optional[List[str]] = None):
        self.data = data
        self.sliding_window = sliding_window
        self.labels = labels

    def __len__(self):
        return self.data.shape[0]

    @property
    def extent(self) -> Segment:
        return self.sliding_window.range_to_segment(0, len(self))

    def __iter__(self) -> Iterator[Tuple[Segment, np.ndarray]]:
        for i in range(len(self)):
            yield self.sliding_window[i], self.data[i]

    def crop(self, focus: Segment, mode: str = "loose", return_data: bool = True):
        i, j = self.sliding_window.crop(focus, mode=mode)
        n = self.data.shape[0]
        if j < 0 or i >= n:
            data, i0 = self.data[:0], 0
        else:
            i0, j0 = max(i, 0), min(j, n)
            data = self.data[i0:j0]
        if return_data:
            return data
        sw = SlidingWindow(start=self.sliding_window[i0].start, duration=self.sliding_window.duration,
                           step=self.sliding_window.step)
        return SlidingWindowFeature(data, sw, labels=self.labels)


class Annotation:
    """Ordered collection of (segment, track, label); just what DiarizeOutput / RTTM writing need.

    Bulk-constructed annotations (``from_rows``) keep numpy arrays and only build ``Segment`` objects when iterated.
    """

    def __init__(self, uri: Optional[str] = None):
        self.uri = uri
        self._tracks: Optional[List[Tuple[Segment, object, object]]] = []
        self._sorted = True
        self._rows = None      # (starts f64, ends f64, labels object array), already in itertracks order

    @classmethod
    def from_rows(cls, starts: np.ndarray, ends: np.ndarray, labels, uri: Optional[str] = None) -> "Annotation":
        """Rows must already be sorted by (start, end, track).  ``labels`` may be an integer array (kept as such:
        relabelling thousands of segments then is one table lookup instead of a Python loop)."""
        a = cls(uri=uri)
        lab = np.asarray(labels)
        if lab.dtype.kind not in "iu":
            lab = np.asarray(labels, dtype=object)
        a._rows = (np.asarray(starts, dtype=np.float64), np.asarray(ends, dtype=np.float64), lab)
        a._tracks = None
        return a

    def _materialise(self):
        if self._tracks is None:
            st, en, lab = self._rows
            self._tracks = [(Segment(float(a), float(b)), i, l) for i, (a, b, l) in enumerate(zip(st, en, lab.tolist()))]
            self._sorted = True
            self._rows = None

    def __setitem__(self, key, label):
        segment, track = key
        self.add(segment, track, label)

    def add(self, segment: Segment, track, label):
        if not segment:            # pyannote.core: "do not add empty track"
            return
        self._materialise()
        self._tracks.append((segment, track, label))
        self._sorted = False

    def _sort(self):
        self._materialise()
        if not self._sorted:
            self._tracks.sort(key=lambda r: (r[0].start, r[0].end))   # stable: ties keep insertion (track) order
            self._sorted = True

    def itertracks(self, yield_label: bool = False):
        self._sort()
        for segment, track, label in self._tracks:
            yield (segment, track, label) if yield_label else (segment, track)

    def itersegments(self):
        for segment, _ in self.itertracks():
            yield segment

    def labels(self):
        if self._tracks is None:
            lab = self._rows[2]
            if lab.dtype.kind in "iu":
                return np.unique(lab).tolist()
            return sorted(set(lab.tolist()), key=lambda v: (str(type(v)), v))
        return sorted({label for _, _, label in self._tracks}, key=lambda v: (str(type(v)), v))

    def rename_labels(self, mapping: dict) -> "Annotation":
        if self._tracks is None:
            st, en, lab = self._rows
            if len(lab) and lab.dtype.kind in "iu":
                uniq, inv = np.unique(lab, return_inverse=True)
                table = np.empty(len(uniq), dtype=object)
                table[:] = [mapping.get(u, u) for u in uniq.tolist()]
                new = table[inv]
            else:
                new = np.array([mapping.get(l, l) for l in lab.tolist()], dtype=object) if len(lab) else lab
            return Annotation.from_rows(st, en, new, uri=self.uri)
        out = Annotation(uri=self.uri)
        out._tracks = [(s, t, mapping.get(lab, lab)) for s, t, lab in self._tracks]
        out._sorted = self._sorted
        return out

    def support(self, collar: float = 0.0) -> "Annotation":
        """Per label (sorted), merge segments that touch / overlap or whose gap is strictly shorter than `collar`
        seconds (pyannote.core Timeline.support_iter: ``not gap or gap.duration < collar``)."""
        out = Annotation(uri=self.uri)
        by_label = {}
        for s, _, lab in self.itertracks(yield_label=True):
            by_label.setdefault(lab, []).append(s)
        n = 0
        for lab in sorted(by_label, key=lambda v: (str(type(v)), v)):
            segs = sorted(by_label[lab])
            cur = segs[0]
            for s in segs[1:]:
                gap = s.start - cur.end
                if gap <= SEGMENT_PRECISION or gap < collar:
                    cur = Segment(cur.start, max(cur.end, s.end))
                else:
                    out.add(cur, n, lab)
                    n += 1
                    cur = s
            out.add(cur, n, lab)
            n += 1
        return out

    def __len__(self):
        return len(self._rows[0]) if self._tracks is None else len(self._tracks)

    def __bool__(self):
        return len(self) > 0

    def _iter_rttm(self):
        """pyannote.core Annotation._iter_rttm: one SPEAKER line per track; names with spaces are refused."""
        uri = self.uri if self.uri else "<NA>"
        if isinstance(uri, str) and " " in uri:
            raise ValueError(f'Space-separated RTTM file format does not allow file URIs containing spaces (got: "{uri}").')
        for s, _, lab in self.itertracks(yield_label=True):
            if isinstance(lab, str) and " " in lab:
                raise ValueError(f'Space-separated RTTM file format does not allow labels containing spaces (got: "{lab}").')
            yield f"SPEAKER {uri} 1 {s.start:.3f} {s.duration:.3f} <NA> <NA> {lab} <NA> <NA>\n"

    def to_rttm(self) -> str:
        return "".join(self._iter_rttm())

    def write_rttm(self, file) -> None:
        """Dump to an open text file, as the reference's CLI does (__main__.py:705-706)."""
        for line in self._iter_rttm():
            file.write(line)

    def get_timeline(self) -> List[Segment]:
        """Segments in chronological order (pyannote.core returns a Timeline; a sorted list of unique segments here)."""
        return sorted(set(self.itersegments()))

    def label_duration(self, label) -> float:
        """Total duration of the union of `label`'s segments (pyannote.core Annotation.label_duration)."""
        total, end = 0.0, -np.inf
        for s in sorted(seg for seg, _, lab in self.itertracks(yield_label=True) if lab == label):
            if s.end > end:
                total += s.end - max(s.start, end)
                end = s.end
        return total

    def chart(self) -> List[Tuple[object, float]]:
        """(label, duration) pairs, longest first (pyannote.core Annotation.chart)."""
        return sorted(((lab, self.label_duration(lab)) for lab in self.labels()), key=lambda x: x[1], reverse=True)
