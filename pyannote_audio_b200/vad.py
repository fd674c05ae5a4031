"""Voice activity detection pipeline (mirror of /root/reference/src/pyannote/audio/pipelines/
voice_activity_detection.py:66-204) reusing the diarization kernels: PyanNet sliding window -> speech indicator per
frame (max over the speakers of the powerset multilabel) -> Hamming-windowed overlap-add on the device
(b200_aggregate) -> Binarize.  SURVEY.md section 8(f) row 3."""
from __future__ import annotations

from typing import Callable, Mapping, Optional, Union

import numpy as np
import torch

from . import ops
from .audio import AudioFile
from .core import Annotation, Segment, SlidingWindow, SlidingWindowFeature
from .inference import Inference
from .models import PyanNet
from .signal import Binarize


class VoiceActivityDetection:
    def __init__(self, segmentation: Union[PyanNet, Mapping, None] = None, fscore: bool = False, token=None,
                 cache_dir=None, device: Optional[torch.device] = None, **inference_kwargs):
        from .loading import get_model, is_checkpoint_spec

        if is_checkpoint_spec(segmentation):               # path / {"checkpoint": ...} from Pipeline.from_pretrained
            segmentation = get_model(segmentation, token=token, cache_dir=cache_dir)
        if isinstance(segmentation, Mapping):
            model = PyanNet()
            model.load_state_dict(segmentation)
            segmentation = model
        if not isinstance(segmentation, PyanNet):
            raise ValueError("`segmentation` must be a PyanNet instance or its state dict (no hub access here)")
        self.segmentation, self.fscore = segmentation, fscore
        device = device or torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        segmentation.to(device)
        inference_kwargs["pre_aggregation_hook"] = lambda scores: np.max(scores, axis=-1, keepdims=True)
        self._segmentation = Inference(segmentation, **inference_kwargs)
        # powerset model: thresholds are fixed (voice_activity_detection.py:117-118)
        self.onset = self.offset = 0.5
        self.min_duration_on = self.min_duration_off = 0.0
        self.initialize()

    def default_parameters(self):
        return {"min_duration_on": 0.0, "min_duration_off": 0.0}

    def instantiate(self, params: dict):
        for k in ("onset", "offset", "min_duration_on", "min_duration_off"):
            if k in params:
                setattr(self, k, float(params[k]))
        self.initialize()
        return self

    def classes(self):
        return ["SPEECH"]

    def initialize(self):
        self._binarize = Binarize(onset=self.onset, offset=self.offset, min_duration_on=self.min_duration_on,
                                  min_duration_off=self.min_duration_off)

    def speech_scores(self, file: AudioFile, hook: Optional[Callable] = None) -> SlidingWindowFeature:
        """Aggregated speech score per frame, (num_frames, 1) float32: what `self._segmentation(file)` returns in the
        reference (Inference with the max-over-speakers pre-aggregation hook), computed without leaving the device
        between the network and the overlap-add."""
        inf = self._segmentation
        waveform, sample_rate = inf.model.audio(file)
        cls, _, off, _ = inf.slide_device(waveform, sample_rate)
        if hook is not None:
            hook(completed=len(off), total=len(off))
        ctx = inf.model._ctx()
        speech = ctx.powerset_speech(cls)                                           # (C,589,1) f32 on the device
        chunks_sw = SlidingWindow(start=0.0, duration=inf.duration, step=inf.step)
        agg = inf.aggregate_device(SlidingWindowFeature(speech, chunks_sw), inf.model.receptive_field,
                                   warm_up=inf.warm_up, hamming=True, missing=0.0)
        num_samples = waveform.shape[1]
        if (num_samples < ops.CHUNK) or (num_samples - ops.CHUNK) % round(inf.step * sample_rate) > 0:
            agg.data = agg.crop(Segment(0.0, num_samples / sample_rate), mode="loose")
        return agg

    def apply(self, file: AudioFile, hook: Optional[Callable] = None) -> Annotation:
        file = self._segmentation.model.audio.validate_file(file)
        user_hook = hook
        hook = (lambda *a, **k: user_hook(*a, file=file, **k)) if user_hook is not None else (lambda *a, **k: None)
        segmentations = self.speech_scores(file, hook=lambda **k: hook("segmentation", None, **k))
        hook("segmentation", segmentations)
        speech = self._binarize(segmentations)
        speech.uri = file.get("uri")
        return speech.rename_labels({label: "SPEECH" for label in speech.labels()})

    __call__ = apply
