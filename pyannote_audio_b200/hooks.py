"""Pipeline hooks (mirror of /root/reference/src/pyannote/audio/pipelines/utils/hook.py:37-239).

``hook(step_name, step_artifact, file=..., total=..., completed=...)`` is called by ``SpeakerDiarization`` with the
step names "segmentation", "speaker_counting", "embeddings", "discrete_diarization", with real
SlidingWindowFeature / ndarray artifacts (copied device -> host only when a hook is installed) and the reference's
progress calls (artifact None, total / completed set).
"""
from __future__ import annotations

import time
from copy import deepcopy
from typing import Any, Mapping, Optional, Text


class ArtifactHook:
    """Stores every step artifact in file[file_key][step_name] (hook.py:37-82)."""

    def __init__(self, *artifacts, file_key: str = "artifact"):
        self.artifacts = artifacts
        self.file_key = file_key

    def __enter__(self):
        return self

    def __exit__(self, *args):
        pass

    def __call__(self, step_name: Text, step_artifact: Any, file: Optional[Mapping] = None,
                 total: Optional[int] = None, completed: Optional[int] = None):
        if (step_artifact is None) or (self.artifacts and step_name not in self.artifacts):
            return
        file.setdefault(self.file_key, dict())[step_name] = deepcopy(step_artifact)


class TimingHook:
    """Wall-clock per step in file[file_key][step_name] (hook.py:150-203)."""

    def __init__(self, file_key: str = "timing"):
        self.file_key = file_key

    def __enter__(self):
        self._pipeline_start_time = time.time()
        self._start_time = dict()
        self._end_time = dict()
        return self

    def __exit__(self, *args):
        _pipeline_end_time = time.time()
        processing_time = dict()
        processing_time["total"] = _pipeline_end_time - self._pipeline_start_time
        for step_name, _start_time in self._start_time.items():
            _end_time = self._end_time[step_name]
            processing_time[step_name] = _end_time - _start_time
        self._file[self.file_key] = processing_time

    def __call__(self, step_name: Text, step_artifact: Any, file: Optional[Mapping] = None,
                 total: Optional[int] = None, completed: Optional[int] = None):
        if not hasattr(self, "_file"):
            self._file = file
        if completed is None:
            return
        if completed == 0:
            self._start_time[step_name] = time.time()
        if completed >= total:
            self._end_time[step_name] = time.time()


class Hooks:
    """Combines several hooks (hook.py:206-239)."""

    def __init__(self, *hooks):
        self.hooks = hooks

    def __enter__(self):
        for hook in self.hooks:
            if hasattr(hook, "__enter__"):
                hook.__enter__()
        return self

    def __exit__(self, *args):
        for hook in self.hooks:
            if hasattr(hook, "__exit__"):
                hook.__exit__(*args)

    def __call__(self, step_name: Text, step_artifact: Any, file: Optional[Mapping] = None,
                 total: Optional[int] = None, completed: Optional[int] = None):
        for hook in self.hooks:
            hook(step_name, step_artifact, file=file, total=total, completed=completed)
