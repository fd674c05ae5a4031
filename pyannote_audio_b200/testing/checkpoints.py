"""Test fixture: a Lightning-format ``pytorch_model.bin`` laid out as the reference writes it
(/root/reference/src/pyannote/audio/core/model.py:244-256), built without lightning / pyannote.audio."""
import torch

from . import synthetic as syn


def reference_style_checkpoint(kind):
    """A Lightning-format pytorch_model.bin as the reference writes it (model.py:244-256): state_dict +
    hyper_parameters + checkpoint["pyannote.audio"] whose `specifications` is pickled under the REFERENCE's module
    path pyannote.audio.core.task (registered here only while pickling, then removed again)."""
    import dataclasses
    import enum
    import io
    import sys
    import types

    names = ("pyannote", "pyannote.audio", "pyannote.audio.core", "pyannote.audio.core.task")
    saved = {n: sys.modules.get(n) for n in names}
    mods = {n: types.ModuleType(n) for n in names}
    sys.modules.update(mods)
    try:
        class Problem(enum.Enum):
            BINARY_CLASSIFICATION = 0
            MONO_LABEL_CLASSIFICATION = 1
            MULTI_LABEL_CLASSIFICATION = 2
            REPRESENTATION = 3
            REGRESSION = 4

        class Resolution(enum.Enum):
            FRAME = 1
            CHUNK = 2

        @dataclasses.dataclass
        class Specifications:
            problem: Problem
            resolution: Resolution
            duration: float
            min_duration: float = None
            warm_up: tuple = (0.0, 0.0)
            classes: list = None
            powerset_max_classes: int = None
            permutation_invariant: bool = False

        for c in (Problem, Resolution, Specifications):
            c.__module__, c.__qualname__ = "pyannote.audio.core.task", c.__name__
            setattr(mods["pyannote.audio.core.task"], c.__name__, c)
        if kind == "seg":
            ck = {"state_dict": syn.make_segmentation_state_dict(0),
                  "hyper_parameters": {"sincnet": {"stride": 10}, "linear": {"hidden_size": 128, "num_layers": 2},
                                       "lstm": {"hidden_size": 128, "num_layers": 4, "bidirectional": True,
                                                "monolithic": True, "dropout": 0.0},
                                       "sample_rate": 16000, "num_channels": 1},
                  "pyannote.audio": {"versions": {"pyannote.audio": "4.0.0"},
                                     "architecture": {"module": "pyannote.audio.models.segmentation.PyanNet",
                                                      "class": "PyanNet"},
                                     "specifications": Specifications(
                                         Problem.MONO_LABEL_CLASSIFICATION, Resolution.FRAME, 10.0,
                                         classes=["speaker#1", "speaker#2", "speaker#3"], powerset_max_classes=2,
                                         permutation_invariant=True)}}
        else:
            ck = {"state_dict": syn.make_embedding_state_dict(1),
                  "hyper_parameters": {"sample_rate": 16000, "num_channels": 1, "num_mel_bins": 80,
                                       "frame_length": 25, "frame_shift": 10, "dither": 0.0,
                                       "window_type": "hamming", "use_energy": False},
                  "pyannote.audio": {"versions": {"pyannote.audio": "4.0.0"},
                                     "architecture": {"module": "pyannote.audio.models.embedding.wespeaker",
                                                      "class": "WeSpeakerResNet34"},
                                     "specifications": Specifications(Problem.REPRESENTATION, Resolution.CHUNK, 10.0)}}
        ck["pytorch-lightning_version"] = "2.6.1"
        buf = io.BytesIO()
        torch.save(ck, buf)
        return buf.getvalue(), ck["state_dict"]
    finally:
        for n in names:
            if saved[n] is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = saved[n]
