"""Generates tests/golden/reference_apply_vectors.npz by EXECUTING the reference's model wrappers and its
SpeakerDiarization.apply where they lie under /root/reference -- run once in the build container:

    PYTHONPATH=. python tests/golden/make_golden_apply.py

On top of tests/golden/make_golden_pipeline.py (import stubs + pyannote.core stand-in) this loads
  models/blocks/sincnet.py, models/segmentation/PyanNet.py      (real wiring; `asteroid_filterbanks` = the oracle's
                                                                 ParamSincFB / Encoder restatement, which stays unpinned)
  models/embedding/wespeaker/__init__.py (+ resnet.py, pooling.py)   compute_fbank, forward(waveforms, weights)
  core/io.py                                                    Audio.crop of in-memory waveforms (mode="pad")
  pipelines/speaker_diarization.py                              SpeakerDiarization.apply / get_embeddings / reconstruct
behind a minimal stand-in for `pyannote.audio.core.model.Model` (an nn.Module with lightning's save_hyperparameters
and the LogSoftmax activation of a powerset model), loads the seeded synthetic weights into the reference's own modules,
and runs them on a synthetic conversation.  The reference's `apply` is executed verbatim on an instance whose
collaborators are: `_segmentation` = the reference PyanNet slid over the file (chunking as core/inference.py:235-257,
powerset -> multilabel by the reference's Powerset), `_embedding` = the reference WeSpeakerResNet34 called like
PyannoteAudioPretrainedSpeakerEmbedding.__call__ (pipelines/speaker_verification.py:704-716), `_audio` = the reference
Audio, `clustering` = the reference VBxClustering with the reference PLDA.
"""
import copy
import inspect
import os
import sys
import tempfile
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden_pipeline as G  # noqa: E402

from oracle import nets  # noqa: E402
from pyannote_audio_b200 import synthetic as syn  # noqa: E402


class Model(torch.nn.Module):
    """Stand-in for pyannote.audio.core.model.Model (a LightningModule): just what the two architectures use."""

    def __init__(self, sample_rate=16000, num_channels=1, task=None):
        super().__init__()
        self.hparams = types.SimpleNamespace(sample_rate=sample_rate, num_channels=num_channels)
        self.specifications = None

    def save_hyperparameters(self, *names):
        caller = inspect.currentframe().f_back.f_locals
        for n in names:
            setattr(self.hparams, n, caller[n])

    def default_activation(self):
        return torch.nn.LogSoftmax(dim=-1)                  # core/model.py: mono-label (powerset) problems


def load_models(ref):
    sys.modules["pyannote.audio.core.model"].Model = Model
    ast = G.stub("asteroid_filterbanks")
    ast.Encoder, ast.ParamSincFB = nets.Encoder, nets.ParamSincFB
    G.stub("pyannote.audio.models"), G.stub("pyannote.audio.models.blocks"), G.stub("pyannote.audio.models.segmentation")
    G.stub("pyannote.audio.models.embedding")
    out = {}
    out["receptive_field"] = G.load("pyannote.audio.utils.receptive_field", "utils/receptive_field.py")
    out["params"] = G.load("pyannote.audio.utils.params", "utils/params.py")
    out["pooling"] = G.load("pyannote.audio.models.blocks.pooling", "models/blocks/pooling.py")
    out["sincnet"] = G.load("pyannote.audio.models.blocks.sincnet", "models/blocks/sincnet.py")
    out["pyannet"] = G.load("pyannote.audio.models.segmentation.PyanNet", "models/segmentation/PyanNet.py")
    import importlib.util

    name, d = "pyannote.audio.models.embedding.wespeaker", os.path.join(G.SPA, "models", "embedding", "wespeaker")
    spec = importlib.util.spec_from_file_location(name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    out["wespeaker"] = mod
    out["io"] = G.load("pyannote.audio.core.io", "core/io.py")
    return out


def main():
    ref = G.load_reference()
    ref.update(load_models(ref))
    out = {}
    torch.manual_seed(0)
    seg_sd, emb_sd, plda_arrays = syn.make_segmentation_state_dict(0), syn.make_embedding_state_dict(1), syn.make_plda(2)

    # ---- the reference's PyanNet and WeSpeakerResNet34 with the synthetic weights ---------------------------------
    seg_model = ref["pyannet"].PyanNet(lstm={"num_layers": 4})          # community-1 hyper-parameters (4 BiLSTM layers)
    seg_model.specifications = types.SimpleNamespace(powerset=True, num_powerset_classes=7, classes=["a", "b", "c"],
                                                     powerset_max_classes=2, duration=10.0)
    seg_model.build()
    missing = seg_model.load_state_dict(seg_sd, strict=True)
    seg_model.eval()
    emb_model = ref["wespeaker"].WeSpeakerResNet34()
    emb_model.load_state_dict(emb_sd, strict=True)
    emb_model.eval()
    print("state dicts loaded strictly into the reference modules:", missing)

    seconds, seed = 23.4, 77
    wav = syn.make_conversation(seconds, seed=seed)
    out["wav_seconds"], out["wav_seed"] = np.array(seconds), np.array(seed)
    file = {"waveform": wav, "sample_rate": 16000, "uri": "golden"}
    audio = ref["io"].Audio(sample_rate=16000, mono="downmix")
    frames = G.SlidingWindow(start=0.0, **G.FRAMES)
    chunks_sw = G.SlidingWindow(start=0.0, duration=10.0, step=1.0)
    powerset = ref["powerset"].Powerset(3, 2)

    # the reference's Inference.slide / infer (core/inference.py:182-373) run verbatim on a hand-assembled instance
    # (Inference.__init__ needs the Lightning model API: device moves, example outputs, ...)
    Inference = ref["inference"].Inference
    Spec = sys.modules["pyannote.audio.core.model"].Specifications
    spec = Spec()
    spec.powerset, spec.num_powerset_classes, spec.classes, spec.powerset_max_classes = True, 7, ["a", "b", "c"], 2
    spec.permutation_invariant, spec.duration = True, 10.0
    spec.resolution = sys.modules["pyannote.audio.core.task"].Resolution.FRAME
    seg_model.specifications, seg_model.audio, seg_model.receptive_field = spec, audio, frames

    def make_inference(**kw):
        inf = object.__new__(Inference)
        inf.model, inf.duration, inf.step, inf.batch_size, inf.device = seg_model, 10.0, 1.0, 4, torch.device("cpu")
        inf.conversion, inf.warm_up = powerset, (0.0, 0.0)
        inf.skip_aggregation, inf.pre_aggregation_hook = kw.get("skip_aggregation", False), kw.get("pre_aggregation_hook")
        return inf

    inf_skip = make_inference(skip_aggregation=True)

    def slide(f, hook=None):
        waveform, sr = audio(f)
        return Inference.slide(inf_skip, waveform, sr, hook=hook)

    SD = ref["speaker_diarization"].SpeakerDiarization
    art = {}

    def hook(name, artifact, **kw):
        if artifact is not None:
            art[name] = copy.deepcopy(artifact)             # the last call of a step carries its final artifact
                                                            # (a copy: apply() caps `count` in place afterwards)

    class _Emb:
        sample_rate, dimension, min_num_samples = 16000, 256, 400

        def __call__(self, waveforms, masks=None):
            with torch.inference_mode():
                return emb_model(waveforms, weights=masks).numpy()

    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, "xvec_transform.npz"), mean1=plda_arrays["mean1"], mean2=plda_arrays["mean2"],
                 lda=plda_arrays["lda"])
        np.savez(os.path.join(td, "plda.npz"), mu=plda_arrays["mu"], tr=plda_arrays["tr"], psi=plda_arrays["psi"])
        plda = ref["plda"].PLDA(os.path.join(td, "xvec_transform.npz"), os.path.join(td, "plda.npz"))
    for name, exclude_overlap in (("std", False), ("xo", True)):
        sd = object.__new__(SD)
        sd.training, sd.legacy, sd.klustering, sd._expects_num_speakers = False, False, "VBxClustering", False
        sd.embedding_exclude_overlap, sd.embedding_batch_size = exclude_overlap, 32
        sd._segmentation = types.SimpleNamespace(
            __call__=None, model=types.SimpleNamespace(specifications=seg_model.specifications, receptive_field=frames))
        seg_callable = type("_Seg", (), {"model": sd._segmentation.model, "__call__": staticmethod(slide)})()
        sd._segmentation = seg_callable
        sd.segmentation = types.SimpleNamespace(min_duration_off=0.0, threshold=0.5)
        sd._embedding, sd._audio = _Emb(), audio
        cl = ref["clustering"].VBxClustering(plda)
        cl.threshold, cl.Fa, cl.Fb = 0.6, 0.07, 0.8
        sd.clustering = cl
        sd.setup_hook = lambda f, hook=None: hook
        art.clear()
        res = SD.apply(sd, dict(file), hook=hook)
        if name == "std":
            out["segmentations"] = art["segmentation"].data.astype(np.uint8)
            out["count"] = art["speaker_counting"].data
        out[f"{name}_embeddings"] = art["embeddings"]
        out[f"{name}_discrete"] = art["discrete_diarization"].data
        for key, ann in (("diar", res.speaker_diarization), ("excl", res.exclusive_speaker_diarization)):
            rows = [(s.start, s.end, int(str(lab).split("_")[1])) for s, _, lab in ann.itertracks(yield_label=True)]
            out[f"{name}_{key}"] = np.array(rows, dtype=np.float64).reshape(-1, 3)
        import json

        out[f"{name}_serialized"] = np.array(json.dumps(res.serialize()))        # DiarizeOutput.serialize verbatim
        out[f"{name}_labels"] = np.array(res.speaker_diarization.labels())
        out[f"{name}_speaker_embeddings"] = res.speaker_embeddings
        print(name, "segments", len(out[f"{name}_diar"]), "exclusive", len(out[f"{name}_excl"]), "labels",
              out[f"{name}_labels"], "embeddings", art["embeddings"].shape, "nan", int(np.isnan(art["embeddings"]).sum()),
              "max count", int(art["speaker_counting"].data.max()))
    # ---- the same file with constraints on the number of speakers (KMeans branch, count capped by max_speakers) and a
    # file on which no speaker is ever active (early exit, speaker_diarization.py:617-628) ----------------------------
    for name, kw in (("forced3", dict(num_speakers=3)), ("max1", dict(max_speakers=1)), ("min3", dict(min_speakers=3))):
        sd.embedding_exclude_overlap = False
        art.clear()
        import warnings as _w

        with _w.catch_warnings():
            _w.simplefilter("ignore")
            res = SD.apply(sd, dict(file), hook=hook, **kw)
        out[f"{name}_discrete"] = art["discrete_diarization"].data
        for key, ann in (("diar", res.speaker_diarization), ("excl", res.exclusive_speaker_diarization)):
            rows = [(sg.start, sg.end, int(str(lab).split("_")[1])) for sg, _, lab in ann.itertracks(yield_label=True)]
            out[f"{name}_{key}"] = np.array(rows, dtype=np.float64).reshape(-1, 3)
        out[f"{name}_labels"] = np.array(res.speaker_diarization.labels())
        out[f"{name}_speaker_embeddings"] = res.speaker_embeddings
        print(name, "segments", len(out[f"{name}_diar"]), "labels", out[f"{name}_labels"], res.speaker_embeddings.shape)
    silent = object.__new__(SD)
    silent.__dict__.update(sd.__dict__)
    silent._segmentation = type("_Seg", (), {"model": sd._segmentation.model, "__call__": staticmethod(
        lambda f, hook=None: G.SlidingWindowFeature(np.zeros((15, 589, 3), dtype=np.float32), chunks_sw))})()
    res = SD.apply(silent, dict(file), hook=hook)
    assert len(list(res.speaker_diarization.itertracks())) == 0 and res.speaker_embeddings.shape == (0, 256)
    out["silent_speaker_embeddings_shape"] = np.array(res.speaker_embeddings.shape)

    with torch.inference_mode():
        w16 = audio(file)[0]
        ch = w16.unfold(1, 160000, 16000).permute(1, 0, 2)
        last = torch.nn.functional.pad(w16[:, ch.shape[0] * 16000:], (0, 160000 - (w16.shape[1] - ch.shape[0] * 16000)))
        out["logp"] = torch.vstack([seg_model(ch), seg_model(last[None])]).numpy()

    # ---- VoiceActivityDetection.apply (pipelines/voice_activity_detection.py:66-204) verbatim ------------------------
    G.stub("pyannote.metrics.detection")
    vad_mod = G.load("pyannote.audio.pipelines.voice_activity_detection", "pipelines/voice_activity_detection.py")
    VAD = vad_mod.VoiceActivityDetection
    inf_vad = make_inference(pre_aggregation_hook=lambda scores: np.max(scores, axis=-1, keepdims=True))
    for name, seconds in (("vad", 23.4), ("vad_short", 6.3)):
        wav_v = syn.make_conversation(seconds, seed=seed)
        f = {"waveform": wav_v, "sample_rate": 16000, "uri": "golden"}
        vad = object.__new__(VAD)
        vad.training = False
        vad._segmentation = type("_Seg", (), {"__call__": staticmethod(
            lambda ff, hook=None: Inference.slide(inf_vad, *audio(ff), hook=None))})()
        vad.setup_hook = lambda ff, hook=None: (lambda *a, **k: None)
        for sub, (mon, moff) in (("", (0.0, 0.0)), ("_on", (0.25, 0.0))):
            vad._binarize = ref["signal"].Binarize(onset=0.5, offset=0.5, min_duration_on=mon, min_duration_off=moff)
            speech = VAD.apply(vad, dict(f))
            out[f"{name}{sub}_rows"] = np.array([(sg.start, sg.end) for sg, _ in speech.itertracks()],
                                                dtype=np.float64).reshape(-1, 2)
            assert speech.labels() in ([], ["SPEECH"])
        scores = Inference.slide(inf_vad, *audio(f), hook=None)
        out[f"{name}_scores"] = scores.data
        out[f"{name}_scores_sw"] = np.array([scores.sliding_window.start, scores.sliding_window.duration,
                                             scores.sliding_window.step])
        print(name, "frames", scores.data.shape, "speech regions", len(out[f"{name}_rows"]), len(out[f"{name}_on_rows"]))
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_apply_vectors.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}: {len(out)} arrays, {os.path.getsize(dst) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
