"""CPU: the oracle's networks and its SpeakerDiarization.apply restatement against vectors produced by EXECUTING the
reference's own model wrappers (PyanNet wiring, WeSpeakerResNet34.compute_fbank / forward with weights), its Audio.crop,
its VBxClustering and its SpeakerDiarization.apply / get_embeddings / reconstruct on a synthetic conversation
(tests/golden/make_golden_apply.py -> reference_apply_vectors.npz; the ParamSincFB filters and the four pyannote.core
classes behind them are the oracle's / a stand-in: the two pieces whose sources are not available)."""
import os

import numpy as np
import pytest
import torch

from oracle import nets, pipeline as P
from pyannote_audio_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHUNKS = P.SW(0.0, 10.0, 1.0)


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_apply_vectors.npz"))


@pytest.fixture(scope="module")
def models():
    seg, emb = nets.PyanNet(), nets.WeSpeakerResNet34()
    seg.load_state_dict(syn.make_segmentation_state_dict(0))
    emb.load_state_dict(syn.make_embedding_state_dict(1))
    return seg.eval(), emb.eval()


@pytest.fixture(scope="module")
def wav(ref):
    return syn.make_conversation(float(ref["wav_seconds"]), seed=int(ref["wav_seed"]))


def test_pyannet_matches_the_reference_module(ref, models, wav):
    seg_model, _ = models
    with torch.inference_mode():
        logp = seg_model(P.chunk_waveform(wav)).numpy()
    assert logp.shape == ref["logp"].shape == (15, 589, 7)
    np.testing.assert_allclose(logp, ref["logp"], rtol=0, atol=2e-5)
    seg = P.slide(seg_model, wav)
    assert np.array_equal(seg.data.astype(np.uint8), ref["segmentations"])
    count = P.speaker_count(seg, P.SW(*nets.sincnet_receptive_field()), warm_up=(0.0, 0.0))
    assert np.array_equal(count.data, ref["count"]) and ref["count"].max() == 2


@pytest.mark.parametrize("name,exclude_overlap", [("std", False), ("xo", True)])
def test_get_embeddings_matches_the_reference_pipeline(ref, models, wav, name, exclude_overlap):
    # the reference: Audio.crop(mode="pad") per chunk, one WeSpeakerResNet34 forward per (chunk, speaker) with the
    # (clean) mask as pooling weights, batches of 32 (speaker_diarization.py:332-478)
    _, emb_model = models
    seg = P.SWF(ref["segmentations"].astype(np.float32), CHUNKS)
    want = ref[f"{name}_embeddings"]
    got = P.get_embeddings(emb_model, wav, seg, exclude_overlap=exclude_overlap, share_trunk=False)
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)
    shared = P.get_embeddings(emb_model, wav, seg, exclude_overlap=exclude_overlap, share_trunk=True)
    cos = (shared * want).sum(-1) / (np.linalg.norm(shared, axis=-1) * np.linalg.norm(want, axis=-1))
    assert (1 - cos).max() < 1e-6                            # one trunk pass + three poolings = three forwards
    if exclude_overlap:
        assert np.abs(want - ref["std_embeddings"]).max() > 1e-3      # the clean masks were actually used


@pytest.mark.parametrize("name,exclude_overlap", [("std", False), ("xo", True)])
def test_apply_matches_the_reference_apply(ref, name, exclude_overlap):
    plda = P.PLDA(**syn.make_plda(2))
    seg = P.SWF(ref["segmentations"].astype(np.float32), CHUNKS)
    out = P.apply(None, None, plda, None, segmentations=seg, embeddings=ref[f"{name}_embeddings"],
                  exclude_overlap=exclude_overlap)
    want = ref[f"{name}_discrete"]
    assert out.discrete.data.shape == want.shape
    differ = np.nonzero((out.discrete.data != want).any(axis=1))[0]
    assert len(differ) == 0, "discrete diarization differs from the reference's run (no activation ties in this file)"
    for key, times in (("diar", out.times), ("excl", out.exclusive_times)):
        rows = ref[f"{name}_{key}"]
        assert len(times) == len(rows) and len(rows) > 0
        np.testing.assert_array_equal(np.array([[a, b] for a, b, _ in times]), rows[:, :2])
        assert [lab for _, _, lab in times] == [f"SPEAKER_{int(k):02d}" for k in rows[:, 2]]
    assert out.labels == list(ref[f"{name}_labels"])
    np.testing.assert_allclose(out.speaker_embeddings, ref[f"{name}_speaker_embeddings"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["vad", "vad_short"])
def test_vad_matches_the_reference_pipeline(ref, models, name):
    # the reference: Inference.slide verbatim (pre_aggregation_hook = max over speakers, Hamming overlap-add, padded
    # tail cropped; core/inference.py:217-373) and VoiceActivityDetection.apply (Binarize at 0.5, min_duration_on)
    seg_model, _ = models
    seconds = {"vad": float(ref["wav_seconds"]), "vad_short": 6.3}[name]
    wav = syn.make_conversation(seconds, seed=int(ref["wav_seed"]))
    scores = P.vad_scores(seg_model, wav)
    want = ref[f"{name}_scores"]
    assert scores.data.shape == want.shape and scores.data.dtype == want.dtype
    np.testing.assert_allclose(scores.data, want, rtol=0, atol=1e-6)
    np.testing.assert_allclose([scores.sw.start, scores.sw.duration, scores.sw.step], ref[f"{name}_scores_sw"], atol=1e-15)
    for sub, mon in (("", 0.0), ("_on", 0.25)):
        got = P.binarize_scores(P.SWF(want, scores.sw), onset=0.5, offset=0.5, min_duration_on=mon)
        rows = ref[f"{name}{sub}_rows"]
        assert len(got) == len(rows) and len(rows) >= 1
        np.testing.assert_array_equal(np.array([[a, b] for a, b, _ in got]), rows)


@pytest.mark.parametrize("name,kw", [("forced3", dict(num_speakers=3)), ("max1", dict(max_speakers=1)),
                                     ("min3", dict(min_speakers=3))])
def test_apply_with_speaker_bounds_matches_the_reference(ref, name, kw):
    # forced / minimum number of speakers -> KMeans on the normalised training embeddings (clustering.py:626-642),
    # max_speakers=1 -> the instantaneous count is capped (speaker_diarization.py:735)
    plda = P.PLDA(**syn.make_plda(2))
    seg = P.SWF(ref["segmentations"].astype(np.float32), CHUNKS)
    out = P.apply(None, None, plda, None, segmentations=seg, embeddings=ref["std_embeddings"], **kw)
    want = ref[f"{name}_discrete"]
    assert out.discrete.data.shape == want.shape
    differ = np.nonzero((out.discrete.data != want).any(axis=1))[0]
    if len(differ):                                         # only activation ties may differ (numpy's default argsort)
        act = P.aggregate(P.clustered_segmentations(seg, out.hard_clusters), out.count.sw, hamming=False, missing=0.0,
                          skip_average=True).data[: len(want)]
        np.testing.assert_array_equal(out.discrete.data.sum(axis=1), want.sum(axis=1))
        for t in differ:
            assert sorted(act[t][out.discrete.data[t] > 0]) == sorted(act[t][want[t] > 0]), f"frame {t}: not a tie"
        print(f"[apply {name}] {len(differ)} of {len(want)} frames differ from the reference's run, all of them ties")
    else:
        rows = ref[f"{name}_diar"]
        np.testing.assert_array_equal(np.array([[a, b] for a, b, _ in out.times]), rows[:, :2])
        assert [lab for _, _, lab in out.times] == [f"SPEAKER_{int(k):02d}" for k in rows[:, 2]]
    assert out.labels == list(ref[f"{name}_labels"])
    np.testing.assert_allclose(out.speaker_embeddings, ref[f"{name}_speaker_embeddings"], rtol=0, atol=1e-6)


def test_apply_on_a_silent_file_matches_the_reference(ref):
    plda = P.PLDA(**syn.make_plda(2))
    out = P.apply(None, None, plda, None, segmentations=P.SWF(np.zeros((15, 589, 3), dtype=np.float32), CHUNKS))
    assert out.segments == [] and tuple(out.speaker_embeddings.shape) == tuple(ref["silent_speaker_embeddings_shape"])
