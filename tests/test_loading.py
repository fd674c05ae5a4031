"""CPU: local `Pipeline.from_pretrained` plumbing (pyannote_audio_b200/loading.py) -- `expand_subfolders` with the
cases of /root/reference/tests/test_pipeline_subfolder.py, config resolution from a directory / file / dict, and a
community-1 style directory (config.yaml + Lightning-format checkpoints + PLDA npz) resolved and loaded on the host."""
import numpy as np
import pytest
import torch
import yaml

from pyannote_audio_b200 import synthetic as syn
from pyannote_audio_b200.loading import Pipeline, expand_subfolders, get_model, get_plda, is_checkpoint_spec, resolve_pipeline
from pyannote_audio_b200.testing.checkpoints import reference_style_checkpoint as _reference_style_checkpoint


def _expanded(subfolder, model_id="org/repo", revision=None, token=None, cache_dir=None):
    return {"checkpoint": model_id, "revision": revision, "subfolder": subfolder, "token": token, "cache_dir": cache_dir}


def test_expand_subfolders_cases():
    # test_pipeline_subfolder.py: no references / dict reference / parent subfolder / lists / nesting / revisions
    config = {"key": "plain_value", "nested": {"k": 42}}
    expand_subfolders(config, model_id="org/repo")
    assert config == {"key": "plain_value", "nested": {"k": 42}}
    config = {"embedding": "$model/embeddings"}
    expand_subfolders(config, model_id="org/repo", token="tok")
    assert config["embedding"] == _expanded("embeddings", token="tok")
    config = {"segmentation": "$model/seg"}
    expand_subfolders(config, model_id="org/repo", parent_subfolder="pipelines/v1/")
    assert config["segmentation"] == _expanded("pipelines/v1/seg")
    config = {"models": ["$model/a", "plain", {"deep": "$model/b/c@rev2"}], "n": 3}
    expand_subfolders(config, model_id="org/repo", parent_revision="rev1", cache_dir="/tmp/c")
    assert config["models"][0] == _expanded("a", revision="rev1", cache_dir="/tmp/c")
    assert config["models"][1] == "plain" and config["n"] == 3
    assert config["models"][2]["deep"] == _expanded("b/c", revision="rev2", cache_dir="/tmp/c")
    assert is_checkpoint_spec("some/dir") and is_checkpoint_spec(_expanded("x")) and not is_checkpoint_spec({"w": 1})
    assert not is_checkpoint_spec(syn.make_plda(2)) and not is_checkpoint_spec(torch.nn.Linear(1, 1))


@pytest.fixture()
def community_dir(tmp_path):
    root = tmp_path / "community-1"
    for sub, kind in (("segmentation", "seg"), ("embedding", "emb")):
        (root / sub).mkdir(parents=True)
        (root / sub / "pytorch_model.bin").write_bytes(_reference_style_checkpoint(kind)[0])
    (root / "plda").mkdir()
    p = syn.make_plda(2)
    np.savez(root / "plda" / "xvec_transform.npz", mean1=p["mean1"], mean2=p["mean2"], lda=p["lda"])
    np.savez(root / "plda" / "plda.npz", mu=p["mu"], tr=p["tr"], psi=p["psi"])
    config = {"version": "4.0.0",
              "pipeline": {"name": "pyannote.audio.pipelines.SpeakerDiarization",
                           "params": {"clustering": "VBxClustering", "segmentation": "$model/segmentation",
                                      "segmentation_batch_size": 32, "embedding": "$model/embedding",
                                      "embedding_batch_size": 32, "embedding_exclude_overlap": True,
                                      "plda": "$model/plda"}},
              "params": {"clustering": {"threshold": 0.6, "Fa": 0.07, "Fb": 0.8},
                         "segmentation": {"min_duration_off": 0.0}}}
    (root / "config.yaml").write_text(yaml.dump(config))
    return root


def test_resolve_pipeline_from_directory_file_and_dict(community_dir):
    from pyannote_audio_b200.clustering import PLDA
    from pyannote_audio_b200.models import PyanNet, WeSpeakerResNet34
    from pyannote_audio_b200.pipeline import SpeakerDiarization

    for checkpoint in (community_dir, str(community_dir / "config.yaml")):
        klass, params, hyper = resolve_pipeline(checkpoint)
        assert klass is SpeakerDiarization and hyper["clustering"] == {"threshold": 0.6, "Fa": 0.07, "Fb": 0.8}
        assert params["segmentation"]["subfolder"] == "segmentation" and params["embedding_exclude_overlap"] is True
        assert str(params["plda"]["checkpoint"]) == str(community_dir) and params["token"] is None
    # the expanded entries load on the host exactly like the reference's get_model / get_plda
    seg, emb, plda = get_model(params["segmentation"]), get_model(params["embedding"]), get_plda(params["plda"])
    assert isinstance(seg, PyanNet) and isinstance(emb, WeSpeakerResNet34) and isinstance(plda, PLDA)
    assert not seg.training and seg.specifications.powerset and seg.specifications.duration == 10.0
    assert torch.equal(seg.state_dict()["classifier.weight"], syn.make_segmentation_state_dict(0)["classifier.weight"])
    assert np.array_equal(plda.phi, PLDA(syn.make_plda(2)).phi)
    assert get_model(seg) is seg and get_plda(plda) is plda
    with pytest.raises(TypeError):
        get_model(3.14)
    # a config dictionary; subfolder / revision misuse; hub ids; unknown pipelines
    cfg = yaml.safe_load((community_dir / "config.yaml").read_text())
    klass, params, _ = resolve_pipeline(cfg)
    assert klass is SpeakerDiarization and params["segmentation"]["subfolder"] == "segmentation"
    with pytest.raises(ValueError, match="Subfolder cannot be used"):
        resolve_pipeline(cfg, subfolder="x")
    with pytest.raises(ValueError, match="Revisions cannot be used"):
        resolve_pipeline(community_dir, revision="main")
    with pytest.raises(ValueError, match="not a local pipeline checkpoint"):
        resolve_pipeline("pyannote/speaker-diarization-community-1")
    with pytest.raises(ValueError, match="Revisions must be passed"):
        resolve_pipeline("pyannote/speaker-diarization-community-1@main")
    with pytest.raises(NotImplementedError):
        resolve_pipeline({"pipeline": {"name": "pyannote.audio.pipelines.SpeechSeparation"}})
    # config.yaml inside a subfolder of the directory: children resolve under it (parent_subfolder)
    sub = community_dir / "pipelines" / "v2"
    sub.mkdir(parents=True)
    (sub / "config.yaml").write_text(yaml.dump({"pipeline": {"name": "pyannote.audio.pipelines.VoiceActivityDetection",
                                                              "params": {"segmentation": "$model/seg"}}}))
    klass, params, hyper = resolve_pipeline(community_dir, subfolder="pipelines/v2")
    assert klass.__name__ == "VoiceActivityDetection" and params["segmentation"]["subfolder"] == "pipelines/v2/seg"
    assert hyper is None


def test_from_pretrained_builds_the_pipeline_up_to_the_device(community_dir):
    """Everything of Pipeline.from_pretrained runs on the host except moving the models to the GPU: on a machine
    without one the constructor must get that far (models and PLDA loaded) and then fail loudly -- no CPU fallback."""
    if torch.cuda.is_available():
        pipeline = Pipeline.from_pretrained(community_dir)
        assert pipeline.embedding_exclude_overlap is True and pipeline.clustering.threshold == 0.6
        return
    with pytest.raises((RuntimeError, AssertionError)) as err:
        Pipeline.from_pretrained(community_dir)
    assert "NVIDIA" in str(err.value) or "CUDA" in str(err.value) or "cuda" in str(err.value)
