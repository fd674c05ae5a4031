"""Fits the synthetic segmentation classifier (last Linear 128->7) by ridge regression.

Random LSTM weights alone give an (almost) constant powerset class; to get non-degenerate segmentations
(turn taking, overlaps, several local speakers) on the synthetic conversations, the last layer is fitted
in closed form on 100 s of seeded synthetic audio against its known turn structure.  The result is stored
in pyannote_audio_b200/data/synthetic_classifier_seed0.npz and picked up by
``synthetic.make_segmentation_state_dict(seed=0)``.  Uses the CPU oracle as feature extractor (this is a
fixture generator, run once in the build container:  PYTHONPATH=. python tests/golden/make_synthetic_classifier.py).
"""
import numpy as np
import torch

from oracle import nets, pipeline as P
from pyannote_audio_b200 import synthetic as syn

torch.set_num_threads(8)
seg = nets.PyanNet()
seg.load_state_dict(syn.make_segmentation_state_dict(0, fitted_classifier=False))
seg.eval()
mapping = nets.powerset_mapping(3, 2).numpy()


def features_and_targets(dur, seed):
    wav, turns = syn.make_conversation(dur, seed=seed, return_turns=True)
    chunks = P.chunk_waveform(wav)
    feats = []
    with torch.inference_mode():
        for c in range(0, chunks.shape[0], 32):
            x = seg.sincnet(chunks[c:c + 32])
            x, _ = seg.lstm(x.transpose(1, 2))
            for lin in seg.linear:
                x = torch.nn.functional.leaky_relu(lin(x))
            feats.append(x.numpy())
    feats = np.concatenate(feats)
    C = feats.shape[0]
    tgt = np.zeros((C, 589), dtype=np.int64)
    fr_step, fr_dur = 270 / 16000, 991 / 16000
    for c in range(C):
        mid = c * 1.0 + np.arange(589) * fr_step + fr_dur / 2
        act = np.zeros((589, 3), bool)
        for a, b, k in turns:
            act[:, k] |= (mid >= a) & (mid < b)
        first = [(np.argmax(act[:, k]) if act[:, k].any() else 10 ** 9, k) for k in range(3)]
        act = act[:, [k for _, k in sorted(first)]]
        for t in range(589):
            a = act[t].astype(float)
            if a.sum() > 2:
                a[2] = 0
            tgt[c, t] = int(np.argmax((mapping == a).all(1)))
    return feats.reshape(-1, 128), tgt.reshape(-1)


X, y = features_and_targets(100.0, 777)
Y = np.eye(7)[y]
mu = X.mean(0)
Xc = X - mu
W = np.linalg.solve(Xc.T @ Xc + 1e-3 * len(X) * np.eye(128), Xc.T @ (Y - Y.mean(0)))
b = Y.mean(0) - mu @ W
SCALE = 8.0
np.savez("pyannote_audio_b200/data/synthetic_classifier_seed0.npz",
         weight=(SCALE * W.T).astype(np.float32), bias=(SCALE * b).astype(np.float32))
print("train acc", ((X @ W + b).argmax(1) == y).mean())

# ---- embedding bias: centre the synthetic embeddings (see synthetic.make_embedding_state_dict) ----
emb = nets.WeSpeakerResNet34()
sd = syn.make_embedding_state_dict(1, centered=False)
emb.load_state_dict(sd)
emb.eval()
seg.load_state_dict(syn.make_segmentation_state_dict(0))
wav = syn.make_conversation(60.0, seed=777)
s = P.slide(seg, wav)
E = P.get_embeddings(emb, wav, s)
train, _, _ = P.filter_embeddings(E, s.data)
bias = sd["resnet.seg_1.bias"].numpy() - train.mean(0)
np.savez("pyannote_audio_b200/data/synthetic_embedding_bias_seed1.npz", bias=bias.astype(np.float32))
print("embedding centre norm", np.linalg.norm(train.mean(0)), "residual std", (train - train.mean(0)).std())
