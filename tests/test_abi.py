"""CPU: the C-ABI library loads and exports every symbol include/b200diar.h declares (no compute calls)."""
import ctypes
import os
import re

from pyannote_audio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b200diar.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"missing exports: {missing}"


def test_python_binding_covers_header():
    names = set(declared_functions())
    bound = set(_lib._PROTOS)
    assert names <= bound, f"not bound in _lib.py: {sorted(names - bound)}"
    assert bound <= names, f"bound but not declared in the header: {sorted(bound - names)}"


def test_error_path_without_gpu_or_bad_args():
    lib = _lib.load()
    assert lib.b200_version() >= 100
    # host-only entry point: bad arguments -> negative status + message, no exception across the boundary
    rc = lib.b200_fcluster_distance(None, 0, ctypes.c_double(0.5), None)
    assert rc == -1
    assert b"bad arguments" in lib.b200_last_error()


def _plan(off, valid, sub_batch, share):
    import numpy as np

    lib = _lib.load()
    off = np.ascontiguousarray(off, dtype=np.int64)
    valid = np.ascontiguousarray(valid, dtype=np.int32)
    n = len(off)
    frame0 = np.zeros(n, dtype=np.int32)
    rows = np.zeros((n + sub_batch - 1) // sub_batch, dtype=np.int32)
    runs = lib.b200_emb_fbank_plan(off.ctypes.data, valid.ctypes.data, n, sub_batch, int(share), frame0.ctypes.data,
                                   rows.ctypes.data)
    return int(runs), frame0, rows


def test_fbank_plan_shares_frames_of_overlapping_chunks():
    """Host-only entry point: the shared-frame fbank layout (include/b200diar.h: b200_emb_fbank_plan).  A chunk's
    998 frames must map to rows that another chunk maps to only when both read the very same samples."""
    import numpy as np

    CHUNK, STEP, HOP, FR = 160000, 16000, 160, 998
    # a pipeline batch: two files back to back (591 chunks each, the last one short), sub-batches of 256 chunks
    off, valid = [], []
    base = 0
    for _ in range(2):
        o = np.arange(591, dtype=np.int64) * STEP
        v = np.full(591, CHUNK, dtype=np.int32)
        v[-1] = 9600000 - int(o[-1])
        off.append(base + o)
        valid.append(v)
        base += int(o[-1]) + CHUNK
    off, valid = np.concatenate(off), np.concatenate(valid)
    runs, frame0, rows = _plan(off, valid, 256, True)
    runs0, frame00, rows0 = _plan(off, valid, 256, False)
    assert runs0 == len(off) and np.array_equal(frame00, (np.arange(len(off)) % 256) * FR)
    assert np.array_equal(rows0, [256 * FR] * 4 + [(len(off) - 1024) * FR])
    assert rows.sum() * 8 < rows0.sum()                    # ~10x fewer frames
    # every row of every chunk names the right samples: row r of a sub-batch <-> first sample of its frame
    for s in range(len(rows)):
        first_sample = {}
        for c in range(s * 256, min(len(off), (s + 1) * 256)):
            assert 0 <= frame0[c] and frame0[c] + FR <= rows[s]
            for k in (0, 1, 499, 997):
                sample = int(off[c]) + k * HOP
                key = (int(frame0[c]) + k, valid[c] == CHUNK)
                assert first_sample.setdefault(key, sample) == sample
            if valid[c] < CHUNK:                            # a short chunk never shares its rows
                others = [d for d in range(s * 256, min(len(off), (s + 1) * 256)) if d != c]
                assert all(abs(int(frame0[d]) - int(frame0[c])) >= FR for d in others)
    # unaligned, reversed and gapped chunks fall back to private runs; adjacency (gap of exactly 998 hops) extends
    off2 = np.array([0, 16000, 16037, 8000, 16000 + FR * HOP, 10 ** 7], dtype=np.int64)
    runs2, f2, rows2 = _plan(off2, np.full(6, CHUNK, np.int32), 256, True)
    assert list(f2[:2]) == [0, 100] and f2[2] == 100 + FR and f2[3] == f2[2] + FR
    assert f2[4] == f2[3] + FR and f2[5] == f2[4] + FR and runs2 == 5 and rows2[0] == f2[5] + FR
    off3 = np.array([0, FR * HOP], dtype=np.int64)
    runs3, f3, rows3 = _plan(off3, np.full(2, CHUNK, np.int32), 256, True)
    assert runs3 == 1 and list(f3) == [0, FR] and rows3[0] == 2 * FR
    assert _plan(np.array([-1]), np.array([CHUNK]), 4, True)[0] == -1
