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
