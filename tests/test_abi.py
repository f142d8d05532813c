"""CPU: the C-ABI library loads and exports every symbol include/sgbpo.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

from safegbpo_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def test_header_symbols_exported():
    header = (ROOT / "include" / "sgbpo.h").read_text()
    declared = set(re.findall(r"^\s*(?:const\s+char\*|int)\s+(sgbpo_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed"
    _lib.build()
    handle = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in sgbpo.h but not exported"
    assert declared == set(_lib.EXPORTED), (declared, set(_lib.EXPORTED))


def test_struct_sizes_match_header():
    n, m, hp = _lib.MAXN, _lib.MAXM, _lib.MAXHP
    ints = 10 + 1 + 2
    doubles = 3 * n + 2 * m + 3 * hp + m + n * n + n + n + n * m + 2 * hp + 3 * hp + n
    # natural alignment: int32 blocks are padded to 8 bytes before each double block
    assert ctypes.sizeof(_lib.Consts) >= ints * 4 + doubles * 8
    assert ctypes.sizeof(_lib.StepShared) == 23 * 8


def test_version_and_dims_without_gpu():
    lib = _lib.lib()
    assert b"sm_100a" in lib.sgbpo_version()
    vals = [ctypes.c_int() for _ in range(4)]
    assert lib.sgbpo_env_dims(1, *[ctypes.byref(v) for v in vals]) == 0
    assert [v.value for v in vals] == [4, 1, 5, 0]
    assert lib.sgbpo_env_dims(99, *[ctypes.byref(v) for v in vals]) != 0
    assert b"unknown env_id" in lib.sgbpo_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    monkeypatch.setattr(_lib, "_lib", None)
    try:
        _lib.lib()
    except _lib.SgbpoError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
