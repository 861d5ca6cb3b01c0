"""The C-ABI library loads and exports exactly the entry points include/mixstage_b200.h
declares, and the ctypes prototype table covers them all (no compute calls: CPU box)."""
import ctypes
import os
import re

from mixstage_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mixstage_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(ms_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build(verbose=False)
    lib = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    assert lib.ms_version() >= 100


def test_prototype_table_matches_header():
    syms = _header_symbols()
    assert sorted(_lib.PROTOTYPES) == syms
    src = open(os.path.join(ROOT, "include", "mixstage_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in _lib.PROTOTYPES.items():
        m = re.search(r"\bint\s+%s\s*\((.*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = [p.strip() for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_lib.MixStageError):
        _lib.load()
