"""CPU-only checks of the C-ABI shared library: it builds for sm_100a, loads, and exports every symbol that
include/trpx.h declares (no compute call is made without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__ as ge
    ge.build()
    from trphysx_b200 import _native
    return _native.LIB_PATH


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "trpx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trpx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(libpath):
    syms = _declared_symbols()
    assert len(syms) >= 28
    lib = ctypes.CDLL(libpath)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/trpx.h but not exported by libtrpx.so"


def test_binding_table_covers_header(libpath):
    from trphysx_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared_symbols()
    lib = _native.lib()
    assert lib.trpx_version() == 100
    assert lib.trpx_last_error() is not None


def test_library_is_sm100a_only_and_has_tma(libpath):
    out = subprocess.run(["cuobjdump", "--list-elf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    sass = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass                 # cp.async.bulk (TMA bulk copy) staging the weights
    assert "LDG.E.ENL2.256" in sass         # 256-bit KV window loads


def test_no_oracle_import_in_product():
    """The product package must not import the oracle (or the reference shim)."""
    pkg = os.path.join(ROOT, "transformer-physx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f"{f} imports the oracle"
                assert "ref_shim" not in src


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from trphysx_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_native.TrpxError):
        _native.lib()
