"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/aznet_b200.h declares (no compute calls without a GPU); argument errors surface as
Python exceptions; the product never imports the oracle."""
import os
import re
import subprocess
import sys

import pytest

from aznet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    hdr = open(_lib.HEADER).read()
    declared = set(re.findall(r"\b(azn_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert declared == set(_lib.EXPORTS)


def test_version_and_error_text(lib):
    assert b"sm_100a" in lib.azn_version()
    assert lib.azn_last_error() is not None


def test_sass_is_blackwell_native():
    """tcgen05 / TMA evidence in the shipped SASS (B200_PROFILING.md: UTC*MMA, UTMALDG, LDTM)."""
    so = _lib.build()
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100a" in sass.upper() or "EF_CUDA_SM100" in sass
    assert re.search(r"UTC[A-Z]*MMA", sass), "no tcgen05.mma (UTC*MMA) in SASS"
    assert "UTMALDG" in sass, "no TMA tensor load (UTMALDG) in SASS"
    assert "LDTM" in sass, "no tcgen05.ld (LDTM) in SASS"


def test_struct_mirror_size():
    """ctypes mirror of azn_search_state matches the C layout (compile a sizeof probe with gcc)."""
    src = '#include "%s"\n#include <stdio.h>\nint main(){printf("%%zu", sizeof(azn_search_state));return 0;}' % _lib.HEADER
    exe = "/tmp/azn_sizeof"
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    size = int(subprocess.run([exe], capture_output=True, text=True).stdout)
    import ctypes
    assert ctypes.sizeof(_lib.SearchState) == size


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aznet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), "%s imports the oracle" % f
                assert "liboracle" not in txt and "/root/reference" not in txt.replace("relative to /root/reference", "")


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "SO_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device()
    from aznet_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.nms(torch.zeros((4, 5)), 0.5)
