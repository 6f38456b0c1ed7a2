"""The C-ABI library loads on a CPU-only box and exports every symbol include/f3ps.h declares;
no compute entry point is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import f3ps
    f3ps.build()
    return f3ps


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "f3ps.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(f3ps_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "libf3ps.so does not export " + n
    assert set(built.EXPORTED) == set(names)


def test_version_and_no_cpu_fallback(built):
    import torch
    L = built.lib()
    assert b"sm_100a" in L.f3ps_version()
    if not torch.cuda.is_available():
        with pytest.raises(built.F3psError):
            built.Segmenter(device=0)      # f3ps_create returns F3PS_ERR_CUDA: the product never falls back to the CPU


def test_library_is_sm100a_only(built):
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % built.LIB_PATH).read()
    if out.strip():
        assert "sm_100a" in out and "sm_90" not in out


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "fast-3d-pointcloud-segmentation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
