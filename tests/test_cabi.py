"""CPU tests: the C-ABI library loads and exports every symbol include/deepfluids_b200.h declares (no compute calls)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "deepfluids_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    from deepfluids_b200 import cabi
    return cabi


def test_header_and_binding_tables_agree(built):
    assert _declared() == sorted(built.SIGNATURES)


def test_library_exports_every_declared_symbol(built):
    out = subprocess.check_output(["nm", "-D", "--defined-only", built.LIB_PATH]).decode()
    exported = set(re.findall(r" T (dfl_[a-z0-9_]+)", out))
    assert set(_declared()) <= exported
    lib = built.load()
    assert lib.dfl_version() >= 100


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.DflError):
        built.lib()


def test_sass_is_blackwell_native(built):
    """The tensor-core kernels must lower to tcgen05 (UTCHMMA) + TMA (UTMALDG), not legacy HMMA."""
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("no cuobjdump")
    sass = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-sass", built.LIB_PATH]).decode()
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
