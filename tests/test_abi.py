"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/mmc_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mmc_b200.h"


def declared_symbols():
    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmc_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from metropolismontecarlo_b200 import _lib
    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    return _lib


def test_header_symbols_are_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib.LIB_PATH)], capture_output=True, text=True, check=True)
    exported = set(re.findall(r" T (mmc_[a-z0-9_]+)", out.stdout))
    assert set(syms) <= exported, sorted(set(syms) - exported)
    assert set(syms) == set(lib.SIGNATURES), (sorted(set(syms) ^ set(lib.SIGNATURES)))
    L = lib.load()
    assert L.mmc_version() == 100


def test_library_is_sm100a_cuda_not_torch(lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(lib.LIB_PATH)], capture_output=True, text=True)
    assert "sm_100a" in out.stdout
    ldd = subprocess.run(["ldd", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libtorch" not in ldd and "libc10" not in ldd


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = lib.load()
    cfg = lib.Config(0, 0, 1, 0, None)
    h = lib.H()
    rc = L.mmc_create(C.byref(cfg), C.byref(h))
    assert rc == lib.MMC_ECUDA
    assert b"no CPU fallback" in L.mmc_last_error(None)
    from metropolismontecarlo_b200.energy import Engine, MMCError
    with pytest.raises(MMCError):
        Engine()


def test_product_never_imports_oracle():
    pkg = ROOT / "metropolismontecarlo_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.inl")):
        txt = p.read_text()
        assert "oracle" not in txt.lower(), p
