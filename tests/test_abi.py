"""The C-ABI library loads, exports every symbol include/noahmp_b200.h declares, and the ctypes mirrors of
its structs have the library's sizes.  No GPU needed (no compute calls)."""
import ctypes as C
import os
import re

from noahmp_b200 import _capi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "noahmp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(noahmp_b200_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    L = C.CDLL(_lib.SO_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_struct_sizes_match_library(built):
    L = _lib.lib()
    assert L.noahmp_b200_sizeof_tables() == C.sizeof(_capi.NoahmpTables)
    assert L.noahmp_b200_sizeof_args() == C.sizeof(_capi.NoahmpLsmArgs)
    assert L.noahmp_b200_sizeof_init_args() == C.sizeof(_capi.NoahmpInitArgs)


def test_no_cpu_fallback(built):
    """Without a CUDA device create() must fail loudly, never compute on the CPU."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from noahmp_b200 import NoahMP, NoahmpError, tables
    with pytest.raises(NoahmpError):
        NoahMP(tables.default_tables("USGS"), 4, 4)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under noahmp_b200/ may reference it."""
    pkg = os.path.join(ROOT, "noahmp_b200")
    bad = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|#\s*include\s*[\"<][^\n]*(oracle|nmo)|CDLL\([^\n]*nmo|dlopen\([^\n]*nmo|-lnmo",
                             txt, flags=re.M):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
