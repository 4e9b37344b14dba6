"""The C ABI without a GPU: the library loads, exports exactly the symbols include/ort_b200.h declares,
struct layouts match, and compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ort_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ort_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(ort):
    L = ort._lib.load()
    decl = _declared()
    assert decl == sorted(ort._lib.SYMBOLS), "binding symbol list out of sync with include/ort_b200.h"
    for name in decl:
        assert hasattr(L, name), f"libort_b200.so does not export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", ort._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == decl, "the library exports symbols the header does not declare (or vice versa)"


def test_struct_layouts(ort):
    assert C.sizeof(ort._lib.Field) == 80
    assert C.sizeof(ort._lib.Opts) == 48
    assert C.sizeof(ort._lib.Stats) == 112 == ort.STATS_DTYPE.itemsize == ort.STATS_BYTES
    assert C.sizeof(ort._lib.GridOut) == 11 * C.sizeof(C.c_void_p)
    assert ort._lib.load().ort_version() == 200


def test_header_compiles_as_c():
    """the boundary is plain C: the header must compile with a C compiler"""
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_gpu_fails_loudly(ort):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ort.OrtError) as ei:
        ort.Context(0)
    assert "no CPU fallback" in str(ei.value)
    assert ort._lib.load().ort_sync(None) == ort._lib.ORT_EINVAL        # NULL context is an error, not a crash


def test_product_never_imports_oracle():
    """the product path must not route through the oracle"""
    pkg = os.path.join(ROOT, "opticalraytracing.jl_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn
                assert "ort_oracle" not in txt, fn
