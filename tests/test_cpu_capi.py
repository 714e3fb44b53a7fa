"""CPU checks of the boundary: libvcgpu.so loads and exports every symbol include/vcgpu.h declares;
without a GPU it fails loudly (there is no CPU fallback); struct layouts match the header."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "vcgpu.h")).read()
    return sorted(set(re.findall(r"\b(vcgpu_[a-z0-9_]+)\s*\(", txt)) - {"vcgpu_iter_cb"})


def test_library_exports_every_declared_symbol():
    from vicalib_b200 import capi

    L = capi.lib()
    syms = _header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == syms


def test_struct_sizes_match_header():
    from vicalib_b200 import capi

    assert C.sizeof(capi.Flags) == 7 * 4 + 4 + 2 * 8      # 7 ints (+pad) + 2 doubles
    assert C.sizeof(capi.Options) == 4 + 4 + 4 * 8 + 4 * 4  # int, pad, 4 doubles, 4 ints
    assert C.sizeof(capi.Summary) == 4 * 4 + 3 * 8 + 8
    assert C.sizeof(capi.Iteration) == 2 * 4 + 7 * 8


def test_no_gpu_means_loud_failure():
    """On a box without a CUDA device the product path must refuse to run (no oracle / CPU fallback)."""
    try:
        cuda = C.CDLL("libcuda.so.1")
        n = C.c_int(0)
        has_gpu = cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(C.byref(n)) == 0 and n.value > 0
    except OSError:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    from vicalib_b200 import capi

    with pytest.raises(capi.VcgpuError):
        capi.Calibrator()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under vicalib_b200/ may import, link or execute it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle\.binding|liboracle|#include\s+\"[^\"]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "vicalib_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                hits = [m.group(0) for m in pat.finditer(src)]
                assert not hits, (f, hits)
