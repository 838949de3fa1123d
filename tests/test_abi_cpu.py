"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/gdk.h declares,
struct layouts match, and the host logic fails loudly without a GPU (no silent fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from getdist_b200 import build

    build.build()
    from getdist_b200 import _abi

    return _abi.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "gdk.h")).read()
    names = set(re.findall(r"\b(gdk_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layout_matches_header(tmp_path):
    from getdist_b200 import _abi

    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gdk.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(gdk_spec1d), '
                   'sizeof(gdk_result1d), sizeof(gdk_spec2d), sizeof(gdk_result2d)); return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_abi.Spec1D), C.sizeof(_abi.Result1D), C.sizeof(_abi.Spec2D), C.sizeof(_abi.Result2D)]


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from getdist_b200 import MCSamples, _abi

    with pytest.raises(_abi.GdkError):
        MCSamples(samples=np.random.default_rng(0).normal(size=(100, 2)), sampler="uncorrelated")


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under getdist_b200/ may import, include or load it"""
    pkg = os.path.join(ROOT, "getdist_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|#include.*oracle|hostsim|CDLL\([^)]*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                m = bad.search(txt)
                assert m is None or "tests/hostsim" in txt[max(0, m.start() - 200): m.end() + 50] and f.endswith(".cuh"), (f, m.group(0))
