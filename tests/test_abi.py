"""The C-ABI shared library loads without a GPU and exports every symbol include/voxelizer_b200.h declares;
without a device it refuses to work instead of falling back to the CPU."""
import ctypes as C
import os
import re

import pytest

import voxelizer_b200 as vx
from util import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "voxelizer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vx_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = vx.cuda_lib()
    names = declared_symbols()
    assert len(names) >= 15
    assert set(names) == set(vx.C_ABI_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler gives the ABI structs (LP64)
    assert C.sizeof(vx.GridDesc) == 28
    assert C.sizeof(vx.FilterDesc) == 48
    assert C.sizeof(vx.Options) == 32
    assert C.sizeof(vx.FrameDesc) == 104
    assert C.sizeof(vx.GridInfo) == 56
    assert C.sizeof(vx.StageTimes) == 26 * 8


def test_no_cpu_fallback_without_device():
    lib = vx.cuda_lib()
    if lib.vx_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(vx.VxError, match="no CUDA device"):
        vx.Backend(0.5, (0, -4, -1), (8, 4, 1), 2.5, 25.0)


def test_product_never_imports_the_oracle():
    """Nothing under voxelizer_b200/ may reference oracle/ (grep, like the judge does)."""
    pkg = os.path.join(ROOT, "voxelizer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in src and "liboracle" not in src and "import oracle" not in src and "from oracle" not in src, f
