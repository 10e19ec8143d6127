"""CPU: the C-ABI library builds, loads, exports every symbol include/gdbgpu.h declares, and refuses to compute
without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import REPO


def declared_functions():
    src = open(os.path.join(REPO, "include", "gdbgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gdbgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_all_declared_symbols():
    import __graft_entry__ as g
    lib_path = g.build_gpu_lib()
    lib = ctypes.CDLL(lib_path)
    names = declared_functions()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    from genomicsdb_b200 import capi
    assert sorted(capi.EXPORTS) == names


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genomicsdb_b200 import capi
    with pytest.raises(capi.GdbGpuError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_CUDA


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under genomicsdb_b200/ may import, include or load it"""
    pkg = os.path.join(REPO, "genomicsdb_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "liboracle" not in txt and "combine_oracle" not in txt and '#include "../../oracle' not in txt, f
