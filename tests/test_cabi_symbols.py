"""The C-ABI library loads on a CPU-only box and exports every symbol include/edgefem_b200.h declares
(no compute calls here)."""
import ctypes
import os
import re

import pytest

from edgefem_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "edgefem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(efb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(cabi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_ctypes_table_matches_header():
    assert sorted(cabi.SIGNATURES.keys()) == header_functions()


def test_abi_version_and_no_gpu_error():
    lib = cabi.load()
    assert lib.efb_abi_version() == 1
    if lib.efb_device_count() == 0:
        with pytest.raises(cabi.EfbError, match="no CUDA device"):
            cabi.Ctx(0)
        assert b"no CPU fallback" in lib.efb_last_error(None)


def test_struct_layouts_match_c():
    # sizes the C side relies on (include/edgefem_b200.h)
    assert ctypes.sizeof(cabi.Model) == 40 and ctypes.sizeof(cabi.Pole) == 24 and ctypes.sizeof(cabi.Pml) == 64
    assert ctypes.sizeof(cabi.SolveOpts) == 40 and ctypes.sizeof(cabi.SolveResult) == 24
