"""CPU-side checks of the drop-in boundary: the shared library loads and exports every
symbol include/barcoder_b200.h declares.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from barcoder_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "barcoder_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bc_[a-z_]+)\s*\(", text)))


def test_header_and_loader_agree():
    assert _declared() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert _native.load().bc_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.NativeError) as ei:
        _native.Searcher(0)
    assert ei.value.code in (-2, -3)


def test_stats_struct_size_matches_header():
    # 2*u64 + 6*u32 + 3*u64 + 5*f32 + 8*u32 with natural alignment
    assert ctypes.sizeof(_native.BcStats) == 16 + 24 + 24 + 20 + 32 + 4 + 8


def test_constants_match_header():
    """Every BC_PARAM_* / BC_E* / BC_PAM_* value the Python mirror carries equals the header's."""
    text = open(os.path.join(ROOT, "include", "barcoder_b200.h")).read()
    defs = {}
    for name, val in re.findall(r"#define\s+(BC_[A-Z_0-9]+)\s+\(?(-?\d+)u?\)?", text):
        defs[name] = int(val)
    assert defs["BC_OK"] == 0 and defs["BC_ELIMIT"] == -5
    checked = 0
    for name, val in defs.items():
        if hasattr(_native, name):
            assert getattr(_native, name) == val, name
            checked += 1
    for must in ("BC_PARAM_BLOCKS", "BC_PARAM_PATH", "BC_PARAM_HIT_CAPACITY", "BC_PARAM_SPACER_ID_BASE",
                 "BC_PARAM_SCAN_PART", "BC_PARAM_WINDOW_SORT", "BC_ELIMIT"):
        assert hasattr(_native, must) and must in defs, must
    assert checked >= 10
