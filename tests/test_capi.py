"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports every function
include/hippopt_b200.h declares.  No compute calls here.  CPU only."""
import ctypes
import re

from hippopt_b200 import _capi


def declared_functions():
    text = open(_capi.HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_library):
    names = declared_functions()
    assert set(names) == set(_capi.EXPORTED_SYMBOLS)
    raw = ctypes.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in the header but not exported"


def test_header_tables_are_consistent():
    H = _capi.H
    assert H["HB_KF_COUNT"] == 12 * 8 + 21
    assert H["HB_KI_COUNT"] == H["HB_KI_FAM0"] + 4 * H["HB_KF_COUNT"]
    assert H["HB_KD_COUNT"] == H["HB_KD_BODY0"] + H["HB_KD_BODY_STRIDE"] * H["HB_MAX_BODIES"]
    assert (H["HB_EVAL_F"], H["HB_EVAL_GRAD_F"], H["HB_EVAL_G"], H["HB_EVAL_JAC_G"], H["HB_EVAL_HESS_L"]) == (1, 2, 4, 8, 16)


def test_error_path_without_gpu(built_library):
    """hb_dims on a NULL handle must return an error code and a message, not crash."""
    rc = built_library.hb_dims(None, None, None, None, None, None)
    assert rc != 0 and b"null" in built_library.hb_last_error()
