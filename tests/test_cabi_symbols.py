"""CPU: the C-ABI shared library loads and exports every entry point declared in include/os2d_b200.h
(no compute call is made - there is no GPU here)."""
import ctypes
import os
import re

from os2d_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "os2d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(os2d_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for required in ("os2d_pack_class_features", "os2d_pack_image_features", "os2d_correlate", "os2d_transform_conv",
                     "os2d_resample_boxes", "os2d_decode_boxes", "os2d_nms_segments", "os2d_b200_last_error"):
        assert required in syms


def test_library_exports_every_declared_symbol():
    path = _cabi.library_path()
    assert os.path.exists(path), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(path)
    for name in _declared_symbols():
        assert hasattr(lib, name), "missing export " + name


def test_ctypes_binding_covers_the_header():
    assert sorted(_cabi.SIGNATURES.keys()) == _declared_symbols()
    lib = _cabi.load()
    assert lib.os2d_b200_abi_version() == 1
    assert lib.os2d_conv_weight_blob_bytes(7, 15) == 15 * 49 * 4096
