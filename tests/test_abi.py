"""The C-ABI library loads and exports every symbol include/wasmsnark_b200.h declares
(no compute calls: this runs on the CPU-only box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "wasmsnark_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_header_symbols():
    from wasmsnark_b200 import build as wsbuild
    from wasmsnark_b200 import _lib
    wsbuild.build()
    lib = _lib.load_library()
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
        assert s in _lib.SIGNATURES, "python binding lacks %s" % s
    assert lib.ws_version().startswith(b"wasmsnark_b200")


def test_no_device_fails_loudly():
    """Without a GPU the product path must raise, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wasmsnark_b200 import _lib
    with pytest.raises(_lib.WsError):
        _lib.Context(0)
