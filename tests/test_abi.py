"""The C-ABI library loads and exports every symbol include/wasmsnark_b200.h declares
(no compute calls: this runs on the CPU-only box)."""
import os
import sys
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


def test_missing_library_fails_loudly(monkeypatch):
    """The product path never falls back: a missing .so is an error, not a switch to some CPU path."""
    from wasmsnark_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.LibraryNotBuilt):
        _lib.load_library("/nonexistent/libwasmsnark_b200.so")
    monkeypatch.setattr(_lib, "_lib", None)
    assert _lib.load_library() is not None


def test_product_package_does_not_import_oracle():
    """Only tests/, smoke() and the bench's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "wasmsnark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("the oracle/", ""), f


def test_product_never_reaches_for_the_cpu_kernel_emulator():
    """tests/host/cuemu (the kernels executed on the CPU, for tests) is scaffolding: neither the package, the bench, the
    driver hooks nor the Node binding may name it or the library built from it -- only tests/ does (and the report generators
    under tools/, which read the census it produced)."""
    banned = ("cuemu", "libwasmsnark_b200_emu", "WS_TEST_EMULATED")
    files = [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py")]
    for top in ("wasmsnark_b200", "node", "include", "oracle"):
        for dirpath, dirs, names in os.walk(os.path.join(ROOT, top)):
            dirs[:] = [d for d in dirs if d not in ("_obj", "_ref", "__pycache__")]
            files += [os.path.join(dirpath, f) for f in names if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".js", ".gyp", ".json"))]
    for path in files:
        src = open(path, errors="ignore").read()
        assert not any(b in src for b in banned), path


def test_device_code_is_the_gpu_verified_one():
    """The SASS of every kernel in the built library hashes to the value recorded after the last GPU run of the round
    (profiles/r2_device_sass.sha256, tools/sass_hash.py): later commits changed host code, tests and documents only.  A change
    that touches device code must be re-verified on a GPU and the hash re-recorded (python tools/sass_hash.py --write)."""
    import shutil
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_hash
    if not (os.path.exists(os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")) or shutil.which("cuobjdump")):
        pytest.skip("cuobjdump not installed")
    want = open(os.path.join(ROOT, "profiles", "r2_device_sass.sha256")).read().split()[0]
    got, kernels = sass_hash.sass_hash()
    assert kernels > 100 and got == want, "device code differs from the GPU-verified build"


def test_node_addon_type_checks_against_the_header():
    """node/wasmsnark_b200_addon.c (the N-API binding of INTEGRATION.md) cannot be built here -- no Node.js -- but
    it must at least compile against include/wasmsnark_b200.h and a declaration-only N-API stub."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    assert cc, "no C compiler"
    r = subprocess.run([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "host", "napi_mock"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "node", "wasmsnark_b200_addon.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
