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
    """Kernel by kernel, the SASS of the built library is the one recorded in profiles/r2_device_sass_per_kernel.txt
    (tools/sass_hash.py), where every kernel carries its state: "gpu-verified" = byte-identical to the build the round's last
    B200 run executed (whole-library hash profiles/r2_device_sass.sha256), or "changed-after-last-gpu-run".  Exactly six kernels
    are in the second state: the witness conversion k_fr_to_mont and the five toMontgomery instantiations of k_fp_batch, which
    differ by the operand order of ONE Montgomery product (fp.cuh::to_mont, a raw value must be the second operand:
    tests/test_fp_asm_model.py) and whose compiled PTX -- all six -- is executed on the CPU by tests/test_ptx_kernels.py.  A change that touches
    any other kernel must be verified on a GPU and re-recorded (python tools/sass_hash.py --write)."""
    import shutil
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_hash
    if not (os.path.exists(os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")) or shutil.which("cuobjdump")):
        pytest.skip("cuobjdump not installed")
    rec = sass_hash.read_record()
    now = sass_hash.per_kernel()
    assert len(now) > 100 and set(now) == set(rec), sorted(set(now) ^ set(rec))
    differ = sorted(k for k in now if now[k] != rec[k][0])
    assert not differ, "device code differs from the record: %s" % differ
    changed = sorted(k for k, (_, status) in rec.items() if status != sass_hash.VERIFIED)
    assert len(changed) == 6 and all("k_fr_to_mont" in k or ("k_fp_batch" in k and "ELi4E" in k) for k in changed), changed


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
