"""tests/host/cuemu/run_emulated.py -- TEST SCAFFOLDING: runs a Python script (bench.py, tools/run_*_once.py, ...) with the library's
kernels executed on the CPU emulator and torch's CUDA buffers / streams / events replaced by host stand-ins:

    python tests/host/cuemu/run_emulated.py bench.py --steps 2 --warmup 3 --only-prove

Timings printed by such a run are meaningless (CPU emulation); the point is that the script's own logic -- argument handling,
section order, golden checks, JSON assembly -- can be exercised on a box without a GPU (tests/test_emulated_library.py)."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import build_emu
    import emulated_torch
    from wasmsnark_b200 import _lib
    _lib.load_library(path=build_emu.build())
    emulated_torch.patch()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
