"""tools/sass_hash.py -- sha256 of the device code (SASS of every kernel) of wasmsnark_b200/libwasmsnark_b200.so.

    python tools/sass_hash.py            # prints the hash
    python tools/sass_hash.py --write    # records it in profiles/r2_device_sass.sha256

The last GPU run of round 2 (gpurun session 10: 86 GPU tests, smoke, bench, launch list) exercised exactly this device code; every
later commit of the round changed host code, tests and documents only.  tests/test_abi.py compares the built library with the
recorded hash, so a change that touches device code cannot slip in without being noticed (and re-verified on a GPU)."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wasmsnark_b200", "libwasmsnark_b200.so")
REC = os.path.join(ROOT, "profiles", "r2_device_sass.sha256")


def sass_hash(lib=LIB):
    cuobjdump = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
    out = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    keep, n = [], 0
    for line in out.splitlines():
        if line.startswith(("Fatbin", "=====")) or re.match(r"\s*(host|compile_size|identifier|arch|code version|producer)\b", line):
            continue                      # container headers (paths, tool versions); the code and the kernel names stay
        n += line.lstrip().startswith("Function :")
        keep.append(line)
    return hashlib.sha256("\n".join(keep).encode()).hexdigest(), n


if __name__ == "__main__":
    h, n = sass_hash()
    if "--write" in sys.argv:
        with open(REC, "w") as f:
            f.write("%s  %d kernels  (cuobjdump -sass of libwasmsnark_b200.so, container headers removed; tools/sass_hash.py)\n" % (h, n))
    print(h, n)
