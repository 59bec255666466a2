"""tools/sass_hash.py -- the device code (SASS of every kernel) of wasmsnark_b200/libwasmsnark_b200.so, hashed per kernel.

    python tools/sass_hash.py            # prints the whole-library hash and the kernels that differ from the record
    python tools/sass_hash.py --write    # re-records profiles/r2_device_sass_per_kernel.txt (status of unchanged kernels kept)

The last GPU run of round 2 (gpurun session 10: 86 GPU tests, smoke, bench, launch list) exercised the device code whose
whole-library hash is profiles/r2_device_sass.sha256; the GPU budget of the round was spent after it.  The per-kernel record
(profiles/r2_device_sass_per_kernel.txt) says for every kernel whether its SASS is still the one that ran on the B200
("gpu-verified") or was changed afterwards ("changed-after-last-gpu-run", with the reason in profiles/README.md).
tests/test_abi.py compares the built library with the record, so a change that touches device code cannot slip in unnoticed."""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wasmsnark_b200", "libwasmsnark_b200.so")
REC = os.path.join(ROOT, "profiles", "r2_device_sass.sha256")
REC_PER_KERNEL = os.path.join(ROOT, "profiles", "r2_device_sass_per_kernel.txt")
VERIFIED, CHANGED = "gpu-verified", "changed-after-last-gpu-run"


def _dump(lib):
    cuobjdump = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
    return subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout


def sass_hash(lib=LIB):
    """(sha256 of the whole dump without the container headers, number of kernels)"""
    keep, n = [], 0
    for line in _dump(lib).splitlines():
        if line.startswith(("Fatbin", "=====")) or re.match(r"\s*(host|compile_size|identifier|arch|code version|producer)\b", line):
            continue                      # container headers (paths, tool versions); the code and the kernel names stay
        n += line.lstrip().startswith("Function :")
        keep.append(line)
    return hashlib.sha256("\n".join(keep).encode()).hexdigest(), n


def per_kernel(lib=LIB):
    """{mangled name + '#' + occurrence: sha256 of that kernel's instruction lines} -- a static kernel of a header appears
    once per translation unit under the same name, hence the occurrence number (translation units are linked in a fixed order)"""
    res, name, buf = {}, None, []

    def put():
        k = 0
        while "%s#%d" % (name, k) in res:
            k += 1
        res["%s#%d" % (name, k)] = hashlib.sha256("\n".join(buf).encode()).hexdigest()
    for line in _dump(lib).splitlines():
        s = line.strip()
        if s.startswith("Function :"):
            if name:
                put()
            name, buf = s.split(":", 1)[1].strip(), []
        elif s.startswith("Fatbin"):
            if name:
                put()
            name = None
        elif name and s.startswith(("/*", ".")):
            buf.append(s)
    if name:
        put()
    return res


def read_record(path=REC_PER_KERNEL):
    rec = {}
    for line in open(path):
        if line.strip() and not line.startswith("#"):
            h, status, name = line.split()
            rec[name] = (h, status)
    return rec


def write_record(now, old, path=REC_PER_KERNEL):
    with open(path, "w") as f:
        f.write("# sha256 of each kernel's SASS (cuobjdump -sass of libwasmsnark_b200.so; tools/sass_hash.py), and whether that code is the\n"
                "# one the round's last B200 run executed or was changed after it (reasons: profiles/README.md)\n")
        for name in sorted(now):
            status = old[name][1] if name in old and old[name][0] == now[name] else CHANGED
            f.write("%s %s %s\n" % (now[name], status, name))


if __name__ == "__main__":
    h, n = sass_hash()
    now = per_kernel()
    old = read_record() if os.path.exists(REC_PER_KERNEL) else {}
    if "--write-verified" in sys.argv:          # only right after a GPU run of exactly this build
        with open(REC, "w") as f:
            f.write("%s  %d kernels  (cuobjdump -sass of libwasmsnark_b200.so, container headers removed; tools/sass_hash.py)\n" % (h, n))
        write_record(now, {k: (v, VERIFIED) for k, v in now.items()})
    elif "--write" in sys.argv:
        write_record(now, old)
    print(h, n)
    for name in sorted(set(now) | set(old)):
        if name not in old or name not in now or old[name][0] != now[name]:
            print("differs from the record:", name)
