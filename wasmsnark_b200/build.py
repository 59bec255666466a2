"""Builds libwasmsnark_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m wasmsnark_b200.build [--verbose] [--force]

Each .cu under csrc/ is its own translation unit, compiled in parallel, then linked.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
OUT = os.path.join(HERE, "libwasmsnark_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "wasmsnark_b200.h"))
    return hs


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hdr_t, os.path.getmtime(src)):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return job, r

    failed = False
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        for (src, obj), r in ex.map(compile_one, jobs):
            if verbose or r.returncode:
                sys.stderr.write("== %s\n%s%s" % (src, r.stdout, r.stderr))
            failed |= r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources]
    if jobs or not os.path.exists(OUT):
        r = subprocess.run([nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    build(verbose="--verbose" in sys.argv, force="--force" in sys.argv)
    print(OUT)
