#!/usr/bin/env python3
"""build_ref.py -- TEST INFRASTRUCTURE.  Recipe for oracle/_ref/: translates the reference's
pre-generated WASM images (read in place from /root/reference/build/) to C with wasm2c.py and
compiles them with gcc.  Outputs only under oracle/_ref/ (git-ignored, shipped to the GPU box).
No-op when the reference is not mounted and the libraries already exist."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WASMSNARK_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ("bn128", "mnt6753"):
        js = os.path.join(REF, "build", name + "_wasm.js")
        c = os.path.join(OUT, name + "_wasm.c")
        so = os.path.join(OUT, "lib%s_ref.so" % name)
        if not os.path.exists(js):
            print("reference not mounted; keeping %s" % ("existing " + so if os.path.exists(so) else "NOTHING (oracle/_ref unavailable)"))
            continue
        if os.path.exists(so) and os.path.getmtime(so) > max(os.path.getmtime(js), os.path.getmtime(os.path.join(HERE, "wasm2c.py"))):
            continue
        subprocess.check_call([sys.executable, os.path.join(HERE, "wasm2c.py"), js, c])
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-o", so, c])


def copy_example():
    """The reference's example key / witness / golden proof (35 MB of DATA, not sources) into the
    git-ignored oracle/_ref/example_bn128/, so that the C4 parity test can run on the GPU box,
    where /root/reference does not exist."""
    import shutil
    src = os.path.join(REF, "example", "bn128")
    dst = os.path.join(OUT, "example_bn128")
    if not os.path.isdir(src):
        return
    os.makedirs(dst, exist_ok=True)
    for f in ("proving_key.bin", "witness.bin", "proof_good0.json", "proof_good.json", "verification_key.json", "public.json"):
        if not os.path.exists(os.path.join(dst, f)):
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))


if __name__ == "__main__":
    copy_example()
    main()
