"""tests/host/cuemu/build_emu.py -- TEST SCAFFOLDING: builds tests/host/libwasmsnark_b200_emu.so, the library's own .cu sources
compiled by g++ against the CPU kernel emulator in this directory (cuda_runtime.h, cuemu.cpp).

The sources are copied into tests/host/cuemu/_build/pkg/csrc/ with two mechanical rewrites (nothing else is touched):
  kernel<<<grid, block, smem, stream>>>(args);   ->  { auto a = std::make_tuple(args); cuemu::launch_on(stream, dim3(grid), dim3(block), smem, "kernel", [a]() { std::apply(kernel, a); }); }
  extern __shared__ T name[];                    ->  T* name = (T*)cuemu::dyn_smem();
  __noinline__                                   ->  __attribute__((noinline))
plus one counting line in the host path of fp.cuh's product and squaring (operation census, cuemu.cpp).
Only tests load the result (tests/test_emulated_library.py); see cuda_runtime.h for what it does and does not cover.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, "wasmsnark_b200", "csrc")
BUILD = os.path.join(HERE, "_build", "pkg", "csrc")      # so that "../../include/wasmsnark_b200.h" stays inside _build/
OUT = os.path.join(ROOT, "tests", "host", "libwasmsnark_b200_emu.so")
OUT_ASAN = os.path.join(ROOT, "tests", "host", "libwasmsnark_b200_emu_asan.so")


def _matching(s, i, open_c, close_c, step):
    """index of the bracket matching s[i] (scanning forwards, step = 1, or backwards, step = -1)"""
    depth = 0
    while 0 <= i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += step
    raise ValueError("unbalanced brackets")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            break
        # kernel expression: identifier, optionally followed by a balanced <...> template argument list
        k = i - 1
        while src[k].isspace():
            k -= 1
        if src[k] == ">":
            k = _matching(src, k, ">", "<", -1) - 1
        while src[k].isalnum() or src[k] in "_:":
            k -= 1
        kern = src[k + 1:i].strip()
        j = src.index(">>>", i)
        cfg = _split_top(src[i + 3:j])
        assert 2 <= len(cfg) <= 4, cfg
        a0 = j + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == "(", src[i:a0 + 1]
        a1 = _matching(src, a0, "(", ")", 1)
        end = a1 + 1
        while src[end].isspace():
            end += 1
        assert src[end] == ";", src[i:end + 1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:k + 1])
        name = re.sub(r"\s+", " ", kern).replace('"', "")
        # the name the census files the launch under: the kernel expression + the template bindings of the launch site
        # the arguments are evaluated HERE, at the launch, and copied into the closure -- as the hardware launch does; the
        # closure may run much later (CUEMU_ASYNC: deferred execution of the streams, cuemu.cpp)
        stream = cfg[3] if len(cfg) > 3 else "nullptr"
        out.append('{ auto cuemu_args = std::make_tuple%s; cuemu::launch_on((cudaStream_t)(%s), dim3(%s), dim3(%s), %s, '
                   'cuemu::kernel_label("%s", __PRETTY_FUNCTION__), [=]() { std::apply([=](auto&... cuemu_a) { %s(cuemu_a...); }, cuemu_args); }); }'
                   % (src[a0:a1 + 1], stream, cfg[0], cfg[1], smem, name, kern))
        pos = end + 1
    out.append(src[pos:])
    return "".join(out)


# operation census: one counting line in the portable host path of the COPIED field layer (product and squaring)
CENSUS = (("    // portable CIOS on 64-bit limbs (host:", "    cuemu::count_op(N, 0);\n"),
          ("#else\n    return mul(a, a);\n#endif", None))


def add_census(name, src):
    if name != "fp.cuh":
        return src
    anchor, line = CENSUS[0]
    assert src.count(anchor) == 1, "fp.cuh: host path of mul_inline not found"
    src = src.replace(anchor, line + anchor)
    anchor = CENSUS[1][0]
    assert src.count(anchor) == 1, "fp.cuh: host path of sqr not found"
    return src.replace(anchor, "#else\n    cuemu::count_op(P::N, 1);\n    return mul(a, a);\n#endif")


def rewrite(src):
    src = re.sub(r"\b__noinline__\b", "__attribute__((noinline))", src)
    src = re.sub(r"extern\s+__shared__\s+([\w:]+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)cuemu::dyn_smem();", src)
    return rewrite_launches(src)


def build(verbose=False, asan=False, ubsan=False):
    """asan=True: the same build with AddressSanitizer (tests/host/libwasmsnark_b200_emu_asan.so; run python with
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0): out-of-bounds accesses of kernels and host
    code are reported -- the GPU-less counterpart of compute-sanitizer's memcheck -- and so are misaligned accesses (a 16-byte
    vector load from an address that is not 16-byte aligned faults on the GPU)."""
    global BUILD, OUT
    if asan:
        BUILD = os.path.join(HERE, "_build", "asan", "csrc")
        OUT = OUT_ASAN
    elif ubsan:        # UndefinedBehaviorSanitizer alone (no LD_PRELOAD needed): shifts by >= the width, misaligned accesses, ...
        BUILD = os.path.join(HERE, "_build", "ubsan", "csrc")
        OUT = OUT_ASAN.replace("_asan", "_ubsan")
    os.makedirs(BUILD, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    deps = [os.path.join(CSRC, f) for f in srcs] + [os.path.join(HERE, f) for f in ("cuda_runtime.h", "cuemu.cpp", "build_emu.py")] + \
           [os.path.join(ROOT, "include", "wasmsnark_b200.h")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    for f in srcs:
        text = add_census(f, rewrite(open(os.path.join(CSRC, f)).read()))
        assert "<<<" not in text, f
        # the sources include the public header as "../../include/wasmsnark_b200.h": same relative position under _build/
        with open(os.path.join(BUILD, f), "w") as o:
            o.write(text)
    inc = os.path.join(BUILD, "..", "..", "include")
    os.makedirs(inc, exist_ok=True)
    with open(os.path.join(inc, "wasmsnark_b200.h"), "w") as o:
        o.write(open(os.path.join(ROOT, "include", "wasmsnark_b200.h")).read())
    flags = ["-std=c++17", "-O2", "-fno-strict-aliasing", "-fPIC", "-fvisibility=hidden", "-w", "-D__CUDACC__=1", "-I" + HERE, "-include", os.path.join(HERE, "cuda_runtime.h")]
    if asan:
        flags = [f for f in flags if f != "-O2"] + ["-O1", "-g", "-fsanitize=address,alignment", "-fno-sanitize-recover=alignment", "-fno-omit-frame-pointer", "-DCUEMU_ASAN=1"]
    if ubsan and not asan:
        flags = [f for f in flags if f != "-O2"] + ["-O1", "-g", "-fsanitize=undefined", "-fno-sanitize=vptr", "-fno-sanitize-recover=undefined"]
    units = [(os.path.join(BUILD, f), os.path.join(BUILD, f[:-3] + ".o")) for f in srcs if f.endswith(".cu")]
    units.append((os.path.join(HERE, "cuemu.cpp"), os.path.join(BUILD, "cuemu.o")))

    def cc(unit):
        src, obj = unit
        return unit, subprocess.run(["g++"] + flags + ["-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)
    failed = False
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
        for (src, obj), r in ex.map(cc, units):
            if r.returncode or verbose:
                sys.stderr.write("== %s\n%s%s" % (src, r.stdout, r.stderr[:6000]))
            failed |= r.returncode != 0
    if failed:
        raise RuntimeError("emulated build failed")
    r = subprocess.run(["g++", "-shared", "-o", OUT] + (["-fsanitize=address,alignment"] if asan else ["-fsanitize=undefined"] if ubsan else []) + [o for _, o in units] + ["-lpthread"], capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr)
        raise RuntimeError("emulated link failed")
    return OUT


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv, asan="--asan" in sys.argv, ubsan="--ubsan" in sys.argv))
