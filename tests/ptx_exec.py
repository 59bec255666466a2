"""tests/ptx_exec.py -- TEST SCAFFOLDING: executes the PTX that nvcc emits for one of the library's element-wise kernels on
the CPU, one CUDA thread at a time, with Python integers.

Why: the kernel emulator (tests/host/cuemu) compiles the .cu sources with g++, so the field layer takes its portable host
path there and the PTX carry chains of fp_asm.cuh -- mad.lo.cc / madc.hi.cc sequences written as inline asm -- are never
executed without a GPU.  tests/test_fp_asm_model.py models the generator's instruction lists as a human composed them; this
file runs what the COMPILER composed: `nvcc -ptx` of the real translation unit, the kernel's entry as it stands (inline asm,
constant folding, the optimiser's register renaming and all).  What is left to the hardware run is ptxas (PTX -> SASS).

Supported: the integer subset these kernels use (mov, add/sub/mul/mad with .cc / addc / subc / madc carry forms, mul.wide,
shifts, logic, setp / selp, predicated execution and branches, ld.param, ld/st scalar and .v2/.v4 vectors on global, local
and generic addresses, cvta, per-frame local memory, calls of .func device functions with .param passing).
Anything else raises -- a kernel outside the subset fails loudly instead of being half-executed."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wasmsnark_b200", "csrc")
CACHE = os.path.join(ROOT, "tests", "host", "_ptx")

M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF


def ptx_of(cu_name):
    """`nvcc -ptx` of wasmsnark_b200/csrc/<cu_name> with the library's own flags (cached under tests/host/_ptx/)."""
    src = os.path.join(CSRC, cu_name)
    out = os.path.join(CACHE, cu_name[:-3] + ".ptx")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(ROOT, "include", "wasmsnark_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(CACHE, exist_ok=True)
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=compute_100a", "-std=c++17", "-O3", "-ptx", src, "-o", out + ".tmp"])
        os.replace(out + ".tmp", out)
    return open(out).read()


def ptx_of_many(cu_names):
    """the same for several translation units at once (nvcc runs side by side: a cold cache costs the slowest, not the sum)"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(cu_names)) as ex:
        return dict(zip(cu_names, ex.map(ptx_of, cu_names)))


class Code(list):
    """statements of one entry / function + what they need around them"""
    funcs = None          # {name: (parameter names, Code)} of the whole translation unit
    local_bytes = 0       # size of the frame's local depot
    depot = None          # its symbol


def _parse_body(body):
    body = re.sub(r"//[^\n]*", "", body)
    code = Code()
    m = re.search(r"\.local\s+\.align\s+\d+\s+\.b8\s+(\w+)\[(\d+)\]", body)
    if m:
        code.depot, code.local_bytes = m.group(1), int(m.group(2))
    for raw in body.replace("\n", " ").split(";"):
        s = raw.strip()
        while s[:1] in ("{", "}"):                         # call sequences sit in their own { } scope
            s = s[1:].strip()
        while True:                                        # labels in front of a statement
            m = re.match(r"(\$?[\w$]+):\s*", s)
            if not m or s.startswith("@"):
                break
            code.append(("label", m.group(1)))
            s = s[m.end():]
        if s and not s.startswith((".reg", ".local", ".shared", ".maxntid", ".reqntid", ".pragma", ".param")):
            code.append(("ins", s))
    return code


def _split_unit(ptx, start):
    """(parameter names, Code) of the .entry / .func whose parameter list opens at ptx[start - 1] == '('"""
    j = ptx.index(")", start)
    params = re.findall(r"\.param\s+\.\w+(?:\s+\.ptr)?(?:\s+\.align\s+\d+)?\s+(\w+)", ptx[start:j])
    b0 = ptx.index("{", j)
    b1 = ptx.index("\n}", b0)
    return params, _parse_body(ptx[b0 + 1:b1])


def functions(ptx):
    """{name: (parameter names, Code)} of every .func DEFINED in the translation unit (no return values: the library's
    out-of-line helpers return through pointers)"""
    funcs = {}
    for m in re.finditer(r"\.func\s+(\w+)\s*\(", ptx):
        end = ptx.index(")", m.end())
        k = end + 1
        while ptx[k].isspace():
            k += 1
        if ptx[k] == "{":                                   # a definition, not a forward declaration
            funcs[m.group(1)] = _split_unit(ptx, m.end())
    for _, code in funcs.values():
        code.funcs = funcs
    return funcs


def entry(ptx, name_part):
    """(parameter names, statements) of the one .entry whose mangled name contains every string of name_part"""
    parts = [name_part] if isinstance(name_part, str) else list(name_part)
    hits = [m for m in re.finditer(r"\.entry\s+(\w+)\s*\(", ptx) if all(p in m.group(1) for p in parts)]
    assert len(hits) == 1, "%d entries match %r" % (len(hits), parts)
    params, code = _split_unit(ptx, hits[0].end())
    code.funcs = functions(ptx) if "call" in ptx[hits[0].end():ptx.index("\n}", hits[0].end())] else {}
    return params, code


class Memory:
    """global memory: named buffers at fixed fake addresses"""

    def __init__(self):
        self.bufs, self.next = [], 0x10000000

    def alloc(self, data):
        base = self.next
        buf = bytearray(data)
        self.bufs.append((base, buf))
        self.next += (len(buf) + 0xFFFF) & ~0xFFFF
        return base

    def free(self, base):
        self.bufs = [(b, buf) for b, buf in self.bufs if b != base]

    def _find(self, addr, size):
        for base, buf in self.bufs:
            if base <= addr and addr + size <= base + len(buf):
                return buf, addr - base
        raise IndexError("global access outside every buffer: 0x%x + %d" % (addr, size))

    def load(self, addr, size):
        assert addr % size == 0, "misaligned %d-byte load at 0x%x" % (size, addr)
        buf, off = self._find(addr, size)
        return int.from_bytes(buf[off:off + size], "little")

    def store(self, addr, size, v):
        assert addr % size == 0, "misaligned %d-byte store at 0x%x" % (size, addr)
        buf, off = self._find(addr, size)
        buf[off:off + size] = (v & ((1 << (8 * size)) - 1)).to_bytes(size, "little")

    def bytes_at(self, base):
        return bytes(next(buf for b, buf in self.bufs if b == base))


_TYPE_BITS = {"u8": 8, "s8": 8, "b8": 8, "u16": 16, "s16": 16, "b16": 16, "u32": 32, "s32": 32, "b32": 32, "u64": 64, "s64": 64, "b64": 64}


def _signed(v, bits):
    return v - (1 << bits) if v >> (bits - 1) else v


def run_thread(params, stmts, args, mem, tid, ntid, ctaid, nctaid=(1, 1, 1), max_steps=5_000_000):
    """one CUDA thread from entry to ret"""
    budget = [max_steps]
    special = {"%tid": tid, "%ntid": ntid, "%ctaid": ctaid, "%nctaid": nctaid}
    _run_frame(params, stmts, args, mem, special, budget)


def _run_frame(params, stmts, args, mem, special, budget):
    """one activation of an entry or a device function: its own registers, carry flag and local depot"""
    reg, cc = {}, 0
    pvals = dict(zip(params, args))
    labels = {s[1]: i for i, s in enumerate(stmts) if s[0] == "label"}
    depot = getattr(stmts, "depot", None)
    local_base = mem.alloc(bytes(stmts.local_bytes)) if depot else None
    outgoing = {}                                  # .param values of the call being set up
    try:
        return _run_body(stmts, reg, cc, pvals, labels, depot, local_base, outgoing, mem, special, budget)
    finally:
        if local_base is not None:
            mem.free(local_base)


def _run_body(stmts, reg, cc, pvals, labels, depot, local_base, outgoing, mem, special, budget):
    def val(tok, bits=32):
        tok = tok.strip()
        if tok.startswith("%"):
            m = re.match(r"(%n?c?taid|%n?tid)\.([xyz])$", tok)
            if m:
                return special[m.group(1)]["xyz".index(m.group(2))]
            return reg[tok]
        if depot and tok == depot:
            return local_base
        return int(tok, 0) & ((1 << bits) - 1)

    def addr_of(tok):
        m = re.match(r"\[\s*([%\w$]+)\s*(?:\+\s*(-?\w+))?\s*\]$", tok.strip())
        assert m, tok
        base = m.group(1)
        off = int(m.group(2), 0) if m.group(2) else 0
        return base, off

    pc = 0
    while pc < len(stmts):
        budget[0] -= 1
        assert budget[0] > 0, "runaway kernel"
        kind, s = stmts[pc]
        pc += 1
        if kind == "label":
            continue
        m = re.match(r"@(!?)(%p\d+)\s+", s)
        if m:
            take = bool(reg[m.group(2)]) != bool(m.group(1))
            s = s[m.end():]
            if not take:
                continue
        op, _, rest = s.partition(" ")
        rest = rest.strip()
        if op in ("ret", "exit"):
            return
        if op in ("bra", "bra.uni"):
            pc = labels[rest.strip()]
            continue
        f = op.split(".")
        base = f[0]
        # operands: split on commas outside braces / brackets
        ops, depth, cur = [], 0, ""
        for ch in rest:
            if ch in "{[(":
                depth += 1
            elif ch in "}])":
                depth -= 1
            if ch == "," and depth == 0:
                ops.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            ops.append(cur.strip())
        ty = f[-1]
        bits = _TYPE_BITS.get(ty)
        if base == "ld" and f[1] == "param":
            name, off = addr_of(ops[1])
            assert off == 0
            reg[ops[0]] = pvals[name] & ((1 << bits) - 1)
        elif base == "st" and f[1] == "param":
            name, off = addr_of(ops[0])
            assert off == 0
            outgoing[name] = val(ops[1], bits)
        elif base == "call":
            assert len(ops) == 2 and ops[1].startswith("("), "ptx_exec: calls with return values are not modelled: " + s
            fparams, fcode = stmts.funcs[ops[0]]
            names = [x.strip() for x in ops[1].strip("() ").split(",")]
            _run_frame(fparams, fcode, [outgoing[x] for x in names], mem, special, budget)
            outgoing.clear()
        elif base == "cvta":
            reg[ops[0]] = val(ops[1], 64)
        elif base in ("ld", "st"):
            assert not ({"shared", "const", "param"} & set(f)), op        # global, local or generic: one flat address space here
            vec = 4 if "v4" in f else 2 if "v2" in f else 1
            size = bits // 8
            if base == "ld":
                b, off = addr_of(ops[1])
                a = reg[b] + off
                assert a % (size * vec) == 0, "misaligned vector load"
                dst = [x.strip() for x in ops[0].strip("{} ").split(",")]
                assert len(dst) == vec
                for k, d in enumerate(dst):
                    reg[d] = mem.load(a + k * size, size)
            else:
                b, off = addr_of(ops[0])
                a = reg[b] + off
                assert a % (size * vec) == 0, "misaligned vector store"
                src = [x.strip() for x in ops[1].strip("{} ").split(",")]
                assert len(src) == vec
                for k, x in enumerate(src):
                    mem.store(a + k * size, size, val(x, bits))
        elif base == "mov":
            reg[ops[0]] = val(ops[1], bits)
        elif base in ("add", "addc", "sub", "subc") and bits:
            a, b = val(ops[1], bits), val(ops[2], bits)
            cin = cc if base in ("addc", "subc") else 0
            if base in ("add", "addc"):
                t = a + b + cin
                cout = t >> bits
            else:
                t = a - b - cin
                cout = 1 if t < 0 else 0
            reg[ops[0]] = t & ((1 << bits) - 1)
            if "cc" in f:
                cc = cout
        elif base in ("mul", "mad", "madc"):
            mode = f[1]
            a, b = val(ops[1], bits), val(ops[2], bits)
            if mode == "wide":
                assert ty in ("u32", "s32")
                pr = (_signed(a, 32) * _signed(b, 32)) if ty == "s32" else a * b
                t = pr + (val(ops[3], 64) if base != "mul" else 0)
                reg[ops[0]] = t & M64
                continue
            if ty.startswith("s") and mode == "hi":
                pr = _signed(a, bits) * _signed(b, bits)
            else:
                pr = a * b
            part = (pr & ((1 << bits) - 1)) if mode == "lo" else ((pr >> bits) & ((1 << bits) - 1))
            if base == "mul":
                reg[ops[0]] = part
            else:
                t = part + val(ops[3], bits) + (cc if base == "madc" else 0)
                reg[ops[0]] = t & ((1 << bits) - 1)
                if "cc" in f:
                    cc = t >> bits
        elif base in ("and", "or", "xor"):
            a, b = int(val(ops[1], bits or 1)), int(val(ops[2], bits or 1))           # .pred operands are Python bools
            reg[ops[0]] = (a & b) if base == "and" else (a | b) if base == "or" else (a ^ b)
        elif base == "neg":
            reg[ops[0]] = (-val(ops[1], bits)) & ((1 << bits) - 1)
        elif base == "not":
            reg[ops[0]] = ~int(val(ops[1], bits or 1)) & ((1 << (bits or 1)) - 1)
        elif base in ("shl", "shr"):
            a, n = val(ops[1], bits), val(ops[2], 32)
            if base == "shl":
                reg[ops[0]] = (a << n) & ((1 << bits) - 1) if n < bits else 0
            elif ty.startswith("s"):
                reg[ops[0]] = (_signed(a, bits) >> min(n, bits - 1)) & ((1 << bits) - 1)
            else:
                reg[ops[0]] = a >> n if n < bits else 0
        elif base == "shf":                       # shf.l.wrap.b32 / shf.l.clamp.b32 d, lo, hi, n  (funnel shift)
            lo, hi, n = val(ops[1]), val(ops[2]), val(ops[3])
            n = n & 31 if "wrap" in f else min(n, 32)
            x = (hi << 32) | lo
            reg[ops[0]] = ((x << n) >> 32) & M32 if f[1] == "l" else (x >> n) & M32
        elif base == "setp":
            cmp_, a, b = f[1], val(ops[1], bits), val(ops[2], bits)
            if ty.startswith("s"):
                a, b = _signed(a, bits), _signed(b, bits)
            reg[ops[0]] = {"eq": a == b, "ne": a != b, "lt": a < b, "le": a <= b, "gt": a > b, "ge": a >= b,
                           "lo": a < b, "ls": a <= b, "hi": a > b, "hs": a >= b}[cmp_]
        elif base == "selp":
            reg[ops[0]] = val(ops[1], bits) if reg[ops[3]] else val(ops[2], bits)
        elif base == "cvt":
            dbits, sbits = _TYPE_BITS[f[-2]], _TYPE_BITS[f[-1]]
            v = val(ops[1], sbits)
            if f[-1].startswith("s"):
                v = _signed(v, sbits)
            reg[ops[0]] = v & ((1 << dbits) - 1)
        else:
            raise NotImplementedError("ptx_exec: %s" % s)
    raise AssertionError("fell off the end of the kernel")


def launch(params, stmts, args, mem, grid, block):
    """every thread of a 1-D launch, one after the other (element-wise kernels: no barriers, no shared memory)"""
    for b in range(grid):
        for t in range(block):
            run_thread(params, stmts, args, mem, (t, 0, 0), (block, 1, 1), (b, 0, 0), (grid, 1, 1))
