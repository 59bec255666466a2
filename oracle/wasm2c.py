#!/usr/bin/env python3
"""wasm2c.py -- TEST INFRASTRUCTURE (oracle), not product code.

Translates the reference's own pre-generated WebAssembly images
(/root/reference/build/bn128_wasm.js, build/mnt6753_wasm.js: a base64 string
inside a JS wrapper) into one C file each, so that the reference's *own*
generated code (f1m_mul, g1m_multiexp2, g2m_multiexp, fft_fft, fft_ifft,
pol_constructLC, int_mul, ...) can be compiled with gcc and called natively
here and on the GPU box, where no JS engine / WASM runtime exists.

Outputs go ONLY to oracle/_ref/ (git-ignored): nothing of the reference is
copied into the repository history.  The translation is mechanical (one C
statement per WASM instruction); the images are integer-only MVP WASM with a
single imported memory, no tables, no globals, no floats.

Every translated function takes the instance's linear memory base `M` as its
first argument, so several instances (the reference's "workers",
src/bn128.js:50-171) can run on different threads.

C ABI of the generated library (prefix `w_`):
    uint8_t *wasm_new(uint32_t pages);   // allocate memory, apply data segments
    void     wasm_free(uint8_t *M);
    <ret>    w_<export>(uint8_t *M, uint32_t a0, ...);   // one per WASM export
"""
import base64
import re
import sys


class Reader:
    def __init__(self, b, pos=0):
        self.b = b
        self.p = pos

    def u8(self):
        v = self.b[self.p]
        self.p += 1
        return v

    def leb(self):
        r = 0
        s = 0
        while True:
            c = self.u8()
            r |= (c & 0x7F) << s
            s += 7
            if not (c & 0x80):
                return r

    def sleb(self, bits):
        r = 0
        s = 0
        while True:
            c = self.u8()
            r |= (c & 0x7F) << s
            s += 7
            if not (c & 0x80):
                if c & 0x40:
                    r -= 1 << s
                return r

    def bytes(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def name(self):
        return self.bytes(self.leb()).decode()


CT = {0x7F: "uint32_t", 0x7E: "uint64_t"}
PFX = {0x7F: "i", 0x7E: "j"}

# opcode -> (C operator, operand type, result type) for plain binary ops
I32, I64 = 0x7F, 0x7E
BINOPS = {
    0x6A: ("+", I32, I32), 0x6B: ("-", I32, I32), 0x6C: ("*", I32, I32),
    0x6E: ("/", I32, I32), 0x70: ("%", I32, I32),
    0x71: ("&", I32, I32), 0x72: ("|", I32, I32), 0x73: ("^", I32, I32),
    0x7C: ("+", I64, I64), 0x7D: ("-", I64, I64), 0x7E: ("*", I64, I64),
    0x80: ("/", I64, I64), 0x82: ("%", I64, I64),
    0x83: ("&", I64, I64), 0x84: ("|", I64, I64), 0x85: ("^", I64, I64),
    # unsigned compares
    0x46: ("==", I32, I32), 0x47: ("!=", I32, I32), 0x49: ("<", I32, I32),
    0x4B: (">", I32, I32), 0x4D: ("<=", I32, I32), 0x4F: (">=", I32, I32),
    0x51: ("==", I64, I32), 0x52: ("!=", I64, I32), 0x54: ("<", I64, I32),
    0x56: (">", I64, I32), 0x58: ("<=", I64, I32), 0x5A: (">=", I64, I32),
}
SIGNED_CMP = {
    0x48: ("<", I32), 0x4A: (">", I32), 0x4C: ("<=", I32), 0x4E: (">=", I32),
    0x53: ("<", I64), 0x55: (">", I64), 0x57: ("<=", I64), 0x59: (">=", I64),
}
SHIFTS = {
    0x74: ("<<", I32, False), 0x75: (">>", I32, True), 0x76: (">>", I32, False),
    0x86: ("<<", I64, False), 0x87: (">>", I64, True), 0x88: (">>", I64, False),
}
# loads: opcode -> (result type, C memory type)
LOADS = {
    0x28: (I32, "uint32_t"), 0x29: (I64, "uint64_t"),
    0x2C: (I32, "int8_t"), 0x2D: (I32, "uint8_t"),
    0x2E: (I32, "int16_t"), 0x2F: (I32, "uint16_t"),
    0x30: (I64, "int8_t"), 0x31: (I64, "uint8_t"),
    0x32: (I64, "int16_t"), 0x33: (I64, "uint16_t"),
    0x34: (I64, "int32_t"), 0x35: (I64, "uint32_t"),
}
STORES = {
    0x36: (I32, "uint32_t"), 0x37: (I64, "uint64_t"),
    0x3A: (I32, "uint8_t"), 0x3B: (I32, "uint16_t"),
    0x3C: (I64, "uint8_t"), 0x3D: (I64, "uint16_t"), 0x3E: (I64, "uint32_t"),
}


class Module:
    def __init__(self, wasm):
        assert wasm[:8] == b"\0asm\1\0\0\0", "not a wasm v1 module"
        self.types = []
        self.func_types = []      # type index per defined function
        self.nimports = 0
        self.exports = []         # (name, funcidx)
        self.bodies = []
        self.data = []            # (offset, bytes)
        self.mem_min = 0
        r = Reader(wasm, 8)
        while r.p < len(wasm):
            sid = r.u8()
            size = r.leb()
            end = r.p + size
            if sid == 1:
                for _ in range(r.leb()):
                    assert r.u8() == 0x60
                    params = [r.u8() for _ in range(r.leb())]
                    results = [r.u8() for _ in range(r.leb())]
                    self.types.append((params, results))
            elif sid == 2:
                for _ in range(r.leb()):
                    r.name(), r.name()
                    kind = r.u8()
                    assert kind == 2, "only a memory import is supported"
                    flags = r.u8()
                    self.mem_min = r.leb()
                    if flags & 1:
                        r.leb()
            elif sid == 3:
                self.func_types = [r.leb() for _ in range(r.leb())]
            elif sid == 7:
                for _ in range(r.leb()):
                    nm = r.name()
                    kind = r.u8()
                    idx = r.leb()
                    if kind == 0:
                        self.exports.append((nm, idx))
            elif sid == 10:
                for _ in range(r.leb()):
                    bsize = r.leb()
                    self.bodies.append(r.bytes(bsize))
            elif sid == 11:
                for _ in range(r.leb()):
                    assert r.leb() == 0
                    assert r.u8() == 0x41
                    off = r.sleb(32)
                    assert r.u8() == 0x0B
                    self.data.append((off & 0xFFFFFFFF, r.bytes(r.leb())))
            r.p = end


class FuncGen:
    def __init__(self, mod, fidx):
        self.mod = mod
        self.fidx = fidx
        self.params, self.results = mod.types[mod.func_types[fidx]]
        self.out = []
        self.stack = []           # list of value types
        self.ctrl = []            # dicts
        self.maxdepth = {I32: 0, I64: 0}
        self.nlabel = 0
        self.ind = 1

    def emit(self, s):
        self.out.append("    " * self.ind + s)

    def var(self, depth, t):
        self.maxdepth[t] = max(self.maxdepth[t], depth + 1)
        return "%s%d" % (PFX[t], depth)

    def push(self, t):
        self.stack.append(t)
        return self.var(len(self.stack) - 1, t)

    def pop(self, t=None):
        tt = self.stack.pop()
        if t is not None:
            assert tt == t, "type mismatch in f%d" % self.fidx
        return self.var(len(self.stack), tt), tt

    def top(self):
        return self.var(len(self.stack) - 1, self.stack[-1])

    def branch(self, depth):
        """C statement(s) for a branch to the depth-th enclosing label."""
        c = self.ctrl[-1 - depth]
        c["used"] = True
        s = ""
        if c["kind"] != "loop" and c["rt"] is not None:
            src = self.top()
            dst = self.var(c["height"], c["rt"])
            if src != dst:
                s = "%s = %s; " % (dst, src)
        return s + "goto L%d;" % c["label"]

    def gen(self):
        body = self.mod.bodies[self.fidx]
        r = Reader(body)
        ltypes = list(self.params)
        for _ in range(r.leb()):
            n = r.leb()
            t = r.u8()
            ltypes += [t] * n
        self.ltypes = ltypes
        self.ctrl.append(dict(kind="func", label=self._newlabel(), rt=self.results[0] if self.results else None,
                              height=0, used=False))
        dead = 0   # >0: skipping unreachable code, counts nested blocks opened while dead
        while r.p < len(body):
            op = r.u8()
            if dead:
                # parse immediates, track nesting, emit nothing
                if op in (0x02, 0x03, 0x04):
                    r.u8()
                    dead += 1
                elif op == 0x0B:
                    dead -= 1
                    if dead == 0:
                        if len(self.ctrl) == 1:
                            break            # function body ended in unreachable code
                        self._end(r, reachable=False)
                elif op == 0x05:
                    if dead == 1:
                        dead = 0
                        self._else()
                elif op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22):
                    r.leb()
                elif 0x28 <= op <= 0x3E:
                    r.leb(), r.leb()
                elif op == 0x41:
                    r.sleb(32)
                elif op == 0x42:
                    r.sleb(64)
                continue
            if op == 0x00:
                self.emit("__builtin_trap();")
                dead = 1
            elif op == 0x01:
                pass
            elif op in (0x02, 0x03):
                bt = r.u8()
                rt = None if bt == 0x40 else bt
                lab = self._newlabel()
                kind = "block" if op == 0x02 else "loop"
                self.ctrl.append(dict(kind=kind, label=lab, rt=rt, height=len(self.stack), used=False))
                if kind == "loop":
                    self.emit("L%d: ;" % lab)
            elif op == 0x04:
                bt = r.u8()
                rt = None if bt == 0x40 else bt
                c, _ = self.pop(I32)
                lab = self._newlabel()
                self.ctrl.append(dict(kind="if", label=lab, rt=rt, height=len(self.stack), used=False, haselse=False))
                self.emit("if (%s) {" % c)
                self.ind += 1
            elif op == 0x05:
                self._else()
            elif op == 0x0B:
                if len(self.ctrl) == 1:
                    # function end
                    c = self.ctrl[-1]
                    if c["rt"] is not None:
                        v, _ = self.pop(c["rt"])
                        dst = self.var(0, c["rt"])
                        if v != dst:
                            self.emit("%s = %s;" % (dst, v))
                    break
                self._end(r, reachable=True)
            elif op == 0x0C:
                self.emit(self.branch(r.leb()))
                dead = 1
            elif op == 0x0D:
                d = r.leb()
                c, _ = self.pop(I32)
                self.emit("if (%s) { %s }" % (c, self.branch(d)))
            elif op == 0x0F:
                self.emit(self.branch(len(self.ctrl) - 1))
                dead = 1
            elif op == 0x10:
                f = r.leb() - self.mod.nimports
                params, results = self.mod.types[self.mod.func_types[f]]
                args = []
                for t in reversed(params):
                    v, _ = self.pop(t)
                    args.append(v)
                args.reverse()
                call = "f%d(%s)" % (f, ", ".join(["M"] + args))
                if results:
                    self.emit("%s = %s;" % (self.push(results[0]), call))
                else:
                    self.emit(call + ";")
            elif op == 0x1A:
                self.pop()
            elif op == 0x1B:
                c, _ = self.pop(I32)
                b, t = self.pop()
                a, _ = self.pop(t)
                d = self.push(t)
                self.emit("%s = %s ? %s : %s;" % (d, c, a, b))
            elif op == 0x20:
                n = r.leb()
                self.emit("%s = l%d;" % (self.push(self.ltypes[n]), n))
            elif op == 0x21:
                n = r.leb()
                v, _ = self.pop(self.ltypes[n])
                self.emit("l%d = %s;" % (n, v))
            elif op == 0x22:
                n = r.leb()
                self.emit("l%d = %s;" % (n, self.top()))
            elif op in LOADS:
                r.leb()
                off = r.leb()
                rt, mt = LOADS[op]
                a, _ = self.pop(I32)
                d = self.push(rt)
                self.emit("{ %s t_; __builtin_memcpy(&t_, M + (uint64_t)%s + %du, sizeof t_); %s = (%s)t_; }"
                          % (mt, a, off, d, CT[rt] if mt[0] == "u" else ("int32_t" if rt == I32 else "int64_t")))
            elif op in STORES:
                r.leb()
                off = r.leb()
                vt, mt = STORES[op]
                v, _ = self.pop(vt)
                a, _ = self.pop(I32)
                self.emit("{ %s t_ = (%s)%s; __builtin_memcpy(M + (uint64_t)%s + %du, &t_, sizeof t_); }"
                          % (mt, mt, v, a, off))
            elif op == 0x41:
                self.emit("%s = %du;" % (self.push(I32), r.sleb(32) & 0xFFFFFFFF))
            elif op == 0x42:
                self.emit("%s = %duLL;" % (self.push(I64), r.sleb(64) & 0xFFFFFFFFFFFFFFFF))
            elif op == 0x45:
                a, _ = self.pop(I32)
                self.emit("%s = (%s == 0);" % (self.push(I32), a))
            elif op == 0x50:
                a, _ = self.pop(I64)
                self.emit("%s = (%s == 0);" % (self.push(I32), a))
            elif op in BINOPS:
                cop, ot, rt = BINOPS[op]
                b, _ = self.pop(ot)
                a, _ = self.pop(ot)
                self.emit("%s = (%s)(%s %s %s);" % (self.push(rt), CT[rt], a, cop, b))
            elif op in SIGNED_CMP:
                cop, ot = SIGNED_CMP[op]
                st = "int32_t" if ot == I32 else "int64_t"
                b, _ = self.pop(ot)
                a, _ = self.pop(ot)
                self.emit("%s = ((%s)%s %s (%s)%s);" % (self.push(I32), st, a, cop, st, b))
            elif op in SHIFTS:
                cop, t, signed = SHIFTS[op]
                mask = 31 if t == I32 else 63
                b, _ = self.pop(t)
                a, _ = self.pop(t)
                d = self.push(t)
                if signed:
                    st = "int32_t" if t == I32 else "int64_t"
                    self.emit("%s = (%s)((%s)%s >> (%s & %d));" % (d, CT[t], st, a, b, mask))
                else:
                    self.emit("%s = %s %s (%s & %d);" % (d, a, cop, b, mask))
            elif op in (0x77, 0x78, 0x89, 0x8A):
                t = I32 if op < 0x80 else I64
                w = 32 if t == I32 else 64
                left = op in (0x77, 0x89)
                b, _ = self.pop(t)
                a, _ = self.pop(t)
                d = self.push(t)
                sh = "(%s & %d)" % (b, w - 1)
                inv = "((%d - %s) & %d)" % (w, sh, w - 1)
                if left:
                    self.emit("%s = (%s << %s) | (%s >> %s);" % (d, a, sh, a, inv))
                else:
                    self.emit("%s = (%s >> %s) | (%s << %s);" % (d, a, sh, a, inv))
            elif op == 0xA7:
                a, _ = self.pop(I64)
                self.emit("%s = (uint32_t)%s;" % (self.push(I32), a))
            elif op == 0xAC:
                a, _ = self.pop(I32)
                self.emit("%s = (uint64_t)(int64_t)(int32_t)%s;" % (self.push(I64), a))
            elif op == 0xAD:
                a, _ = self.pop(I32)
                self.emit("%s = (uint64_t)%s;" % (self.push(I64), a))
            else:
                raise NotImplementedError("opcode 0x%02x in f%d" % (op, self.fidx))
        # assemble
        rt = CT[self.results[0]] if self.results else "void"
        ps = ", ".join(["uint8_t *M"] + ["%s l%d" % (CT[t], i) for i, t in enumerate(self.params)])
        lines = ["static %s f%d(%s) {" % (rt, self.fidx, ps)]
        for i in range(len(self.params), len(self.ltypes)):
            lines.append("    %s l%d = 0;" % (CT[self.ltypes[i]], i))
        for t in (I32, I64):
            if self.maxdepth[t]:
                lines.append("    %s %s;" % (CT[t], ", ".join("%s%d = 0" % (PFX[t], d) for d in range(self.maxdepth[t]))))
        lines += self.out
        lines.append("L%d: ;" % self.ctrl[0]["label"])
        if self.results:
            lines.append("    return %s0;" % PFX[self.results[0]])
        lines.append("}")
        return "\n".join(lines)

    def _newlabel(self):
        self.nlabel += 1
        return self.nlabel

    def _else(self):
        c = self.ctrl[-1]
        assert c["kind"] == "if"
        c["haselse"] = True
        # value (if any) already sits in the right slot by construction
        del self.stack[c["height"]:]
        self.ind -= 1
        self.emit("} else {")
        self.ind += 1

    def _end(self, r, reachable):
        c = self.ctrl.pop()
        del self.stack[c["height"]:]
        if c["rt"] is not None:
            self.stack.append(c["rt"])
            self.var(c["height"], c["rt"])
        if c["kind"] == "if":
            self.ind -= 1
            self.emit("}")
        if c["kind"] != "loop":
            self.emit("L%d: ;" % c["label"])


def translate(wasm, tag):
    mod = Module(wasm)
    out = ["/* GENERATED by oracle/wasm2c.py from the reference's %s image -- do not commit. */" % tag,
           "#include <stdint.h>", "#include <stdlib.h>", "#include <string.h>", "#include <sys/mman.h>", ""]
    for i in range(len(mod.func_types)):
        params, results = mod.types[mod.func_types[i]]
        rt = CT[results[0]] if results else "void"
        ps = ", ".join(["uint8_t *M"] + [CT[t] for t in params])
        out.append("static %s f%d(%s);" % (rt, i, ps))
    out.append("")
    for i in range(len(mod.func_types)):
        out.append(FuncGen(mod, i).gen())
        out.append("")
    for nm, idx in mod.exports:
        params, results = mod.types[mod.func_types[idx]]
        rt = CT[results[0]] if results else "void"
        ps = ", ".join(["uint8_t *M"] + ["%s a%d" % (CT[t], i) for i, t in enumerate(params)])
        args = ", ".join(["M"] + ["a%d" % i for i in range(len(params))])
        out.append("%s w_%s(%s) { %sf%d(%s); }" % (rt, nm, ps, "return " if results else "", idx, args))
    out.append("")
    blob = bytearray()
    segs = []
    for off, data in mod.data:
        segs.append((off, len(blob), len(data)))
        blob += data
    out.append("static const uint8_t data_blob[%d] = {" % max(1, len(blob)))
    for i in range(0, len(blob), 32):
        out.append("    " + ",".join(str(b) for b in blob[i:i + 32]) + ",")
    out.append("};")
    out.append("static const struct { uint32_t off, src, len; } data_segs[%d] = {" % max(1, len(segs)))
    for s in segs:
        out.append("    {%du, %du, %du}," % s)
    out.append("};")
    out.append("""
uint32_t wasm_min_pages(void) { return %du; }
uint8_t *wasm_new(uint32_t pages) {
    size_t sz = (size_t)pages << 16;
    uint8_t *M = mmap(0, sz, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (M == MAP_FAILED) return 0;
    for (unsigned i = 0; i < %du; i++) memcpy(M + data_segs[i].off, data_blob + data_segs[i].src, data_segs[i].len);
    return M;
}
void wasm_free(uint8_t *M, uint32_t pages) { munmap(M, (size_t)pages << 16); }
""" % (mod.mem_min, len(segs)))
    return "\n".join(out), mod


def extract_wasm(js_path):
    src = open(js_path).read()
    m = re.search(r'Buffer\.from\("([A-Za-z0-9+/=]+)",\s*"base64"\)', src)
    consts = dict((k, int(v)) for k, v in re.findall(r"exports\.(\w+)\s*=\s*(\d+);", src))
    return base64.b64decode(m.group(1)), consts


if __name__ == "__main__":
    js, out_c = sys.argv[1], sys.argv[2]
    wasm, consts = extract_wasm(js)
    tag = js.split("/")[-1]
    code, mod = translate(wasm, tag)
    with open(out_c, "w") as f:
        f.write(code)
        f.write("\n/* data-segment pointers exported by the JS wrapper */\n")
        for k, v in consts.items():
            f.write("uint32_t wasm_%s(void) { return %du; }\n" % (k, v))
    print("%s: %d functions, %d exports, %d data segments -> %s" % (tag, len(mod.func_types), len(mod.exports), len(mod.data), out_c))
