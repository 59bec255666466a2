"""ref_wasm.py -- TEST INFRASTRUCTURE (oracle), not product code.

ctypes host for oracle/_ref/libbn128_ref.so, i.e. for the reference's OWN
generated WebAssembly image (/root/reference/build/bn128_wasm.js) translated
to C by oracle/wasm2c.py and compiled with gcc.  It replays, call for call,
what the reference's JavaScript host does around that image:

  * heap / putBin / getBin                      src/bn128.js:302-317
  * worker commands G1_MULTIEXP / G2_MULTIEXP / CALC_H   src/bn128.js:102-166
  * g1_multiexp / g2_multiexp range split + serial combine   src/bn128.js:353-415
  * groth16GenProof                             src/bn128.js:580-720

so the arithmetic that runs is the reference's (bit-exact by construction);
only the 150 lines of JS glue are restated.  r and s (random in the reference,
src/bn128.js:642-664) are injectable, r = s = 0 being the reference's own
debug mode (src/bn128.js:666-668) which example/bn128/proof_good0.json pins.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.
"""
import ctypes
import os
import struct

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libbn128_ref.so")

PAGES = 5000            # src/bn128.js:183,256 (320 MB per instance)


def available():
    return os.path.exists(_LIB)


class RefInstance:
    """One WASM instance = one reference worker (src/bn128.js:50-171)."""

    def __init__(self, pages=PAGES, lib_path=_LIB):
        self.lib = ctypes.CDLL(lib_path)
        self.lib.wasm_new.restype = ctypes.c_void_p
        self.lib.wasm_new.argtypes = [ctypes.c_uint32]
        self.lib.wasm_free.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        self.pages = pages
        self.M = self.lib.wasm_new(pages)
        if not self.M:
            raise MemoryError("wasm_new failed")
        self.size = pages << 16
        self.mem = (ctypes.c_uint8 * self.size).from_address(self.M)
        self._fn = {}

    def close(self):
        if self.M:
            self.lib.wasm_free(self.M, self.pages)
            self.M = None

    def call(self, name, *args):
        f = self._fn.get(name)
        if f is None:
            f = getattr(self.lib, "w_" + name)
            f.restype = ctypes.c_uint32
            self._fn[name] = f
        return f(ctypes.c_void_p(self.M), *[ctypes.c_uint32(a) for a in args])

    def const(self, name):
        f = getattr(self.lib, "wasm_" + name)
        f.restype = ctypes.c_uint32
        return f()

    # --- heap (src/bn128.js:68-77, :302-317) ---
    @property
    def heap(self):
        return struct.unpack_from("<I", self.mem, 0)[0]

    @heap.setter
    def heap(self, v):
        struct.pack_into("<I", self.mem, 0, v)

    def alloc(self, length):
        h = self.heap
        while h & 3:
            h += 1
        res = h
        h += length
        if h > self.size:
            raise MemoryError("reference heap exhausted")
        self.heap = h
        return res

    def put(self, p, b):
        ctypes.memmove(self.M + p, bytes(b) if not isinstance(b, (bytes, bytearray)) else (ctypes.c_char * len(b)).from_buffer_copy(b), len(b))

    def put_bin(self, b):
        p = self.alloc(len(b))
        self.put(p, b)
        return p

    def get(self, p, l):
        return ctypes.string_at(self.M + p, l)

    # --- worker commands (src/bn128.js:102-166) ---
    def cmd_g1_multiexp(self, scalars, points, n, variant="multiexp2", w=7):
        old = self.heap
        ps = self.put_bin(scalars)
        pp = self.put_bin(points)
        pr = self.alloc(96)
        self.call("g1m_zero", pr)
        self.call("g1m_" + variant, ps, pp, n, w, pr)
        res = self.get(pr, 96)
        self.heap = old
        return res

    def cmd_g2_multiexp(self, scalars, points, n, variant="multiexp", w=7):
        old = self.heap
        ps = self.put_bin(scalars)
        pp = self.put_bin(points)
        pr = self.alloc(192)
        self.call("g2m_zero", pr)
        self.call("g2m_" + variant, ps, pp, n, w, pr)
        res = self.get(pr, 192)
        self.heap = old
        return res

    def cmd_calc_h(self, signals, polsA, polsB, nSignals, domainSize):
        old = self.heap
        c = self.call
        pSignals = self.put_bin(signals)
        pPolsA = self.put_bin(polsA)
        pPolsB = self.put_bin(polsB)
        pSignalsM = self.alloc(nSignals * 32)
        pPolA = self.alloc(domainSize * 32)
        pPolB = self.alloc(domainSize * 32)
        pPolA2 = self.alloc(domainSize * 64)
        pPolB2 = self.alloc(domainSize * 64)
        c("fft_toMontgomeryN", pSignals, pSignalsM, nSignals)
        c("pol_zero", pPolA, domainSize)
        c("pol_zero", pPolB, domainSize)
        c("pol_constructLC", pPolsA, pSignalsM, nSignals, pPolA)
        c("pol_constructLC", pPolsB, pSignalsM, nSignals, pPolB)
        c("fft_copyNInterleaved", pPolA, pPolA2, domainSize)
        c("fft_copyNInterleaved", pPolB, pPolB2, domainSize)
        c("fft_ifft", pPolA, domainSize, 0)
        c("fft_ifft", pPolB, domainSize, 0)
        c("fft_fft", pPolA, domainSize, 1)
        c("fft_fft", pPolB, domainSize, 1)
        c("fft_copyNInterleaved", pPolA, pPolA2 + 32, domainSize)
        c("fft_copyNInterleaved", pPolB, pPolB2 + 32, domainSize)
        c("fft_mulN", pPolA2, pPolB2, domainSize * 2, pPolA2)
        c("fft_ifft", pPolA2, domainSize * 2, 0)
        c("fft_fromMontgomeryN", pPolA2 + domainSize * 32, pPolA2 + domainSize * 32, domainSize)
        res = self.get(pPolA2 + domainSize * 32, domainSize * 32)
        self.heap = old
        return res


def bin2int(b):
    return int.from_bytes(b, "little")


class RefBn128:
    """The reference's `Bn128` host class (src/bn128.js:267-720), serial workers."""

    def __init__(self, concurrency=8, pages=PAGES):
        self.main = RefInstance(pages=1000)
        self.pr0 = self.main.alloc(192)
        self.pr1 = self.main.alloc(192)
        self.workers = [RefInstance(pages=pages) for _ in range(1)]
        self.concurrency = concurrency      # split count (src/bn128.js:211-215: 8 under Node)

    def close(self):
        self.main.close()
        for w in self.workers:
            w.close()

    def _multiexp(self, g, scalars, points, psize, variant=None):
        nPoints = len(scalars) // 32
        W = self.concurrency
        per = nPoints // W
        results = []
        for i in range(W):
            n = per if i < W - 1 else nPoints - per * (W - 1)
            s = scalars[i * per * 32: i * per * 32 + n * 32]
            p = points[i * per * psize: i * per * psize + n * psize]
            wk = self.workers[0]
            if g == 1:
                results.append(wk.cmd_g1_multiexp(s, p, n, *( [variant] if variant else [])))
            else:
                results.append(wk.cmd_g2_multiexp(s, p, n, *( [variant] if variant else [])))
        m = self.main
        pfx = "g1m" if g == 1 else "g2m"
        size = 96 if g == 1 else 192
        m.call(pfx + "_zero", self.pr0)
        for rbin in results:
            m.put(self.pr1, rbin)
            m.call(pfx + "_add", self.pr0, self.pr1, self.pr0)
        return m.get(self.pr0, size)

    def g1_multiexp(self, scalars, points):
        return self._multiexp(1, scalars, points, 64)

    def g2_multiexp(self, scalars, points):
        return self._multiexp(2, scalars, points, 128)

    def calcH(self, signals, polsA, polsB, nSignals, domainSize):
        return self.workers[0].cmd_calc_h(signals, polsA, polsB, nSignals, domainSize)

    def _msm_results(self, signals, sec, nPublic):
        h = self.calcH(signals, sec["polsA"], sec["polsB"], sec["nSignals"], sec["domainSize"])
        rH = self.g1_multiexp(h, sec["pointsH"])
        rA = self.g1_multiexp(signals, sec["pointsA"])
        rB1 = self.g1_multiexp(signals, sec["pointsB1"])
        rB2 = self.g2_multiexp(signals, sec["pointsB2"])
        rC = self.g1_multiexp(signals[(nPublic + 1) * 32:], sec["pointsC"])
        return rH, rA, rB1, rB2, rC

    def g1_affine_std(self, p96):
        """g1m_affine + g1m_fromMontgomery (src/bn128.js:706-712)."""
        m = self.main
        m.put(self.pr0, p96)
        m.call("g1m_affine", self.pr0, self.pr0)
        m.call("g1m_fromMontgomery", self.pr0, self.pr0)
        return m.get(self.pr0, 96)

    def g2_affine_std(self, p192):
        m = self.main
        m.put(self.pr0, p192)
        m.call("g2m_affine", self.pr0, self.pr0)
        m.call("g2m_fromMontgomery", self.pr0, self.pr0)
        return m.get(self.pr0, 192)

    def groth16GenProof(self, signals, pkey, r=None, s=None, raw=False):
        """src/bn128.js:580-720.  r, s: 32-byte LE strings (None => os.urandom)."""
        hdr = struct.unpack_from("<10I", pkey, 0)
        nSignals, nPublic, domainSize, pPolsA, pPolsB, pPointsA, pPointsB1, pPointsB2, pPointsC, pHExps = hdr
        polsA = pkey[pPolsA:pPolsB]          # reference slices over-long (:592); contents beyond are unused
        polsB = pkey[pPolsB:pPointsA]
        pointsA = pkey[pPointsA:pPointsA + nSignals * 64]
        pointsB1 = pkey[pPointsB1:pPointsB1 + nSignals * 64]
        pointsB2 = pkey[pPointsB2:pPointsB2 + nSignals * 128]
        pointsC = pkey[pPointsC:pPointsC + (nSignals - nPublic - 1) * 64]
        pointsH = pkey[pHExps:pHExps + domainSize * 64]
        alfa1 = pkey[40:104]
        beta1 = pkey[104:168]
        delta1 = pkey[168:232]
        beta2 = pkey[232:360]
        delta2 = pkey[360:488]

        sec = {"polsA": polsA, "polsB": polsB, "nSignals": nSignals, "domainSize": domainSize, "pointsA": pointsA,
               "pointsB1": pointsB1, "pointsB2": pointsB2, "pointsC": pointsC, "pointsH": pointsH}
        rH, rA, rB1, rB2, rC = self._msm_results(signals, sec, nPublic)

        m = self.main
        c = m.call
        old = m.heap
        pi_a = m.alloc(96)
        pi_b = m.alloc(192)
        pi_c = m.alloc(96)
        pib1 = m.alloc(96)
        m.put(pi_a, rA)
        m.put(pib1, rB1)
        m.put(pi_b, rB2)
        m.put(pi_c, rC)

        def load1(b):
            p = m.alloc(96)
            m.put(p, b)
            c("f1m_one", p + 64)
            return p

        def load2(b):
            p = m.alloc(192)
            m.put(p, b)
            c("f2m_one", p + 128)
            return p

        pAlfa1, pBeta1, pDelta1 = load1(alfa1), load1(beta1), load1(delta1)
        pBeta2, pDelta2 = load2(beta2), load2(delta2)
        aux1 = m.alloc(96)
        aux2 = m.alloc(192)
        pr = m.alloc(32)
        ps = m.alloc(32)
        if r is None:
            r = os.urandom(32)
        if s is None:
            s = os.urandom(32)
        m.put(pr, r)
        m.put(ps, s)

        c("g1m_add", pAlfa1, pi_a, pi_a)
        c("g1m_timesScalar", pDelta1, pr, 32, aux1)
        c("g1m_add", aux1, pi_a, pi_a)

        c("g2m_add", pBeta2, pi_b, pi_b)
        c("g2m_timesScalar", pDelta2, ps, 32, aux2)
        c("g2m_add", aux2, pi_b, pi_b)

        c("g1m_add", pBeta1, pib1, pib1)
        c("g1m_timesScalar", pDelta1, ps, 32, aux1)
        c("g1m_add", aux1, pib1, pib1)

        m.put(aux1, rH)
        c("g1m_add", aux1, pi_c, pi_c)
        c("g1m_timesScalar", pi_a, ps, 32, aux1)
        c("g1m_add", aux1, pi_c, pi_c)
        c("g1m_timesScalar", pib1, pr, 32, aux1)
        c("g1m_add", aux1, pi_c, pi_c)
        prs = m.alloc(64)
        c("int_mul", pr, ps, prs)
        c("g1m_timesScalar", pDelta1, prs, 64, aux1)
        c("g1m_neg", aux1, aux1)
        c("g1m_add", aux1, pi_c, pi_c)

        c("g1m_affine", pi_a, pi_a)
        c("g2m_affine", pi_b, pi_b)
        c("g1m_affine", pi_c, pi_c)
        c("g1m_fromMontgomery", pi_a, pi_a)
        c("g2m_fromMontgomery", pi_b, pi_b)
        c("g1m_fromMontgomery", pi_c, pi_c)
        ba, bb, bc = m.get(pi_a, 96), m.get(pi_b, 192), m.get(pi_c, 96)
        m.heap = old
        if raw:
            return ba, bb, bc
        return format_proof(ba, bb, bc)


class RefBn128Pool(RefBn128):
    """RefBn128 with the reference's worker pool (src/bn128.js:173-265): W instances of the image, one host
    thread each (ctypes releases the GIL inside the image).  Task graph of groth16GenProof as the reference
    builds it (src/bn128.js:607-622): CALC_H is ONE task on one worker (:569-578); every multiexp is split
    into W contiguous ranges (:353-372), one task each; the H multiexp is queued when CALC_H resolves; the
    partial points are added serially on the main instance (:376-380).  Used by the golden-vector scripts and
    as bench.py's CPU baseline (Web-Worker-parallel arm); concurrency=1 is the single-threaded arm."""

    def __init__(self, concurrency=8, pages=PAGES):
        from concurrent.futures import ThreadPoolExecutor
        import queue
        self.main = RefInstance(pages=1000)
        self.pr0 = self.main.alloc(192)
        self.pr1 = self.main.alloc(192)
        self.concurrency = max(1, concurrency)
        self.workers = [RefInstance(pages=pages) for _ in range(self.concurrency)]
        self._free = queue.Queue()
        for w in self.workers:
            self._free.put(w)
        self.ex = ThreadPoolExecutor(max_workers=self.concurrency)

    def close(self):
        self.ex.shutdown(wait=True)
        super().close()

    def _task(self, fn):
        """postAction (src/bn128.js:272-300): the task runs on whichever worker is free."""
        def run():
            wk = self._free.get()
            try:
                return fn(wk)
            finally:
                self._free.put(wk)
        return self.ex.submit(run)

    def _multiexp_async(self, g, scalars, points, psize):
        nPoints = len(scalars) // 32
        W = self.concurrency
        per = nPoints // W
        futs = []
        for i in range(W):
            n = per if i < W - 1 else nPoints - per * (W - 1)
            s = scalars[i * per * 32: i * per * 32 + n * 32]
            p = points[i * per * psize: i * per * psize + n * psize]
            futs.append(self._task((lambda wk, s=s, p=p, n=n: wk.cmd_g1_multiexp(s, p, n)) if g == 1 else
                                   (lambda wk, s=s, p=p, n=n: wk.cmd_g2_multiexp(s, p, n))))
        return futs

    def _combine(self, g, futs):
        m = self.main
        pfx = "g1m" if g == 1 else "g2m"
        m.call(pfx + "_zero", self.pr0)
        for f in futs:
            m.put(self.pr1, f.result())
            m.call(pfx + "_add", self.pr0, self.pr1, self.pr0)
        return m.get(self.pr0, 96 if g == 1 else 192)

    def _multiexp(self, g, scalars, points, psize, variant=None):
        return self._combine(g, self._multiexp_async(g, scalars, points, psize))

    def calcH(self, signals, polsA, polsB, nSignals, domainSize):
        return self._task(lambda wk: wk.cmd_calc_h(signals, polsA, polsB, nSignals, domainSize)).result()

    def _msm_results(self, signals, sec, nPublic):
        """the five multiexps with CALC_H in flight beside them (src/bn128.js:607-622)."""
        fh = self._task(lambda wk: wk.cmd_calc_h(signals, sec["polsA"], sec["polsB"], sec["nSignals"], sec["domainSize"]))
        fA = self._multiexp_async(1, signals, sec["pointsA"], 64)
        fB1 = self._multiexp_async(1, signals, sec["pointsB1"], 64)
        fB2 = self._multiexp_async(2, signals, sec["pointsB2"], 128)
        fC = self._multiexp_async(1, signals[(nPublic + 1) * 32:], sec["pointsC"], 64)
        h = fh.result()
        fH = self._multiexp_async(1, h, sec["pointsH"], 64)
        return self._combine(1, fH), self._combine(1, fA), self._combine(1, fB1), self._combine(2, fB2), self._combine(1, fC)


def groth16_verify(verification_key, inputs, proof, inst=None):
    """Bn128.groth16Verify (src/bn128.js:722-791) replayed on the reference image: the reference's own
    acceptance check for proofs with random r, s (test/bn128_prover.js:73-78).  Returns True / False."""
    own = inst is None
    m = inst or RefInstance(pages=1000)
    c = m.call
    old = m.heap
    S = 32

    def al(n):
        h = m.heap
        while h & 7:                      # setInt requires 8-byte alignment (src/bn128.js:459)
            h += 1
        m.heap = h
        return m.alloc(n)

    def set_f1(p, v):
        m.put(p, int(v).to_bytes(32, "little"))

    def set_g1(p, e):                     # setG1Affine, src/bn128.js:503-507
        set_f1(p, e[0]); set_f1(p + S, e[1]); set_f1(p + 2 * S, 1)

    def set_g2(p, e):                     # setG2Affine, src/bn128.js:509-513
        for k in range(2):
            set_f1(p + 2 * S * k, e[k][0]); set_f1(p + 2 * S * k + S, e[k][1])
        set_f1(p + 4 * S, 1); set_f1(p + 5 * S, 0)
    pA, pB, pIC, pICaux, pICr, pC = al(96), al(192), al(96), al(96), al(32), al(96)
    pDelta2, pGamma2, pAlfa1, pBeta2 = al(192), al(192), al(96), al(192)
    set_g1(pA, proof["pi_a"]); c("g1m_toMontgomery", pA, pA)
    set_g2(pB, proof["pi_b"]); c("g2m_toMontgomery", pB, pB)
    set_g1(pC, proof["pi_c"]); c("g1m_toMontgomery", pC, pC)
    set_g2(pDelta2, verification_key["vk_delta_2"]); c("g2m_toMontgomery", pDelta2, pDelta2)
    set_g2(pGamma2, verification_key["vk_gamma_2"]); c("g2m_toMontgomery", pGamma2, pGamma2)
    set_g1(pAlfa1, verification_key["vk_alfa_1"]); c("g1m_toMontgomery", pAlfa1, pAlfa1)
    set_g2(pBeta2, verification_key["vk_beta_2"]); c("g2m_toMontgomery", pBeta2, pBeta2)
    set_g1(pIC, verification_key["IC"][0]); c("g1m_toMontgomery", pIC, pIC)
    ok = True
    for i, v in enumerate(inputs):
        set_g1(pICaux, verification_key["IC"][i + 1]); c("g1m_toMontgomery", pICaux, pICaux)
        set_f1(pICr, v)
        if c("int_gte", pICr, m.const("pr")):
            ok = False
            break
        c("g1m_timesScalar", pICaux, pICr, 32, pICaux)
        c("g1m_add", pICaux, pIC, pIC)
    valid = 0
    if ok:
        c("g1m_affine", pIC, pIC)
        c("g1m_neg", pIC, pIC); c("g1m_neg", pC, pC); c("g1m_neg", pAlfa1, pAlfa1)
        valid = c("bn128_pairingEq4", pA, pB, pIC, pGamma2, pC, pDelta2, pAlfa1, pBeta2, m.const("pOneT"))
    m.heap = old
    if own:
        m.close()
    return valid == 1


def format_proof(ba, bb, bc):
    """bin2g1 / bin2g2 (src/bn128.js:319-351) + the returned object (:714-718)."""
    g = lambda b, i: str(bin2int(b[32 * i:32 * i + 32]))
    return {
        "pi_a": [g(ba, 0), g(ba, 1), g(ba, 2)],
        "pi_b": [[g(bb, 0), g(bb, 1)], [g(bb, 2), g(bb, 3)], [g(bb, 4), g(bb, 5)]],
        "pi_c": [g(bc, 0), g(bc, 1), g(bc, 2)],
    }
