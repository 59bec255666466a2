"""Host-side mirror of the reference's `Bn128` class (/root/reference/src/bn128.js:267-720)
over the CUDA library.  Same method names, argument meaning and result layouts, so code and
tests written against the reference read the same here:

    bn128 = build_bn128()                      # index.js:20-22 `buildBn128()`
    r = bn128.g1_multiexp(scalars, points)     # src/bn128.js:353-383 -> 96 bytes X|Y|Z Montgomery
    r = bn128.g2_multiexp(scalars, points)     # src/bn128.js:385-415 -> 192 bytes

Difference, by design: calls are synchronous (the reference returns Promises resolved by its
worker pool).  Like the reference, multiexp returns an un-normalised Jacobian triple; use
g1_affine / g2_affine before comparing.
"""
import ctypes

from . import _lib
from ._lib import Context, as_ptr


class Bn128:
    # src/bn128.js:176-177
    q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    n8 = 32

    def __init__(self, device=0):
        self.ctx = Context(device)
        self.lib = self.ctx.lib

    # -- Bn128.terminate(), src/bn128.js:562-566
    def terminate(self):
        self.ctx.close()

    def _multiexp(self, fn, scalars, points, psize, osize):
        if len(scalars) % 32:
            raise ValueError("scalars must be n*32 bytes")
        n = len(scalars) // 32
        if len(points) < n * psize:
            raise ValueError("points buffer too short for %d scalars" % n)
        out = ctypes.create_string_buffer(osize)
        self.ctx.check(fn(self.ctx.h, as_ptr(scalars), as_ptr(points), n, ctypes.cast(out, ctypes.c_void_p)))
        return out.raw

    def g1_multiexp(self, scalars, points):
        """src/bn128.js:353-383: scalars n*32 B raw LE, points n*64 B affine Montgomery."""
        return self._multiexp(self.lib.ws_bn128_g1_multiexp, scalars, points, 64, 96)

    def g2_multiexp(self, scalars, points):
        """src/bn128.js:385-415: points n*128 B."""
        return self._multiexp(self.lib.ws_bn128_g2_multiexp, scalars, points, 128, 192)

    def _affine(self, fn, p, size, from_montgomery):
        n = len(p) // size
        out = ctypes.create_string_buffer(max(1, n * size))
        self.ctx.check(fn(self.ctx.h, as_ptr(p), ctypes.cast(out, ctypes.c_void_p), n, from_montgomery))
        return out.raw[:n * size]

    def g1_affine(self, p):
        """src/bn128.js:417-421 (accepts one or several 96-byte triples)."""
        return self._affine(self.lib.ws_bn128_g1_affine, p, 96, 0)

    def g2_affine(self, p):
        """src/bn128.js:423-427."""
        return self._affine(self.lib.ws_bn128_g2_affine, p, 192, 0)

    def g1_fromMontgomery(self, p):
        """src/bn128.js:429-433 -> g1m_fromMontgomery (build_curve_jacobian_a0.js:258-277): every coordinate of the 96-byte
        triple(s) out of Montgomery form, nothing else (no normalisation)."""
        if len(p) % 96:
            raise ValueError("whole 96-byte G1 triples expected")
        return self.fp_batch(_lib.FIELD_BN128_FQ, _lib.OP_FROMMONT, bytes(p))

    def g2_fromMontgomery(self, p):
        """src/bn128.js:435-439 -> g2m_fromMontgomery: the six Fq components of each 192-byte triple."""
        if len(p) % 192:
            raise ValueError("whole 192-byte G2 triples expected")
        return self.fp_batch(_lib.FIELD_BN128_FQ, _lib.OP_FROMMONT, bytes(p))

    def g1_affine_std(self, p):
        """g1_affine followed by g1_fromMontgomery (src/bn128.js:706-712)."""
        return self._affine(self.lib.ws_bn128_g1_affine, p, 96, 1)

    def g2_affine_std(self, p):
        return self._affine(self.lib.ws_bn128_g2_affine, p, 192, 1)

    # -- NTT (WASM exports fft_fft / fft_ifft, src/build_fft.js:159-221); returns the transformed buffer
    def fft(self, px, odd=0):
        buf = ctypes.create_string_buffer(bytes(px), len(px))
        self.ctx.check(self.lib.ws_bn128_fr_fft(self.ctx.h, ctypes.cast(buf, ctypes.c_void_p), len(px) // 32, odd))
        return buf.raw

    def ifft(self, px, odd=0):
        buf = ctypes.create_string_buffer(bytes(px), len(px))
        self.ctx.check(self.lib.ws_bn128_fr_ifft(self.ctx.h, ctypes.cast(buf, ctypes.c_void_p), len(px) // 32, odd))
        return buf.raw

    def calcH(self, signals, polsA, polsB, nSignals, domainSize):
        """src/bn128.js:569-578 -> worker CALC_H (:126-166): domainSize*32 bytes, non-Montgomery."""
        if len(signals) < nSignals * 32:
            raise ValueError("signals shorter than nSignals * 32 bytes")
        out = ctypes.create_string_buffer(domainSize * 32)
        self.ctx.check(self.lib.ws_bn128_calc_h(self.ctx.h, as_ptr(signals), as_ptr(polsA), len(polsA), as_ptr(polsB), len(polsB),
                                                nSignals, domainSize, ctypes.cast(out, ctypes.c_void_p)))
        return out.raw

    # -- proving key kept in HBM across proofs (the reference re-posts it to its workers every time)
    def load_proving_key(self, pkey):
        h = ctypes.c_void_p()
        self.ctx.check(self.lib.ws_bn128_pkey_load(self.ctx.h, as_ptr(pkey), len(pkey), ctypes.byref(h)))
        return ProvingKey(self, h)

    def groth16GenProof(self, signals, pkey, r=None, s=None, raw=False):
        """src/bn128.js:580-720.  `pkey` is proving_key.bin bytes or a ProvingKey from
        load_proving_key().  r, s: 32-byte strings injected for tests (None => CSPRNG like the
        reference; the values used are kept in self._pr / self._ps like src/bn128.js:663-664 when
        injected).  Returns {pi_a, pi_b, pi_c} as decimal strings (src/bn128.js:714-718)."""
        for nm, v in (("r", r), ("s", s)):
            if v is not None and len(v) != 32:
                raise ValueError("%s must be exactly 32 bytes" % nm)
        pa = ctypes.create_string_buffer(96)
        pb = ctypes.create_string_buffer(192)
        pc = ctypes.create_string_buffer(96)
        outs = [ctypes.cast(x, ctypes.c_void_p) for x in (pa, pb, pc)]
        self._pr, self._ps = r, s
        if isinstance(pkey, ProvingKey):
            self.ctx.check(self.lib.ws_bn128_groth16_prove_key(self.ctx.h, pkey.h, as_ptr(signals), len(signals), as_ptr(r), as_ptr(s), *outs))
        else:
            self.ctx.check(self.lib.ws_bn128_groth16_prove(self.ctx.h, as_ptr(signals), len(signals), as_ptr(pkey), len(pkey), as_ptr(r), as_ptr(s), *outs))
        if raw:
            return pa.raw, pb.raw, pc.raw
        return {"pi_a": self.bin2g1(pa.raw), "pi_b": self.bin2g2(pb.raw), "pi_c": self.bin2g1(pc.raw)}

    def groth16Verify(self, verificationKey, input, proof):
        """src/bn128.js:722-786 (host CPU; see module-level groth16Verify)."""
        return groth16Verify(verificationKey, input, proof)

    # -- field batches (WASM exports f1m_*/frm_*; batch forms fft_mulN etc.)
    def fp_batch(self, field, op, x, y=None):
        return _lib.fp_batch(self.ctx, field, op, x, y)

    def fq2_mul_batch(self, x, y):
        if len(x) % 64 or len(y) != len(x):
            raise ValueError("x and y must be the same whole number of 64-byte Fq2 elements")
        n = len(x) // 64
        out = ctypes.create_string_buffer(max(1, n * 64))
        self.ctx.check(self.lib.ws_bn128_fq2_mul_batch(self.ctx.h, as_ptr(x), as_ptr(y), ctypes.cast(out, ctypes.c_void_p), n))
        return out.raw[:n * 64]

    # -- bin2int / bin2g1 / bin2g2, src/bn128.js:319-351
    @staticmethod
    def bin2int(b):
        return str(int.from_bytes(b[:32], "little"))

    @classmethod
    def bin2g1(cls, b):
        return [cls.bin2int(b[0:32]), cls.bin2int(b[32:64]), cls.bin2int(b[64:96])]

    @classmethod
    def bin2g2(cls, b):
        return [[cls.bin2int(b[0:32]), cls.bin2int(b[32:64])],
                [cls.bin2int(b[64:96]), cls.bin2int(b[96:128])],
                [cls.bin2int(b[128:160]), cls.bin2int(b[160:192])]]


class ProvingKey:
    """Handle on a proving key resident in HBM (ws_bn128_pkey_load)."""

    def __init__(self, bn, h):
        self.bn, self.h = bn, h
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        bn.lib.ws_bn128_pkey_info(h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        self.nVars, self.nPublic, self.domainSize = a.value, b.value, c.value

    def free(self):
        if self.h:
            self.bn.lib.ws_bn128_pkey_free(self.h)
            self.h = None


def groth16GenProof(witness, provingKey, cb=None, device=0):
    """main_bn128.js:27-41 `window.groth16GenProof(witness, provingKey, cb)` (README name
    genZKSnarkProof): cb(err, proof) if given, else returns the proof."""
    try:
        bn = build_bn128(device)
        try:
            proof = bn.groth16GenProof(witness, provingKey)
        finally:
            bn.terminate()
    except Exception as e:      # noqa: BLE001 -- the reference forwards any rejection to cb(err)
        if cb:
            return cb(e, None)
        raise
    if cb:
        return cb(None, proof)
    return proof


genZKSnarkProof = groth16GenProof


def proof_json(proof):
    """The snarkjs proof.json the reference's example ships (example/bn128/proof_good.json): the prover's object
    plus "protocol": "groth", as text."""
    import json
    return json.dumps({"pi_a": proof["pi_a"], "pi_b": proof["pi_b"], "pi_c": proof["pi_c"], "protocol": "groth"}, indent=1)


def _le32(v):
    return int(v).to_bytes(32, "little")


def _g1_bytes(e):
    """[x, y, ("1")] decimal strings -> 64 B plain LE (setG1Affine, src/bn128.js:503-507, without the Montgomery step)."""
    return _le32(e[0]) + _le32(e[1])


def _g2_bytes(e):
    """[[x0, x1], [y0, y1], (["1","0"])] -> 128 B (setG2Affine, src/bn128.js:509-513)."""
    return _le32(e[0][0]) + _le32(e[0][1]) + _le32(e[1][0]) + _le32(e[1][1])


def groth16Verify(verificationKey, input, proof):
    """Bn128.groth16Verify(verificationKey, input, proof), src/bn128.js:722-786: the objects are the parsed
    verification_key.json / public.json / proof.json (decimal strings or ints).  Runs on the host
    (ws_bn128_groth16_verify: four pairings are not GPU work); True / False like the reference.
    Values that do not fit 256 bits cannot be encoded and give False."""
    if input is None:
        input = []
    elif not isinstance(input, (list, tuple)):
        input = [input]                                          # src/bn128.js:724-728
    lib = _lib.load_library()
    try:
        ic = b"".join(_g1_bytes(e) for e in verificationKey["IC"])
        ins = b"".join(_le32(v) for v in input)
        args = [_g1_bytes(verificationKey["vk_alfa_1"]), _g2_bytes(verificationKey["vk_beta_2"]), _g2_bytes(verificationKey["vk_gamma_2"]),
                _g2_bytes(verificationKey["vk_delta_2"])]
        pa, pb, pc = _g1_bytes(proof["pi_a"]), _g2_bytes(proof["pi_b"]), _g1_bytes(proof["pi_c"])
    except OverflowError:
        return False
    if len(verificationKey["IC"]) != len(input) + 1:
        raise ValueError("verificationKey.IC must hold nPublic + 1 points")
    valid = ctypes.c_int(0)
    rc = lib.ws_bn128_groth16_verify(*[as_ptr(a) for a in args], as_ptr(ic), len(verificationKey["IC"]), as_ptr(ins) if input else None,
                                     len(input), as_ptr(pa), as_ptr(pb), as_ptr(pc), ctypes.byref(valid))
    if rc != _lib.WS_OK:
        raise _lib.WsError(rc, "ws_bn128_groth16_verify: bad arguments")
    return valid.value == 1


def pairing(p, q):
    """bn128_pairing (src/bn128/build_bn128.js:1320-1346) on plain-form points: p = [x, y], q = [[x0, x1], [y0, y1]];
    returns the 12 Fq coefficients (ints) in the reference's ftm order."""
    out = ctypes.create_string_buffer(384)
    rc = _lib.load_library().ws_bn128_pairing(as_ptr(_g1_bytes(p)), as_ptr(_g2_bytes(q)), ctypes.cast(out, ctypes.c_void_p))
    if rc != _lib.WS_OK:
        raise _lib.WsError(rc, "ws_bn128_pairing: point not on its curve")
    return [int.from_bytes(out.raw[32 * i:32 * i + 32], "little") for i in range(12)]


class MultiBn128:
    """All the GPUs of one box behind one object (ws_multi_*): the reference's `Bn128` with its worker pool mapped
    onto the GPUs (src/bn128.js:173-265, :353-415, :580-720).  One process, blocking calls, host buffers."""
    bin2int, bin2g1, bin2g2 = Bn128.bin2int, Bn128.bin2g1, Bn128.bin2g2

    def __init__(self, ngpus=0):
        self.lib = _lib.load_library()
        h = ctypes.c_void_p()
        rc = self.lib.ws_multi_create(ngpus, ctypes.byref(h))
        if rc != _lib.WS_OK:
            raise _lib.WsError(rc, "ws_multi_create failed (asked for %d GPUs)" % ngpus)
        self.h = h
        self.ngpus = self.lib.ws_multi_gpus(h)

    def terminate(self):
        if self.h:
            self.lib.ws_multi_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc != _lib.WS_OK:
            raise _lib.WsError(rc, (self.lib.ws_multi_last_error(self.h) or b"").decode())

    def _multiexp(self, fn, scalars, points, psize, osize):
        if len(scalars) % 32:
            raise ValueError("scalars must be n*32 bytes")
        n = len(scalars) // 32
        if len(points) < n * psize:
            raise ValueError("points buffer too short for %d scalars" % n)
        out = ctypes.create_string_buffer(osize)
        self._check(fn(self.h, as_ptr(scalars), as_ptr(points), n, ctypes.cast(out, ctypes.c_void_p)))
        return out.raw

    def g1_multiexp(self, scalars, points):
        return self._multiexp(self.lib.ws_multi_bn128_g1_multiexp, scalars, points, 64, 96)

    def g2_multiexp(self, scalars, points):
        return self._multiexp(self.lib.ws_multi_bn128_g2_multiexp, scalars, points, 128, 192)

    RESIDENT = object()       # pass as `pkey` to prove with the key made resident by load_proving_key

    def load_proving_key(self, pkey):
        """Makes `pkey` the handle's resident key (its shards stay on the GPUs); returns the token to pass as `pkey`."""
        self._check(self.lib.ws_multi_bn128_pkey_load(self.h, as_ptr(pkey), len(pkey)))
        return self.RESIDENT

    def groth16GenProof(self, signals, pkey, r=None, s=None, raw=False):
        """src/bn128.js:580-720 over all the GPUs of the handle; pkey: proving_key.bin bytes, or the token returned by
        load_proving_key()."""
        for nm, v in (("r", r), ("s", s)):
            if v is not None and len(v) != 32:
                raise ValueError("%s must be exactly 32 bytes" % nm)
        pa, pb, pc = ctypes.create_string_buffer(96), ctypes.create_string_buffer(192), ctypes.create_string_buffer(96)
        outs = [ctypes.cast(x, ctypes.c_void_p) for x in (pa, pb, pc)]
        if pkey is self.RESIDENT:
            self._check(self.lib.ws_multi_bn128_groth16_prove_resident(self.h, as_ptr(signals), len(signals), as_ptr(r), as_ptr(s), *outs))
        else:
            self._check(self.lib.ws_multi_bn128_groth16_prove(self.h, as_ptr(signals), len(signals), as_ptr(pkey), len(pkey), as_ptr(r), as_ptr(s), *outs))
        if raw:
            return pa.raw, pb.raw, pc.raw
        return {"pi_a": self.bin2g1(pa.raw), "pi_b": self.bin2g2(pb.raw), "pi_c": self.bin2g1(pc.raw)}


def build_bn128_multi(ngpus=0):
    """buildBn128() with the worker pool = the GPUs of the box (ngpus <= 0: all visible devices)."""
    return MultiBn128(ngpus)


def build_bn128(device=0):
    """index.js:20-22 / src/bn128.js:173-265 `buildBn128()`."""
    return Bn128(device)
