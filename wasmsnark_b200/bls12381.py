"""BLS12-381 host mirror over the CUDA library: the operations the reference's bls12381 module exports on the hot
path (/root/reference/src/bls12381/build_bls12381.js:33-43: f1m / frm, g1m / g2m + multiexp, fft, f2m), with the
calling conventions of the BN128 mirror (wasmsnark_b200/bn128.py).  The reference has no JS host class for this
curve; method names follow its BN128 host (src/bn128.js:353-439)."""
import ctypes

from . import _lib
from ._lib import Context, as_ptr


class Bls12381:
    q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    n8q, n8r = 48, 32

    def __init__(self, device=0):
        self.ctx = Context(device)
        self.lib = self.ctx.lib

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
        """scalars n*32 B raw LE, points n*96 B affine Montgomery -> 144 B Jacobian X|Y|Z."""
        return self._multiexp(self.lib.ws_bls12381_g1_multiexp, scalars, points, 96, 144)

    def g2_multiexp(self, scalars, points):
        """points n*192 B -> 288 B."""
        return self._multiexp(self.lib.ws_bls12381_g2_multiexp, scalars, points, 192, 288)

    def _affine(self, fn, p, size, from_montgomery):
        n = len(p) // size
        out = ctypes.create_string_buffer(max(1, n * size))
        self.ctx.check(fn(self.ctx.h, as_ptr(p), ctypes.cast(out, ctypes.c_void_p), n, from_montgomery))
        return out.raw[:n * size]

    def g1_affine(self, p, from_montgomery=0):
        return self._affine(self.lib.ws_bls12381_g1_affine, p, 144, from_montgomery)

    def g2_affine(self, p, from_montgomery=0):
        return self._affine(self.lib.ws_bls12381_g2_affine, p, 288, from_montgomery)

    def fft(self, px, odd=0):
        buf = ctypes.create_string_buffer(bytes(px), len(px))
        self.ctx.check(self.lib.ws_bls12381_fr_fft(self.ctx.h, ctypes.cast(buf, ctypes.c_void_p), len(px) // 32, odd))
        return buf.raw

    def ifft(self, px, odd=0):
        buf = ctypes.create_string_buffer(bytes(px), len(px))
        self.ctx.check(self.lib.ws_bls12381_fr_ifft(self.ctx.h, ctypes.cast(buf, ctypes.c_void_p), len(px) // 32, odd))
        return buf.raw

    def fp_batch(self, field, op, x, y=None):
        return _lib.fp_batch(self.ctx, field, op, x, y)

    def fq2_mul_batch(self, x, y):
        if len(x) % 96 or len(y) != len(x):
            raise ValueError("x and y must be the same whole number of 96-byte Fq2 elements")
        n = len(x) // 96
        out = ctypes.create_string_buffer(max(1, n * 96))
        self.ctx.check(self.lib.ws_bls12381_fq2_mul_batch(self.ctx.h, as_ptr(x), as_ptr(y), ctypes.cast(out, ctypes.c_void_p), n))
        return out.raw[:n * 96]


def build_bls12381(device=0):
    return Bls12381(device)
