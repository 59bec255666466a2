"""ctypes binding of libwasmsnark_b200.so (C ABI: include/wasmsnark_b200.h).

Fails loudly: if the shared library has not been built, or no CUDA device can be opened,
an exception is raised -- there is no fallback implementation.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwasmsnark_b200.so")

WS_OK = 0
FIELD_BN128_FQ, FIELD_BN128_FR, FIELD_MNT6753_FQ, FIELD_BLS12381_FQ, FIELD_BLS12381_FR = 0, 1, 2, 3, 4
OP_MUL, OP_ADD, OP_SUB, OP_SQUARE, OP_TOMONT, OP_FROMMONT, OP_INVERSE = range(7)
# bytes per element of each field id (ws_fp_batch reads / writes n * size bytes)
FIELD_ELEMENT_BYTES = {FIELD_BN128_FQ: 32, FIELD_BN128_FR: 32, FIELD_MNT6753_FQ: 96, FIELD_BLS12381_FQ: 48, FIELD_BLS12381_FR: 32}
_BINARY_OPS = (OP_MUL, OP_ADD, OP_SUB)


def fp_batch(ctx, field, op, x, y=None):
    """ws_fp_batch with the buffer lengths checked against the field's element size (shared by the curve mirrors)."""
    esz = FIELD_ELEMENT_BYTES.get(field)
    if esz is None:
        raise ValueError("unknown field id %r" % (field,))
    if len(x) % esz:
        raise ValueError("x must be a whole number of %d-byte elements" % esz)
    if op in _BINARY_OPS:
        if y is None or len(y) != len(x):
            raise ValueError("y must have the same length as x for a binary op")
    elif y is not None and len(y) != len(x):
        raise ValueError("y, if given, must have the same length as x")
    n = len(x) // esz
    out = ctypes.create_string_buffer(max(1, n * esz))
    ctx.check(ctx.lib.ws_fp_batch(ctx.h, field, op, as_ptr(x), as_ptr(y), ctypes.cast(out, ctypes.c_void_p), n))
    return out.raw[:n * esz]


class LibraryNotBuilt(RuntimeError):
    pass


class WsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("wasmsnark_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None
_vp, _sz, _i, _u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p

# name -> (restype, argtypes); kept in step with include/wasmsnark_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "ws_version": (ctypes.c_char_p, []),
    "ws_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "ws_destroy": (None, [_vp]),
    "ws_last_error": (ctypes.c_char_p, [_vp]),
    "ws_kernel_launches": (ctypes.c_ulonglong, [_vp]),
    "ws_set_profiling": (_i, [_vp, _i]),
    "ws_last_timings": (_i, [_vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float), _i]),
    "ws_set_msm_task_cap": (_i, [_vp, ctypes.c_uint]),
    "ws_set_msm_seg_len": (_i, [_vp, ctypes.c_uint]),
    "ws_set_msm_acc_variant": (_i, [_vp, _i]),
    "ws_fp_batch": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    "ws_bn128_fq2_mul_batch": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "ws_bn128_g1_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_bn128_g2_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_bn128_g1_multiexp_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "ws_bn128_g2_multiexp_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "ws_bn128_g1_affine": (_i, [_vp, _vp, _vp, _sz, _i]),
    "ws_bn128_g2_affine": (_i, [_vp, _vp, _vp, _sz, _i]),
    "ws_mnt6753_fq3_mul_batch": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "ws_mnt6753_g1_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_mnt6753_g1_multiexp_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "ws_mnt6753_g1_affine": (_i, [_vp, _vp, _vp, _sz, _i]),
    "ws_bn128_fr_fft": (_i, [_vp, _vp, _sz, _i]),
    "ws_bn128_fr_ifft": (_i, [_vp, _vp, _sz, _i]),
    "ws_bn128_fr_fft_dev": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "ws_bn128_calc_h": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, ctypes.c_uint32, ctypes.c_uint32, _vp]),
    "ws_bn128_groth16_prove": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "ws_set_key_tables": (_i, [_vp, _i]),
    "ws_set_key_cache": (_i, [_vp, _i]),
    "ws_set_h_priority": (_i, [_vp, _i]),
    "ws_bn128_groth16_verify": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp]),
    "ws_bn128_pairing": (_i, [_vp, _vp, _vp]),
    "ws_bls12381_fq2_mul_batch": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "ws_bls12381_g1_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_bls12381_g1_multiexp_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "ws_bls12381_g1_affine": (_i, [_vp, _vp, _vp, _sz, _i]),
    "ws_bls12381_g2_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_bls12381_g2_multiexp_dev": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp]),
    "ws_bls12381_g2_affine": (_i, [_vp, _vp, _vp, _sz, _i]),
    "ws_bls12381_fr_fft": (_i, [_vp, _vp, _sz, _i]),
    "ws_bls12381_fr_ifft": (_i, [_vp, _vp, _sz, _i]),
    "ws_bls12381_fr_fft_dev": (_i, [_vp, _vp, _sz, _i, _i, _vp]),
    "ws_bn128_pkey_load": (_i, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "ws_bn128_pkey_free": (None, [_vp]),
    "ws_bn128_pkey_info": (_i, [_vp, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]),
    "ws_bn128_groth16_prove_key": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "ws_bn128_groth16_prove_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "ws_bn128_pkey_load_shard": (_i, [_vp, _vp, _sz, _i, _i, ctypes.POINTER(_vp)]),
    "ws_bn128_groth16_partials_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "ws_bn128_groth16_prepare_dev": (_i, [_vp, _vp, _vp, _vp]),
    "ws_bn128_groth16_witness_msms_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "ws_bn128_groth16_h_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "ws_bn128_groth16_h_msm_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "ws_bn128_pkey_shard_info": (_i, [_vp, ctypes.POINTER(ctypes.c_uint32)]),
    "ws_bn128_pkey_load_ranges": (_i, [_vp, _vp, _sz, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(_vp)]),
    "ws_multi_set_split": (_i, [_vp, _i]),
    "ws_balanced_ranges": (_i, [_vp, _sz, _i, ctypes.POINTER(ctypes.c_uint32)]),
    "ws_set_ntt_variant": (_i, [_vp, _i]),
    "ws_multi_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "ws_multi_destroy": (None, [_vp]),
    "ws_multi_gpus": (_i, [_vp]),
    "ws_multi_last_error": (ctypes.c_char_p, [_vp]),
    "ws_multi_bn128_groth16_prove": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "ws_multi_bn128_pkey_load": (_i, [_vp, _vp, _sz]),
    "ws_multi_bn128_groth16_prove_resident": (_i, [_vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "ws_multi_bn128_g1_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_multi_bn128_g2_multiexp": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "ws_bn128_groth16_finish_dev": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "ws_msm_table_windows": (_i, [_i]),
    "ws_bn128_g1_build_table_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp]),
    "ws_bn128_g2_build_table_dev": (_i, [_vp, _vp, _sz, _i, _vp, _vp]),
    "ws_bn128_g1_multiexp_table_dev": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "ws_bn128_g2_multiexp_table_dev": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "ws_bn128_g1_sum_dev": (_i, [_vp, _vp, _i, _vp, _vp]),
    "ws_bn128_g2_sum_dev": (_i, [_vp, _vp, _i, _vp, _vp]),
}


def load_library(path=LIB_PATH):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise LibraryNotBuilt("%s not found: run `python -m wasmsnark_b200.build` (or __graft_entry__.build())" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_ptr(buf):
    """Host pointer of a bytes / bytearray / numpy array / ctypes buffer (no copy for writable ones)."""
    if buf is None:
        return None
    if isinstance(buf, bytes):
        return ctypes.cast(ctypes.c_char_p(buf), ctypes.c_void_p)
    if hasattr(buf, "ctypes"):                       # numpy
        return ctypes.c_void_p(buf.ctypes.data)
    if isinstance(buf, (bytearray, memoryview)):
        return ctypes.cast((ctypes.c_char * len(buf)).from_buffer(buf), ctypes.c_void_p)
    return ctypes.cast(buf, ctypes.c_void_p)


class Context:
    """One ws_ctx = one GPU (replaces the reference's worker pool, src/bn128.js:173-265)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.ws_create(device, ctypes.byref(h))
        if rc != WS_OK:
            raise WsError(rc, "ws_create failed (no usable CUDA device %d?)" % device)
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.ws_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != WS_OK:
            raise WsError(rc, (self.lib.ws_last_error(self.h) or b"").decode())

    def timings(self):
        names = (ctypes.c_char_p * 48)()
        ms = (ctypes.c_float * 48)()
        n = self.lib.ws_last_timings(self.h, names, ms, 48)
        if n < 0:
            self.check(n)
        return [(names[i].decode(), ms[i]) for i in range(n)]
