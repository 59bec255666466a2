"""node/wasmsnark_b200_addon.c LINKED AND EXECUTED against a mock N-API runtime (tests/host/napi_mock) -- SURVEY.md §8(f) rank 1.

The build image has neither Node.js nor its headers, so the real addon build (node/binding.gyp) and the reference's mocha
suite cannot run here.  What can: the addon's C code, compiled unchanged against a header with the documented N-API
signatures and linked with a toy runtime that implements those functions.  This file plays node/index.js: it calls the
exports with ArrayBuffers / typed arrays / Buffers, a callback for prove(), and drains the async-work queue the way the
event loop would (execute on worker threads, complete on the main thread).

CPU tests: module registration, verify() on the reference's example proofs (host-code verifier: the whole path runs
without a GPU), argument validation, exceptions, handle life cycle without a device.
GPU tests (marked): create / prove / multiexp / affine / calcH / destroy against the reference's golden outputs, two
overlapping proofs on one handle (the reference's promise API allows that: src/bn128.js:282-300), terminate() semantics.
"""
import ctypes
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
MOCK = os.path.join(ROOT, "tests", "host", "napi_mock")
SO = os.path.join(ROOT, "tests", "host", "libaddon_mock.so")

K_UNDEFINED, K_NULL, K_BOOL, K_NUMBER, K_STRING, K_OBJECT, K_FUNCTION, K_EXTERNAL, K_ARRAYBUFFER, K_TYPEDARRAY, K_BUFFER, K_ERROR = range(12)
NAPI_UINT8_ARRAY, NAPI_UINT32_ARRAY = 1, 6
HOST_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p)
EXPORTS = ("create", "destroy", "prove", "g1Multiexp", "g2Multiexp", "g1Affine", "g2Affine", "calcH", "verify")


def _build():
    from wasmsnark_b200 import _lib
    _lib.load_library()                     # raises with the build hint when the product library is missing
    srcs = [os.path.join(ROOT, "node", "wasmsnark_b200_addon.c"), os.path.join(MOCK, "napi_mock.c")]
    deps = srcs + [os.path.join(MOCK, "node_api.h"), os.path.join(ROOT, "include", "wasmsnark_b200.h")]
    so, link = SO, ["-L" + os.path.join(ROOT, "wasmsnark_b200"), "-lwasmsnark_b200", "-Wl,-rpath,$ORIGIN/../../wasmsnark_b200"]
    if os.environ.get("WS_TEST_EMULATED"):  # tests/test_emulated_library.py: the same addon over the CPU-emulated kernels
        so, link = SO.replace(".so", "_emu.so"), ["-L" + os.path.join(ROOT, "tests", "host"), "-lwasmsnark_b200_emu", "-Wl,-rpath,$ORIGIN"]
        deps.append(os.path.join(ROOT, "tests", "host", "libwasmsnark_b200_emu.so"))
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["gcc", "-std=gnu11", "-O1", "-g", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-I" + MOCK, "-I" + os.path.join(ROOT, "include")] + srcs + link + ["-lpthread", "-o", so])
    return ctypes.CDLL(so)


class Js:
    """The JavaScript side, as far as the addon can tell."""

    def __init__(self, lib):
        self.lib = lib
        vp, sz = ctypes.c_void_p, ctypes.c_size_t
        for name, res, args in (("mock_env_new", vp, []), ("mock_env_free", ctypes.c_int, [vp]), ("mock_undefined", vp, [vp]), ("mock_null", vp, [vp]),
                                ("mock_object", vp, [vp]), ("mock_number", vp, [vp, ctypes.c_double]), ("mock_string", vp, [vp, ctypes.c_char_p]),
                                ("mock_arraybuffer", vp, [vp, ctypes.c_char_p, sz]), ("mock_typedarray", vp, [vp, ctypes.c_int, vp, sz, sz]),
                                ("mock_buffer", vp, [vp, ctypes.c_char_p, sz]), ("mock_function", vp, [vp, HOST_FN, vp]), ("mock_kind", ctypes.c_int, [vp]),
                                ("mock_num", ctypes.c_double, [vp]), ("mock_str", ctypes.c_char_p, [vp]), ("mock_code", ctypes.c_char_p, [vp]),
                                ("mock_is_type_error", ctypes.c_int, [vp]), ("mock_data", vp, [vp]), ("mock_len", sz, [vp]), ("mock_get", vp, [vp, ctypes.c_char_p]),
                                ("mock_call", vp, [vp, vp, sz, ctypes.POINTER(vp)]), ("mock_pending", vp, [vp]), ("mock_clear_pending", None, [vp]),
                                ("mock_live_refs", ctypes.c_int, [vp]), ("mock_queued", ctypes.c_int, [vp]), ("mock_run_async", ctypes.c_int, [vp]),
                                ("napi_register_module_v1", vp, [vp, vp])):
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        self.env = lib.mock_env_new()
        self.exports = lib.napi_register_module_v1(self.env, lib.mock_object(self.env))
        self._keep = []

    def close(self):
        """environment teardown; returns what the addon leaked (references + async works)"""
        leaked = self.lib.mock_env_free(self.env)
        self.env = None
        return leaked

    # values
    def ab(self, data):
        return self.lib.mock_arraybuffer(self.env, bytes(data), len(data))

    def u8(self, data, offset=0):
        """Uint8Array view at `offset` into a larger ArrayBuffer"""
        whole = self.ab(bytes(offset) + bytes(data))
        return self.lib.mock_typedarray(self.env, NAPI_UINT8_ARRAY, whole, offset, len(data))

    def u32(self, data):
        return self.lib.mock_typedarray(self.env, NAPI_UINT32_ARRAY, self.ab(data), 0, len(data) // 4)

    def buf(self, data):
        return self.lib.mock_buffer(self.env, bytes(data), len(data))

    def num(self, x):
        return self.lib.mock_number(self.env, float(x))

    @property
    def null(self):
        return self.lib.mock_null(self.env)

    @property
    def undefined(self):
        return self.lib.mock_undefined(self.env)

    def fn(self, pyfn):
        """a JS function: pyfn(list of value handles) -> None"""
        def tramp(env, argc, argv, data):
            pyfn([argv[i] for i in range(argc)])
            return None
        c = HOST_FN(tramp)
        self._keep.append(c)
        return self.lib.mock_function(self.env, c, None)

    # inspection
    def bytes_of(self, v):
        assert self.lib.mock_kind(v) == K_ARRAYBUFFER, self.lib.mock_kind(v)
        return ctypes.string_at(self.lib.mock_data(v), self.lib.mock_len(v))

    def get(self, obj, name):
        return self.lib.mock_get(obj, name.encode())

    def call(self, name, *args):
        """exports[name](...args): the returned value handle, or raises JsError for a thrown exception"""
        argv = (ctypes.c_void_p * max(1, len(args)))(*args)
        r = self.lib.mock_call(self.env, self.get(self.exports, name), len(args), argv)
        e = self.lib.mock_pending(self.env)
        if e:
            err = JsError(self.lib.mock_str(e).decode(), bool(self.lib.mock_is_type_error(e)), (self.lib.mock_code(e) or b"").decode())
            self.lib.mock_clear_pending(self.env)
            raise err
        assert r, "call returned no value and threw nothing"
        return r

    def prove(self, handle, witness, pkey, r=None, s=None):
        """queues prove(); returns the dict the callback fills in ({'err': str} or {'pi_a', 'pi_b', 'pi_c'}) once run_async() ran"""
        out = {}

        def cb(args):
            assert len(args) == 2
            if self.lib.mock_kind(args[0]) == K_ERROR:
                out["err"] = self.lib.mock_str(args[0]).decode()
                assert self.lib.mock_kind(args[1]) == K_NULL
            else:
                assert self.lib.mock_kind(args[0]) == K_NULL and self.lib.mock_kind(args[1]) == K_OBJECT
                for k in ("pi_a", "pi_b", "pi_c"):
                    out[k] = self.bytes_of(self.get(args[1], k))
        res = self.call("prove", handle, witness, pkey, self.ab(r) if r is not None else self.null, self.ab(s) if s is not None else self.null, self.fn(cb))
        assert self.lib.mock_kind(res) == K_UNDEFINED
        return out

    def run_async(self):
        return self.lib.mock_run_async(self.env)


class JsError(Exception):
    def __init__(self, msg, type_error, code):
        super().__init__(msg)
        self.type_error, self.code = type_error, code


@pytest.fixture(scope="module")
def addon():
    return _build()


@pytest.fixture()
def js(addon):
    j = Js(addon)
    yield j
    if j.env:
        assert j.close() == 0, "the addon left references or async works behind"


def le32(x):
    return int(x).to_bytes(32, "little")


def g1b(e):
    return le32(e[0]) + le32(e[1])


def g2b(e):
    return le32(e[0][0]) + le32(e[0][1]) + le32(e[1][0]) + le32(e[1][1])


def pack_verify(vk, public, proof):
    """what node/index.js: groth16Verify hands to native.verify"""
    return (b"".join([g1b(vk["vk_alfa_1"]), g2b(vk["vk_beta_2"]), g2b(vk["vk_gamma_2"]), g2b(vk["vk_delta_2"])]), b"".join(g1b(p) for p in vk["IC"]),
            b"".join(le32(x) for x in public), g1b(proof["pi_a"]) + g2b(proof["pi_b"]) + g1b(proof["pi_c"]))


def ints(b, n):
    return [str(int.from_bytes(b[32 * i:32 * i + 32], "little")) for i in range(n)]


def proof_json(p):
    """node/index.js: bin2g1 / bin2g2 on the three buffers (src/bn128.js:329-351, :714-718)"""
    a, b, c = ints(p["pi_a"], 3), ints(p["pi_b"], 6), ints(p["pi_c"], 3)
    return {"pi_a": a, "pi_b": [b[0:2], b[2:4], b[4:6]], "pi_c": c}


# ------------------------------------------------------------------------------------------------ CPU
def test_js_wrapper_constants_and_reference_method_names():
    """node/index.js cannot be executed here (no Node.js); what can be checked from its text: the two BigInt constants of its
    g?_fromMontgomery (q, and R^-1 mod q with R = 2^256) and that the class offers the reference's proving-path methods under
    the reference's names (src/bn128.js:353-439, 562-786)."""
    import re
    src = open(os.path.join(ROOT, "node", "index.js")).read()
    q = int(re.search(r"BN128_Q = (\d+)n", src).group(1))
    rinv = int(re.search(r"BN128_RINV = (\d+)n", src).group(1))
    assert q == 21888242871839275222246405745257275088696311157297823662689037894645226208583
    assert (rinv << 256) % q == 1
    for name in ("groth16GenProof", "groth16Verify", "g1_multiexp", "g2_multiexp", "g1_affine", "g2_affine", "g1_fromMontgomery", "g2_fromMontgomery",
                 "calcH", "bin2g1", "bin2g2", "terminate"):
        assert re.search(r"^\s+(async )?%s\(" % name, src, re.M), name


def test_module_registers_the_documented_exports(js):
    for name in EXPORTS:
        assert js.lib.mock_kind(js.get(js.exports, name)) == K_FUNCTION, name


def test_verify_runs_the_reference_cases_through_the_addon(js):
    """native.verify on the verdict table the reference's own groth16Verify produced (tests/golden/example_bn128.verify.json):
    every buffer flavour index.js may pass (ArrayBuffer, Uint8Array view at an offset, Node Buffer)."""
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    vk = g["verification_key"]
    flavours = (js.ab, lambda d: js.u8(d, 40), js.buf)
    seen = set()
    for i, case in enumerate(g["cases"]):
        parts = pack_verify(case["vk_patch"] or vk, case["public"], case["proof"])
        mk = flavours[i % len(flavours)]
        res = js.call("verify", *[mk(p) for p in parts])
        assert js.lib.mock_kind(res) == K_BOOL
        assert bool(js.lib.mock_num(res)) == case["valid"], case["name"]
        seen.add(case["valid"])
    assert seen == {True, False}


def test_verify_rejects_malformed_arguments(js):
    g = json.load(open(os.path.join(GOLD, "example_bn128.verify.json")))
    vk, case = g["verification_key"], g["cases"][0]
    vkb, ic, ins, pr = pack_verify(vk, case["public"], case["proof"])
    for args in ((vkb[:-1], ic, ins, pr), (vkb, ic[:-64], ins, pr), (vkb, ic, ins + bytes(32), pr), (vkb, ic, ins, pr + b"\0"), (vkb, ic + b"\0", ins, pr)):
        with pytest.raises(JsError) as e:
            js.call("verify", *[js.ab(a) for a in args])
        assert e.value.type_error
    with pytest.raises(JsError) as e:
        js.call("verify", js.ab(vkb), js.num(3), js.ab(ins), js.ab(pr))
    assert e.value.type_error
    with pytest.raises(JsError):
        js.call("verify", js.ab(vkb))               # missing arguments read as undefined
    # a Uint32Array is measured in bytes, not elements
    assert bool(js.lib.mock_num(js.call("verify", js.u32(vkb), js.u32(ic), js.u32(ins), js.u32(pr)))) is True


def test_calls_on_something_that_is_not_a_handle_throw(js):
    junk = js.ab(bytes(32))
    for name, args in (("g1Multiexp", (junk, junk, junk)), ("g2Multiexp", (js.null, junk, junk)), ("g1Affine", (js.num(1), junk)), ("g2Affine", (junk, junk)),
                       ("calcH", (junk, junk, junk, junk, js.num(1), js.num(2))), ("destroy", (junk,)), ("destroy", ()),
                       ("prove", (junk, junk, junk, js.null, js.null, js.fn(lambda a: None))), ("prove", (junk, junk))):
        with pytest.raises(JsError) as e:
            js.call(name, *args)
        assert e.value.type_error, name
    assert js.lib.mock_queued(js.env) == 0 and js.lib.mock_live_refs(js.env) == 0


def _have_gpu():
    from wasmsnark_b200 import _lib
    lib = _lib.load_library()
    h = ctypes.c_void_p()
    if lib.ws_create(0, ctypes.byref(h)) != 0:
        return False
    lib.ws_destroy(h)
    return True


def test_create_without_a_device_throws_and_there_is_no_cpu_fallback(js):
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    with pytest.raises(JsError) as e:
        js.call("create", js.num(0), js.num(1))
    assert "no CPU fallback" in str(e.value) and e.value.code == "WASMSNARK_B200"


# ------------------------------------------------------------------------------------------------ GPU
def _example():
    g = json.load(open(os.path.join(GOLD, "synth_small.golden.json")))
    pk = open(os.path.join(GOLD, "synth_small.pkey.bin"), "rb").read()
    wit = open(os.path.join(GOLD, "synth_small.witness.bin"), "rb").read()
    return g, pk, wit


@pytest.mark.gpu
def test_prove_through_the_addon_equals_the_reference_proof(js):
    """create -> prove (async work) -> callback(null, {pi_a, pi_b, pi_c}); formatted as index.js formats it, the proof equals
    the one the reference's own code produced for the same key, witness and (r, s) (tests/golden/synth_small.golden.json)."""
    g, pk, wit = _example()
    h = js.call("create", js.num(0), js.num(1))
    assert js.lib.mock_kind(h) == K_EXTERNAL
    out = js.prove(h, js.ab(wit), js.u8(pk, 8), bytes.fromhex(g["r"]), bytes.fromhex(g["s"]))
    assert out == {} and js.lib.mock_queued(js.env) == 1           # nothing happens before the event loop turns
    assert js.run_async() == 1
    assert "err" not in out, out
    assert proof_json(out) == g["proof_rs"]
    # random (r, s): a different proof each time
    a, b = js.prove(h, js.ab(wit), js.ab(pk)), js.prove(h, js.buf(wit), js.buf(pk))
    assert js.run_async() == 2
    assert "err" not in a and "err" not in b and a["pi_a"] != b["pi_a"]
    js.call("destroy", h)


@pytest.mark.gpu
def test_overlapping_proofs_on_one_handle(js):
    """Four groth16GenProof promises pending at once on one Bn128 (the reference queues them, src/bn128.js:282-300): the
    blocking C calls run on four worker threads against ONE context and every callback gets the right proof."""
    g, pk, wit = _example()
    h = js.call("create", js.num(0), js.num(1))
    r, s = bytes.fromhex(g["r"]), bytes.fromhex(g["s"])
    outs = [js.prove(h, js.ab(wit), js.ab(pk), r, s) for _ in range(3)] + [js.prove(h, js.ab(wit[:64]), js.ab(pk), r, s)]
    assert js.run_async() == 4
    for o in outs[:3]:
        assert proof_json(o) == g["proof_rs"]
    assert "err" in outs[3] and "witness" in outs[3]["err"]            # a short witness fails that call only
    js.call("destroy", h)


@pytest.mark.gpu
def test_terminate_semantics(js):
    """destroy(handle) frees the GPU state at once -- or, with proofs in flight, when the last one completes; the handle is
    dead afterwards."""
    g, pk, wit = _example()
    h = js.call("create", js.num(0), js.num(1))
    out = js.prove(h, js.ab(wit), js.ab(pk), bytes.fromhex(g["r"]), bytes.fromhex(g["s"]))
    js.call("destroy", h)                                       # terminate() while the proof is still queued
    with pytest.raises(JsError):
        js.prove(h, js.ab(wit), js.ab(pk))
    assert js.run_async() == 1 and proof_json(out) == g["proof_rs"]
    for name, args in (("g1Multiexp", (h, js.ab(bytes(32)), js.ab(bytes(64)))), ("g1Affine", (h, js.ab(bytes(96)))), ("prove", (h, js.ab(wit), js.ab(pk), js.null, js.null, js.fn(lambda a: None)))):
        with pytest.raises(JsError) as e:
            js.call(name, *args)
        assert e.value.type_error
    js.call("destroy", h)                                       # twice is harmless


@pytest.mark.gpu
def test_multiexp_affine_calch_through_the_addon(js):
    """the synchronous exports against the oracle / the reference's calcH"""
    import random
    import util
    from oracle import bn128 as O
    h = js.call("create", js.num(0), js.num(1))
    rnd = random.Random(7)
    n = 200
    aff, P0, D = util.arith_points(O.Fq, O.G1_GEN, n, 3)
    sc = [rnd.getrandbits(256) for _ in range(n)]
    jac = js.bytes_of(js.call("g1Multiexp", h, js.ab(b"".join(util.b32(x) for x in sc)), js.u8(util.g1_points_bytes(aff), 64)))
    assert len(jac) == 96
    got = js.bytes_of(js.call("g1Affine", h, js.ab(jac)))
    assert util.g1_result_tuple(got) == util.closed_form_sum(sc, P0, D)
    from test_host_math import gen_points
    pts2 = gen_points(O.Fq2, O.G2_GEN, rnd, 6)
    sc2 = [rnd.getrandbits(256) for _ in range(6)]
    raw2 = b"".join(util.g2_affine_m(P) for P in pts2)
    got2 = js.bytes_of(js.call("g2Affine", h, js.call("g2Multiexp", h, js.ab(b"".join(util.b32(x) for x in sc2)), js.ab(raw2))))
    assert util.g2_result_tuple(got2) == O.pt_affine(O.Fq2, O.multiexp_naive(O.Fq2, sc2, pts2))
    with pytest.raises(JsError) as e:                               # fewer points than scalars
        js.call("g1Multiexp", h, js.ab(bytes(64)), js.ab(bytes(64)))
    assert e.value.type_error
    # calcH on the small synthetic key vs the h the reference computed (golden)
    g, pk, wit = _example()
    import struct
    hdr = struct.unpack("<10I", pk[:40])
    n_vars, domain, p_a, p_b, p_pts = hdr[0], hdr[2], hdr[3], hdr[4], hdr[5]
    hh = js.bytes_of(js.call("calcH", h, js.ab(wit), js.ab(pk[p_a:p_b]), js.ab(pk[p_b:p_pts]), js.num(n_vars), js.num(domain)))
    assert len(hh) == domain * 32
    assert hh.hex() == g["calcH"]
    js.call("destroy", h)
