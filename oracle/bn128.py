"""oracle/bn128.py -- TEST INFRASTRUCTURE (oracle), not product code.

CPU restatement, with Python big integers, of the reference's Groth16 proving
path on BN128 (iden3/wasmsnark).  Every function cites the reference file:line
it follows.  Pinning: this restatement is checked (tests/test_oracle.py)
against

  * example/bn128/proof_good0.json -- the reference's own proof of
    example/bn128/{proving_key.bin, witness.bin} with r = s = 0
    (debug mode of src/bn128.js:666-668), all of pi_a, pi_b, pi_c;
  * the reference's own generated WASM image, translated to C and executed
    natively (oracle/wasm2c.py -> oracle/_ref/, see oracle/ref_wasm.py), on
    f1m_mul/frm_mul, g1m/g2m add/double/affine, g1m_multiexp2, g2m_multiexp,
    fft_fft/fft_ifft (odd = 0/1), pol_constructLC and calcH;
  * the committed fixtures under tests/golden/ generated from that image.

Representation: field elements are Python ints.  Functions whose name ends in
`_m` take/return Montgomery-form values (x*R mod p, R = 2^256,
src/build_f1m.js:37-39) exactly as they lie in proving_key.bin; internally
the group law runs on plain residues -- Montgomery form is a bijection, every
reference routine returns canonical values in [0,p) (src/build_f1m.js:426-435),
so converting at the boundary gives byte-identical results.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.
"""
import struct

# src/bn128/build_bn128.js:19-20
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_ORDER = 21888242871839275222246405745257275088548364400416034343698204186575808495617
RBITS = 256
RMONT = 1 << RBITS                     # Montgomery radix, src/build_f1m.js:37
MASK32 = 0xFFFFFFFF

G1_GEN = (1, 2)                        # src/bn128/build_bn128.js:51-55
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))   # :79-90


# --------------------------------------------------------------------------
# a2  Montgomery field (src/build_f1m.js)
# --------------------------------------------------------------------------
def to_mont(x, p=Q):
    """f1m_toMontgomery = mul by R^2 (src/build_f1m.js:749-756)."""
    return (x << RBITS) % p


_RINV = {Q: pow(RMONT, -1, Q), R_ORDER: pow(RMONT, -1, R_ORDER)}


def from_mont(x, p=Q):
    """f1m_fromMontgomery (src/build_f1m.js:758-770)."""
    return x * _RINV[p] % p


def mont_mul(x, y, p=Q):
    """f1m_mul / frm_mul: x*y*R^-1 mod p, canonical (src/build_f1m.js:235-436)."""
    return x * y * _RINV[p] % p


def f1m_mul_limbs(x, y, p=Q):
    """Literal restatement of the reference's product-scanning Montgomery
    multiplication on 32-bit limbs (src/build_f1m.js:235-436): column k gathers
    sum x_i*y_{k-i} + sum q_i*m_{k-i}; for k < n32 the quotient digit
    m_k = c0*np32 mod 2^32 is derived and q_0*m_k added; one final conditional
    subtraction (:426-435).  Used to check mont_mul() itself."""
    n32 = 8
    np32 = (-pow(p, -1, 1 << 32)) & MASK32              # :255
    xs = [(x >> (32 * i)) & MASK32 for i in range(n32)]
    ys = [(y >> (32 * i)) & MASK32 for i in range(n32)]
    qs = [(p >> (32 * i)) & MASK32 for i in range(n32)]
    m = [0] * n32
    out = [0] * n32
    acc = 0                                             # (c1:c0) as one integer
    for k in range(2 * n32 - 1):
        for i in range(max(0, k - n32 + 1), min(k, n32 - 1) + 1):
            acc += xs[i] * ys[k - i]
        for i in range(max(1, k - n32 + 1), min(k, n32 - 1) + 1):
            acc += qs[i] * m[k - i]
        if k < n32:
            m[k] = ((acc & MASK32) * np32) & MASK32
            acc += qs[0] * m[k]
        else:
            out[k - n32] = acc & MASK32
        acc >>= 32
    out[n32 - 1] = acc & MASK32
    carry = acc >> 32
    res = sum(v << (32 * i) for i, v in enumerate(out))
    if carry or res >= p:
        res = (res - p) & (RMONT - 1)
    return res


# --------------------------------------------------------------------------
# a1  int_mul (src/build_int.js:285-380): full 2n-limb product
# --------------------------------------------------------------------------
def int_mul(x, y):
    return x * y


# --------------------------------------------------------------------------
# a3  Fq2 = Fq[u]/(u^2+1)  (src/build_f2m.js; non-residue = -1 via f1m_neg,
#     src/bn128/build_bn128.js:40).  Elements are (c0, c1).
# --------------------------------------------------------------------------
class Fq:
    """Prime field ops with the same interface as Fq2 (plain residues)."""
    zero = 0
    one = 1

    @staticmethod
    def add(a, b):
        return (a + b) % Q

    @staticmethod
    def sub(a, b):
        return (a - b) % Q

    @staticmethod
    def neg(a):
        return (-a) % Q

    @staticmethod
    def mul(a, b):
        return a * b % Q

    @staticmethod
    def sqr(a):
        return a * a % Q

    @staticmethod
    def inv(a):
        return pow(a, -1, Q)

    @staticmethod
    def is_zero(a):
        return a == 0


class Fq2:
    zero = (0, 0)
    one = (1, 0)

    @staticmethod
    def add(a, b):
        return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)

    @staticmethod
    def sub(a, b):
        return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)

    @staticmethod
    def neg(a):
        return ((-a[0]) % Q, (-a[1]) % Q)

    @staticmethod
    def mul(a, b):
        """Karatsuba, 3 base muls (src/build_f2m.js:127-163)."""
        A = a[0] * b[0]
        B = a[1] * b[1]
        C = (a[0] + a[1]) * (b[0] + b[1])
        return ((A - B) % Q, (C - A - B) % Q)

    @staticmethod
    def sqr(a):
        """src/build_f2m.js:186-228."""
        return ((a[0] + a[1]) * (a[0] - a[1]) % Q, 2 * a[0] * a[1] % Q)

    @staticmethod
    def inv(a):
        """src/build_f2m.js:275-333: 1/(c0^2 + c1^2) * (c0, -c1)."""
        t = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
        return (a[0] * t % Q, (-a[1] * t) % Q)

    @staticmethod
    def is_zero(a):
        return a[0] == 0 and a[1] == 0


# --------------------------------------------------------------------------
# a4  Jacobian curve y^2 = x^3 + b, a = 0 (src/build_curve_jacobian_a0.js)
#     Points are (X, Y, Z) over field class F; infinity <=> Z == 0 (:33-47).
# --------------------------------------------------------------------------
def pt_zero(F):
    """g?m_zero = (0, 1, 0) (src/build_curve_jacobian_a0.js:87-113)."""
    return (F.zero, F.one, F.zero)


def pt_is_zero(F, p):
    return F.is_zero(p[2])


def pt_double(F, p):
    """dbl-2009-l (src/build_curve_jacobian_a0.js:174-235)."""
    if pt_is_zero(F, p):
        return p
    x, y, z = p
    A = F.sqr(x)
    B = F.sqr(y)
    C = F.sqr(B)
    D = F.sub(F.sub(F.sqr(F.add(x, B)), A), C)
    D = F.add(D, D)
    E = F.add(F.add(A, A), A)
    Fv = F.sqr(E)
    G = F.mul(y, z)
    x3 = F.sub(Fv, F.add(D, D))
    c8 = F.add(C, C)
    c8 = F.add(c8, c8)
    c8 = F.add(c8, c8)
    y3 = F.sub(F.mul(F.sub(D, x3), E), c8)
    z3 = F.add(G, G)
    return (x3, y3, z3)


def pt_add(F, p1, p2):
    """add-2007-bl with the reference's shortcuts (src/build_curve_jacobian_a0.js:280-385)."""
    if pt_is_zero(F, p1):
        return p2
    if pt_is_zero(F, p2):
        return p1
    x1, y1, z1 = p1
    x2, y2, z2 = p2
    Z1Z1 = F.sqr(z1)
    Z2Z2 = F.sqr(z2)
    U1 = F.mul(x1, Z2Z2)
    U2 = F.mul(x2, Z1Z1)
    S1 = F.mul(y1, F.mul(z2, Z2Z2))
    S2 = F.mul(y2, F.mul(z1, Z1Z1))
    if U1 == U2 and S1 == S2:
        return pt_double(F, p1)
    H = F.sub(U2, U1)
    rr = F.sub(S2, S1)
    rr = F.add(rr, rr)
    I = F.sqr(F.add(H, H))
    J = F.mul(H, I)
    V = F.mul(U1, I)
    x3 = F.sub(F.sub(F.sqr(rr), J), F.add(V, V))
    S1J = F.mul(S1, J)
    y3 = F.sub(F.mul(F.sub(V, x3), rr), F.add(S1J, S1J))
    z3 = F.mul(F.sub(F.sub(F.sqr(F.add(z1, z2)), Z1Z1), Z2Z2), H)
    return (x3, y3, z3)


def pt_neg(F, p):
    """src/build_curve_jacobian_a0.js:387-406."""
    return (p[0], F.neg(p[1]), p[2])


def pt_affine(F, p):
    """g?m_affine (src/build_curve_jacobian_a0.js:424-457): infinity -> (0,1,0)."""
    if pt_is_zero(F, p):
        return pt_zero(F)
    zi = F.inv(p[2])
    zi2 = F.sqr(zi)
    return (F.mul(p[0], zi2), F.mul(p[1], F.mul(zi, zi2)), F.one)


def pt_times_scalar(F, p, k):
    """g?m_timesScalar: MSB-first double-and-add over the scalar's bytes
    (src/build_timesscalar.js:20-77); k is a non-negative integer of any size."""
    res = pt_zero(F)
    for i in reversed(range(k.bit_length())):
        res = pt_double(F, res)
        if (k >> i) & 1:
            res = pt_add(F, res, p)
    return res


def load_affine(F, x, y):
    """How multiexp loads a key point (src/build_multiexp.js:335-349):
    Z = 1, or Z = 0 (infinity) if the stored x coordinate is zero."""
    return (x, y, F.zero if F.is_zero(x) else F.one)


# --------------------------------------------------------------------------
# a5/a6  multi-exponentiation  sum s_i * P_i
#   The reference (src/build_multiexp.js:498-744, w = 7) uses subset-sum tables
#   and per-bit accumulators; which bucketing is used cannot change the group
#   element, and results are compared after affine normalisation exactly as the
#   reference's own test does (test/bn128_prover.js:31-46).  This restatement
#   keeps the input conventions that matter: scalars are RAW 256-bit integers
#   (never reduced mod r), x == 0 means infinity.  The reference's literal
#   algorithm runs in oracle/_ref (ref_wasm.py) and is compared in the tests.
# --------------------------------------------------------------------------
def multiexp(F, scalars, points, c=None):
    """scalars: list of ints (< 2^256); points: list of Jacobian triples."""
    n = len(scalars)
    if n == 0:
        return pt_zero(F)
    if c is None:
        c = max(2, min(16, n.bit_length() - 2))
    nwin = (256 + c - 1) // c
    total = pt_zero(F)
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = pt_double(F, total)
        buckets = {}
        for s, p in zip(scalars, points):
            d = (s >> (w * c)) & ((1 << c) - 1)
            if d and not pt_is_zero(F, p):
                b = buckets.get(d)
                buckets[d] = p if b is None else pt_add(F, b, p)
        running = pt_zero(F)
        acc = pt_zero(F)
        prev = None
        for d in sorted(buckets, reverse=True):
            if prev is not None and prev - d > 1:
                acc = pt_add(F, acc, pt_times_scalar(F, running, prev - d - 1))
            running = pt_add(F, running, buckets[d])
            acc = pt_add(F, acc, running)
            prev = d
        if prev is not None and prev > 1:
            acc = pt_add(F, acc, pt_times_scalar(F, running, prev - 1))
        total = pt_add(F, total, acc)
    return total


def multiexp_naive(F, scalars, points):
    """Definitional sum: fold g?m_timesScalar with g?m_add."""
    acc = pt_zero(F)
    for s, p in zip(scalars, points):
        acc = pt_add(F, acc, pt_times_scalar(F, p, s))
    return acc


# --------------------------------------------------------------------------
# a7  NTT over Fr (src/build_fft.js)
# --------------------------------------------------------------------------
S_ADICITY = 28                                   # src/build_fft.js:29-34
NQR = 5                                          # :36


def roots():
    """w[s] = primitive 2^s-th root, w[28] = 5^((r-1)/2^28) (src/build_fft.js:36-47)."""
    w = [0] * (S_ADICITY + 1)
    w[S_ADICITY] = pow(NQR, (R_ORDER - 1) >> S_ADICITY, R_ORDER)
    for i in reversed(range(S_ADICITY)):
        w[i] = w[i + 1] * w[i + 1] % R_ORDER
    return w


_ROOTS = roots()


def _log2(n):
    """fft__log2 traps unless n is a power of two <= 2^28 (src/build_fft.js:137-154)."""
    if n <= 0 or n & (n - 1) or n > (1 << S_ADICITY):
        raise ValueError("fft size must be a power of two <= 2^28")
    return n.bit_length() - 1


def rawfft(x, odd):
    """In-place bit reversal + radix-2 DIT, twiddle W starts at 1 or, when
    `odd`, at w_{2m} (src/build_fft.js:223-372).  Plain residues mod r."""
    n = len(x)
    bits = _log2(n)
    p = R_ORDER
    a = [0] * n
    for i in range(n):
        a[int(format(i, "0%db" % bits)[::-1], 2) if bits else 0] = x[i]
    for s in range(1, bits + 1):
        m = 1 << s
        wm = _ROOTS[s]
        h = m >> 1
        w0 = _ROOTS[s + 1] if odd else 1
        tw = [0] * h
        W = w0
        for j in range(h):
            tw[j] = W
            W = W * wm % p
        for k in range(0, n, m):
            for j in range(h):
                t = tw[j] * a[k + j + h] % p
                u = a[k + j]
                a[k + j] = (u + t) % p
                a[k + j + h] = (u - t) % p
    return a


def fft(x, odd=0):
    """fft_fft (src/build_fft.js:159-187)."""
    return rawfft(x, odd)


def ifft(x, odd=0):
    """fft_ifft = rawfft + __finalInverse: res[i] = raw[(n-i) mod n] / n
    (src/build_fft.js:189-221, :550-648)."""
    n = len(x)
    a = rawfft(x, odd)
    ninv = pow(n, -1, R_ORDER)
    return [a[(n - i) % n] * ninv % R_ORDER for i in range(n)]


# --------------------------------------------------------------------------
# a8  pol_constructLC (src/build_pol.js:61-143) on the key's section layout
#     per signal: u32 ncoefs; ncoefs x { u32 idx; 32-byte Montgomery Fr coef }
# --------------------------------------------------------------------------
def construct_lc(pols, signals, n_signals, domain_size):
    """signals: plain ints (the reference first converts them with
    fft_toMontgomeryN, src/bn128.js:139; a Montgomery product of two Montgomery
    values is the Montgomery form of the plain product).  Returns plain residues."""
    res = [0] * domain_size
    pos = 0
    for i in range(n_signals):
        (ncoefs,) = struct.unpack_from("<I", pols, pos)
        pos += 4
        s = signals[i] % R_ORDER
        for _ in range(ncoefs):
            (idx,) = struct.unpack_from("<I", pols, pos)
            coef = from_mont(int.from_bytes(pols[pos + 4:pos + 36], "little"), R_ORDER)
            pos += 36
            res[idx] = (res[idx] + s * coef) % R_ORDER
    return res


def calc_h(signals, polsA, polsB, n_signals, domain_size):
    """Worker command CALC_H (src/bn128.js:126-166).  signals: plain ints.
    Returns the domain_size values h_j as plain ints (fft_fromMontgomeryN, :160)."""
    n = domain_size
    a = construct_lc(polsA, signals, n_signals, n)
    b = construct_lc(polsB, signals, n_signals, n)
    ac = ifft(a)
    bc = ifft(b)
    ao = fft(ac, 1)
    bo = fft(bc, 1)
    A2 = [0] * (2 * n)
    B2 = [0] * (2 * n)
    A2[0::2] = a
    A2[1::2] = ao
    B2[0::2] = b
    B2[1::2] = bo
    P = [x * y % R_ORDER for x, y in zip(A2, B2)]
    pc = ifft(P)
    return pc[n:]


# --------------------------------------------------------------------------
# binary layouts (tools/buildpkey.js:130-183, tools/buildwitness.js:36-69)
# --------------------------------------------------------------------------
def parse_pkey(pkey):
    hdr = struct.unpack_from("<10I", pkey, 0)
    names = ("nVars", "nPublic", "domainSize", "pPolsA", "pPolsB", "pPointsA",
             "pPointsB1", "pPointsB2", "pPointsC", "pPointsHExps")
    return dict(zip(names, hdr))


def rd_fq_m(buf, off):
    return int.from_bytes(buf[off:off + 32], "little")


def rd_g1(buf, off):
    """64-byte affine Montgomery point -> Jacobian over Fq (plain)."""
    return load_affine(Fq, from_mont(rd_fq_m(buf, off)), from_mont(rd_fq_m(buf, off + 32)))


def rd_g2(buf, off):
    x = (from_mont(rd_fq_m(buf, off)), from_mont(rd_fq_m(buf, off + 32)))
    y = (from_mont(rd_fq_m(buf, off + 64)), from_mont(rd_fq_m(buf, off + 96)))
    return load_affine(Fq2, x, y)


def rd_scalars(buf, n=None, off=0):
    n = (len(buf) - off) // 32 if n is None else n
    return [int.from_bytes(buf[off + 32 * i:off + 32 * i + 32], "little") for i in range(n)]


def g1_to_bytes_m(p):
    """Jacobian (plain) -> 96 bytes Montgomery, the g1_multiexp result format."""
    return b"".join(to_mont(c).to_bytes(32, "little") for c in p)


def g2_to_bytes_m(p):
    return b"".join(to_mont(c).to_bytes(32, "little") for cc in p for c in cc)


# --------------------------------------------------------------------------
# a11  Bn128.groth16GenProof (src/bn128.js:580-720)
# --------------------------------------------------------------------------
def groth16_gen_proof(witness, pkey, r=0, s=0, return_parts=False):
    """witness, pkey: bytes; r, s: integers in [0, 2^256) (the reference draws
    32 random bytes each, src/bn128.js:642-664).  Returns the proof object with
    decimal strings (src/bn128.js:714-718)."""
    k = parse_pkey(pkey)
    N, npub, n = k["nVars"], k["nPublic"], k["domainSize"]
    sig = rd_scalars(witness, N)
    polsA = pkey[k["pPolsA"]:k["pPolsB"]]
    polsB = pkey[k["pPolsB"]:k["pPointsA"]]
    A = [rd_g1(pkey, k["pPointsA"] + 64 * i) for i in range(N)]
    B1 = [rd_g1(pkey, k["pPointsB1"] + 64 * i) for i in range(N)]
    B2 = [rd_g2(pkey, k["pPointsB2"] + 128 * i) for i in range(N)]
    C = [rd_g1(pkey, k["pPointsC"] + 64 * i) for i in range(N - npub - 1)]
    H = [rd_g1(pkey, k["pPointsHExps"] + 64 * i) for i in range(n)]
    alfa1, beta1, delta1 = (rd_g1(pkey, 40 + 64 * i) for i in range(3))
    beta2, delta2 = rd_g2(pkey, 232), rd_g2(pkey, 360)
    # loadPoint1/2 always set Z = 1 for the five fixed points (src/bn128.js:441-453)
    alfa1, beta1, delta1 = [(p[0], p[1], 1) for p in (alfa1, beta1, delta1)]
    beta2, delta2 = [(p[0], p[1], Fq2.one) for p in (beta2, delta2)]

    h = calc_h(sig, polsA, polsB, N, n)
    sH = multiexp(Fq, h, H)
    sA = multiexp(Fq, sig, A)
    sB1 = multiexp(Fq, sig, B1)
    sB2 = multiexp(Fq2, sig, B2)
    sC = multiexp(Fq, sig[npub + 1:], C)

    pi_a = pt_add(Fq, pt_times_scalar(Fq, delta1, r), pt_add(Fq, alfa1, sA))          # :670-673
    pi_b = pt_add(Fq2, pt_times_scalar(Fq2, delta2, s), pt_add(Fq2, beta2, sB2))      # :675-678
    pib1 = pt_add(Fq, pt_times_scalar(Fq, delta1, s), pt_add(Fq, beta1, sB1))         # :680-683
    pi_c = pt_add(Fq, sH, sC)                                                         # :686-688
    pi_c = pt_add(Fq, pt_times_scalar(Fq, pi_a, s), pi_c)                             # :691-693
    pi_c = pt_add(Fq, pt_times_scalar(Fq, pib1, r), pi_c)                             # :695-697
    pi_c = pt_add(Fq, pt_neg(Fq, pt_times_scalar(Fq, delta1, int_mul(r, s))), pi_c)   # :699-704
    pa, pb, pc = pt_affine(Fq, pi_a), pt_affine(Fq2, pi_b), pt_affine(Fq, pi_c)       # :706-712
    proof = {
        "pi_a": [str(v) for v in pa],
        "pi_b": [[str(v) for v in cc] for cc in pb],
        "pi_c": [str(v) for v in pc],
    }
    if return_parts:
        return proof, dict(h=h, A=sA, B1=sB1, B2=sB2, C=sC, H=sH)
    return proof
