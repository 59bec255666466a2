"""oracle/bls12381.py -- TEST INFRASTRUCTURE (oracle), not product code.  PARITY UNPINNED.

CPU restatement (Python big integers) of what the reference's bls12381 module computes on the hot path:
  q, r, R = 2^384 (f1m) / 2^256 (frm)   /root/reference/src/bls12381/build_bls12381.js:20-35, src/build_f1m.js:37-39
  f2m = f1m[u]/(u^2 + 1)                /root/reference/src/bls12381/build_bls12381.js:41, src/build_f2m.js:127-163
  g1m, g2m: y^2 = x^3 + b, a = 0        /root/reference/src/build_curve_jacobian_a0.js (generators build_bls12381.js:50-98)
  g?m_multiexp                          /root/reference/src/build_multiexp.js (definitional sum here)
  fft over frm                          /root/reference/src/build_fft.js:29-72 (roots), :159-372, :550-648

"Parity unpinned": the reference ships NO compiled bls12381 image (build/ holds bn128 and mnt6753 only), no JS host
drives this module, and test/bls12381.js needs that image, so nothing here could be checked against outputs of the
reference itself.  What pins it instead (tests/test_oracle_bls.py): the constants are the reference's own
(build_bls12381.js:20-21, :50-98) and satisfy the curve equations, r * G1 = O and r * G2 = O (the reference's own
"G1n == 0" / "G2n == 0" tests, test/bls12381.js:330-338, :381-389), and the Fr root table follows build_fft.js's
derivation (smallest non-residue, 2-adicity 32).  The group law is the textbook affine chord-and-tangent rule;
results are compared after g?m_affine, where representation does not matter.
"""
Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_ORDER = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
RQ = 1 << 384                      # Montgomery radix of f1m (n64 = 6)
RR = 1 << 256                      # Montgomery radix of frm (n64 = 4)
G1_GEN = (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
          1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569)
G2_GEN = ((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
           3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
          (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
           927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582))
B1 = 4                             # G1: y^2 = x^3 + 4
B2 = (4, 4)                        # G2: y^2 = x^3 + 4 (1 + u)


def to_mont(x, p=Q):
    return x * (RQ if p == Q else RR) % p


def from_mont(x, p=Q):
    return x * pow(RQ if p == Q else RR, -1, p) % p


def mont_mul(x, y, p=Q):
    """f1m_mul / frm_mul on Montgomery residues: x y R^-1 mod p (src/build_f1m.js:235-436)."""
    return x * y * pow(RQ if p == Q else RR, -1, p) % p


class Fq:
    zero, one = 0, 1
    add = staticmethod(lambda a, b: (a + b) % Q)
    sub = staticmethod(lambda a, b: (a - b) % Q)
    neg = staticmethod(lambda a: (-a) % Q)
    mul = staticmethod(lambda a, b: a * b % Q)
    inv = staticmethod(lambda a: pow(a, -1, Q))
    is_zero = staticmethod(lambda a: a == 0)
    b = B1


class Fq2:
    zero, one = (0, 0), (1, 0)
    add = staticmethod(lambda a, b: ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q))
    sub = staticmethod(lambda a, b: ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q))
    neg = staticmethod(lambda a: ((-a[0]) % Q, (-a[1]) % Q))
    mul = staticmethod(lambda a, b: ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q))
    is_zero = staticmethod(lambda a: a[0] == 0 and a[1] == 0)
    b = B2

    @staticmethod
    def inv(a):
        t = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
        return (a[0] * t % Q, (-a[1] * t) % Q)


# affine points: None = infinity
def on_curve(F, P):
    if P is None:
        return True
    x, y = P
    return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), F.b)) == F.zero


def pt_add(F, P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if F.add(y1, y2) == F.zero:
            return None
        three_x2 = F.mul(F.add(F.add(F.one, F.one), F.one), F.mul(x1, x1))
        lam = F.mul(three_x2, F.inv(F.add(y1, y1)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    return (x3, F.sub(F.mul(lam, F.sub(x1, x3)), y1))


def pt_neg(F, P):
    return None if P is None else (P[0], F.neg(P[1]))


def pt_mul(F, P, k):
    """g?m_timesScalar: k is used as a raw non-negative integer (not reduced mod r), like the reference's bit loop."""
    acc = None
    while k:
        if k & 1:
            acc = pt_add(F, acc, P)
        P = pt_add(F, P, P)
        k >>= 1
    return acc


def multiexp(F, scalars, points):
    """sum_i scalars[i] * points[i] (what g?m_multiexp computes, src/build_multiexp.js)."""
    acc = None
    for k, P in zip(scalars, points):
        acc = pt_add(F, acc, pt_mul(F, P, k))
    return acc


# ---- fft over frm (src/build_fft.js)
S_ADICITY = 32


def _nqr():
    nr = 2
    while pow(nr, R_ORDER >> 1, R_ORDER) == 1:      # src/build_fft.js:38-40
        nr += 1
    return nr


NQR = _nqr()


def roots():
    w = [0] * (S_ADICITY + 1)
    w[S_ADICITY] = pow(NQR, (R_ORDER - 1) >> S_ADICITY, R_ORDER)
    for i in reversed(range(S_ADICITY)):
        w[i] = w[i + 1] * w[i + 1] % R_ORDER
    return w


_ROOTS = roots()


def rawfft(x, odd):
    """bit reversal + radix-2 DIT; `odd` starts each stage's twiddle at w_{2m} (src/build_fft.js:223-372)."""
    n = len(x)
    bits = n.bit_length() - 1
    assert n == 1 << bits
    p = R_ORDER
    a = [0] * n
    for i in range(n):
        a[int(format(i, "0%db" % bits)[::-1], 2) if bits else 0] = x[i]
    for s in range(1, bits + 1):
        m, h = 1 << s, 1 << (s - 1)
        W = _ROOTS[s + 1] if odd else 1
        tw = []
        for _ in range(h):
            tw.append(W)
            W = W * _ROOTS[s] % p
        for k in range(0, n, m):
            for j in range(h):
                t = tw[j] * a[k + j + h] % p
                u = a[k + j]
                a[k + j] = (u + t) % p
                a[k + j + h] = (u - t) % p
    return a


def fft(x, odd=0):
    return rawfft(x, odd)


def ifft(x, odd=0):
    """rawfft then res[i] = raw[(n - i) mod n] / n (src/build_fft.js:189-221, :550-648)."""
    n = len(x)
    a = rawfft(x, odd)
    ninv = pow(n, -1, R_ORDER)
    return [a[(n - i) % n] * ninv % R_ORDER for i in range(n)]


# ---- byte formats of the C ABI (Montgomery, little-endian)
def g1_bytes(P):
    if P is None:
        return bytes(96)
    return to_mont(P[0]).to_bytes(48, "little") + to_mont(P[1]).to_bytes(48, "little")


def g2_bytes(P):
    if P is None:
        return bytes(192)
    return b"".join(to_mont(c).to_bytes(48, "little") for c in (P[0][0], P[0][1], P[1][0], P[1][1]))


def g1_from_affine_std(b):
    """144-byte (x, y, z) triple after g1m_affine + fromMontgomery -> affine point or None."""
    x, y, z = (int.from_bytes(b[48 * i:48 * i + 48], "little") for i in range(3))
    return None if z == 0 else (x, y)


def g2_from_affine_std(b):
    v = [int.from_bytes(b[48 * i:48 * i + 48], "little") for i in range(6)]
    return None if v[4] == 0 and v[5] == 0 else ((v[0], v[1]), (v[2], v[3]))
