"""oracle/mnt6753.py -- TEST INFRASTRUCTURE (oracle), not product code.

CPU restatement (Python big integers) of the reference's MNT6753 wide-limb pieces on the hot path:
  f1m_mul over q (R = 2^768)          /root/reference/src/build_f1m.js:235-436, mnt6753/build_mnt6753.js:19
  f3m_mul, Fq3 = Fq[u]/(u^3 - 11)     /root/reference/src/build_f3m.js:125-189, build_mnt6753.js:35-56
  G1: y^2 = x^3 + 11 x + b            /root/reference/src/build_curve_projective.js:202-516 (a = 11: build_mnt6753.js:90-109,
                                       b: :1402, generator :146-150, group order r :234)
The reference has no multiexp for this curve; the MSM oracle is the definitional sum.
Pinned by tests/test_oracle_mnt.py against the reference's hard-coded known answer P1 + P2
(test/mnt6753.js:124-155), r*G = infinity (:199-221) and the reference image (oracle/_ref/libmnt6753_ref.so).
The group law is restated in affine chord-and-tangent form: the reference's projective formulas compute
the same group element and results are compared after g1m_affine."""

Q = 0x01C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB26C5C28C859A99B3EEBCA9429212636B9DFF97634993AA4D6C381BC3F0057974EA099170FA13A4FD90776E240000001
R_ORDER = 0x01C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB117E776F218059DB80F0DA5CB537E38685ACCE9767254A4638810719AC425F0E39D54522CDD119F5E9063DE245E8001
RMONT = 1 << 768
A = 11
B = 11625908999541321152027340224010374716841167701783584648338908235410859267060079819722747939267925389062611062156601938166010098747920378738927832658133625454260115409075816187555055859490253375704728027944315501122723426879114
G1_GEN = (16364236387491689444759057944334173579070747473738339749093487337644739228935268157504218078126401066954815152892688541654726829424326599038522503517302466226143788988217410842672857564665527806044250003808514184274233938437290,
          4510127914410645922431074687553594593336087066778984214797709122300210966076979927285161950203037801392624582544098750667549188549761032654706830225743998064330900301346566408501390638273322467173741629353517809979540986561128)
NR3 = 11
_RINV = pow(RMONT, -1, Q)


def to_mont(x):
    return x * RMONT % Q


def from_mont(x):
    return x * _RINV % Q


def mont_mul(x, y):
    return x * y * _RINV % Q


def f3_mul(x, y):
    """build_f3m.js:158-186 on plain residues."""
    a, b, c = x
    A_, B_, C_ = y
    aA, bB, cC = a * A_ % Q, b * B_ % Q, c * C_ % Q
    r0 = (aA + NR3 * ((b + c) * (B_ + C_) - bB - cC)) % Q
    r1 = ((a + b) * (A_ + B_) - aA - bB + NR3 * cC) % Q
    r2 = ((a + c) * (A_ + C_) - aA - cC + bB) % Q
    return (r0, r1, r2)


INF = None          # affine point or None for infinity


def on_curve(P):
    return P is None or (P[1] * P[1] - (P[0] ** 3 + A * P[0] + B)) % Q == 0


def add(P, Qp):
    if P is None:
        return Qp
    if Qp is None:
        return P
    x1, y1 = P
    x2, y2 = Qp
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1 + A) * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


def neg(P):
    return None if P is None else (P[0], (-P[1]) % Q)


def times_scalar(P, k):
    """g1m_timesScalar semantics: k a non-negative integer of any size."""
    res = None
    for i in reversed(range(k.bit_length())):
        res = add(res, res)
        if (k >> i) & 1:
            res = add(res, P)
    return res


def multiexp_naive(scalars, points):
    acc = None
    for s, P in zip(scalars, points):
        acc = add(acc, times_scalar(P, s))
    return acc


def affine_bytes_m(P):
    """192-byte affine Montgomery x|y; infinity -> all zero (x == 0 rule)."""
    if P is None:
        return bytes(192)
    return to_mont(P[0]).to_bytes(96, "little") + to_mont(P[1]).to_bytes(96, "little")


def affine_result(P):
    """what g1m_affine + fromMontgomery returns: (x, y, 1) or (0, 1, 0)."""
    return (0, 1, 0) if P is None else (P[0], P[1], 1)
