"""Pins oracle/mnt6753.py: the reference's hard-coded known answer (test/mnt6753.js:124-155), r*G = infinity
(:199-221), and its compiled WASM image (f1m_mul, f3m_mul, g1m_add / g1m_timesScalar / g1m_affine).  CPU only."""
import os
import random

import pytest

from oracle import mnt6753 as M
from oracle import ref_wasm

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libmnt6753_ref.so")

P1 = (2071893303198007985737678972190309212568452221625132024511988170095494148670997278812694070338313361389889122280160253462982652030041813566301365289695187505618174204273471887226695702458395861269694368663558765191107385382142,
      17187187414417664367585796530257262302159176591062800465884265459977066325098901507827719965058588341044788483232395252403515861767227243983849894797683644816538861625368393588001624014759720661490214325432345769098675755344007)
P2 = (12212700530208157134689256057121042620633735483309261868159828729358269133353025097021648766749096328904625282610227267815597560656189994727627613599055093979638719153187781645363642530065802177696707603573183038198049837281284,
      25783690089010390455572974279288664362239817189553105020262800542618336981545722940934884637537924384027963108482794866752096889610121505188228870561534980224779056104062778099246844511975082198175449341056361726557212816273203)
P12 = (24530173387584474493637199745573842996932458279480996633536599083023400776896850895927928300031540755827270957795883601164540695965966931068707096855134911813023933604810285308821937900908148310002633109781840767650016069631684,
       2176866345050373929048969848689325360044216568746784615606447613684338532909692353308386803545656617464229816849438838053550146942264975470731462382861977242181149443006406892148357437369226748506091130883950310503563066671343)


def test_known_answer_and_group_order():
    assert M.on_curve(M.G1_GEN) and M.on_curve(P1) and M.on_curve(P2)
    assert M.add(P1, P2) == P12                              # test/mnt6753.js:124-155
    assert M.times_scalar(M.G1_GEN, M.R_ORDER) is None       # test/mnt6753.js:199-221
    assert M.add(M.G1_GEN, None) == M.G1_GEN


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built")
def test_against_reference_image():
    rnd = random.Random(11)
    inst = ref_wasm.RefInstance(pages=1000, lib_path=LIB)
    px, py, pr = inst.alloc(288), inst.alloc(288), inst.alloc(288)
    b96 = lambda v: v.to_bytes(96, "little")
    for _ in range(50):
        x, y = rnd.randrange(M.Q), rnd.randrange(M.Q)
        inst.put(px, b96(x))
        inst.put(py, b96(y))
        inst.call("f1m_mul", px, py, pr)
        assert int.from_bytes(inst.get(pr, 96), "little") == M.mont_mul(x, y)
    for _ in range(20):
        x = [rnd.randrange(M.Q) for _ in range(3)]
        y = [rnd.randrange(M.Q) for _ in range(3)]
        inst.put(px, b"".join(b96(M.to_mont(v)) for v in x))
        inst.put(py, b"".join(b96(M.to_mont(v)) for v in y))
        inst.call("f3m_mul", px, py, pr)
        got = inst.get(pr, 288)
        assert tuple(M.from_mont(int.from_bytes(got[96 * i:96 * i + 96], "little")) for i in range(3)) == M.f3_mul(x, y)

    def put_pt(p, P):
        inst.put(p, M.affine_bytes_m(P) + b96(M.to_mont(1)))

    def get_aff(p):
        inst.call("g1m_affine", p, p)
        g = inst.get(p, 288)
        return tuple(M.from_mont(int.from_bytes(g[96 * i:96 * i + 96], "little")) for i in range(3))
    pk = inst.alloc(96)
    for _ in range(4):
        A = M.times_scalar(M.G1_GEN, rnd.randrange(1, M.R_ORDER))
        Bp = M.times_scalar(M.G1_GEN, rnd.randrange(1, M.R_ORDER))
        put_pt(px, A)
        put_pt(py, Bp)
        inst.call("g1m_add", px, py, pr)
        assert get_aff(pr) == M.affine_result(M.add(A, Bp))
        k = rnd.getrandbits(753)
        inst.put(pk, k.to_bytes(96, "little"))
        put_pt(px, A)
        inst.call("g1m_timesScalar", px, pk, 96, pr)
        assert get_aff(pr) == M.affine_result(M.times_scalar(A, k))
    inst.close()
