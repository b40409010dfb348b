"""The oracle (oracle/paillier_oracle.py, oracle/paillier_oracle.c) pinned against the reference:
its own known answers and the golden vectors generated from the unmodified reference."""
import ctypes
import os

import numpy as np
import pytest

from oracle.golden import H, load_golden
from oracle import paillier_oracle as orc

KEYS = [64, 256, 512, 1024, 2048, 3072]


@pytest.fixture(scope="module")
def coracle():
    import __graft_entry__ as ge
    return ctypes.CDLL(ge.build_oracle())


def limbs(vals, L):
    return np.frombuffer(b"".join(v.to_bytes(4 * L, "little") for v in vals), dtype=np.uint32).reshape(len(vals), L).copy()


def ints(arr):
    raw = arr.tobytes()
    nb = 4 * arr.shape[1]
    return [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(arr.shape[0])]


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_reference_known_answers():
    """phe/tests/paillier_test.py:128-149, phe/tests/util_test.py:31-44"""
    kat = load_golden("kat_reference_tests.json")
    pub = orc.PublicConsts(kat["n"])
    priv = orc.PrivateConsts(pub, kat["p"], kat["q"])
    assert orc.raw_encrypt(pub, kat["m"], kat["r"]) == kat["c"] == 935906717
    assert orc.raw_decrypt(priv, kat["c"]) == kat["m"]
    assert orc.raw_encrypt(pub, 1, 1) == kat["encrypt_1_r_1"] == 126870
    assert (priv.psquare, priv.qsquare, priv.p_inverse, priv.hp, priv.hq) == (85849, 187489, 300, 203, 133)
    for a, b, c, o in kat["powmod"]:
        assert orc.powmod(a, b, c) == o
    assert [orc.invert(a, 101) for a in range(1, 101)] == kat["invert_mod_101"]
    for a in range(1, 101):
        assert orc.invert(a, 101) * a % 101 == 1
    assert orc.invert(1, 4) == 1 and orc.invert(3, 4) == 3
    with pytest.raises(ZeroDivisionError):
        orc.invert(2, 4)


@pytest.mark.parametrize("backend", ["python", "gmp"])
@pytest.mark.parametrize("kb", KEYS)
def test_python_oracle_vs_golden(kb, backend):
    if backend == "gmp" and not orc.have_gmp():
        pytest.skip("libgmp not present")
    if backend == "python" and kb > 2048:
        pytest.skip("slow; covered by the gmp backend")
    fx = load_golden("vectors_%d.json" % kb)
    orc.BACKEND = backend
    try:
        pub = orc.PublicConsts(H(fx["n"]))
        priv = orc.PrivateConsts(pub, H(fx["q"]), H(fx["p"]))
        assert (priv.p, priv.q, priv.p_inverse, priv.hp, priv.hq) == tuple(H(fx[k]) for k in ("p", "q", "p_inverse", "hp", "hq"))
        step = 1 if kb <= 1024 or backend == "gmp" else 4
        for e in fx["encrypt"][::step]:
            assert orc.raw_encrypt(pub, H(e["m"]), H(e["r"])) == H(e["c"])
            assert orc.raw_decrypt(priv, H(e["c"])) == H(e["d"])
        for e in fx["decrypt_any"][::step]:
            assert orc.raw_decrypt(priv, H(e["c"])) == H(e["d"])
        for e in fx["add"]:
            assert orc.raw_add(pub, H(e["a"]), H(e["b"])) == H(e["s"])
        for e in fx["mul"][::step]:
            if "error" in e:
                with pytest.raises(ZeroDivisionError):
                    orc.raw_mul(pub, H(e["c"]), H(e["k"]))
            else:
                assert orc.raw_mul(pub, H(e["c"]), H(e["k"])) == H(e["o"])
    finally:
        orc.BACKEND = "python"


@pytest.mark.parametrize("kb", [64, 256, 1024, 2048])
def test_c_oracle_vs_golden(coracle, kb):
    fx = load_golden("vectors_%d.json" % kb)
    n, p, q = H(fx["n"]), H(fx["p"]), H(fx["q"])
    ln = max(2, (kb + 31) // 32)
    ln += ln % 2
    lp = ln // 2
    N = limbs([n], ln)
    enc = [e for e in fx["encrypt"] if 0 <= H(e["m"]) < 2 ** (32 * ln)][: 12 if kb >= 2048 else 64]
    m, r = limbs([H(e["m"]) for e in enc], ln), limbs([H(e["r"]) for e in enc], ln)
    out = np.zeros((len(enc), 2 * ln), dtype=np.uint32)
    assert coracle.orc_raw_encrypt(P(N), ln, P(m), P(r), P(out), ctypes.c_long(len(enc))) == 0
    assert ints(out) == [H(e["c"]) for e in enc]
    dec = np.zeros((len(enc), ln), dtype=np.uint32)
    assert coracle.orc_raw_decrypt(P(N), ln, P(limbs([p], lp)), P(limbs([q], lp)), lp, P(out), P(dec), ctypes.c_long(len(enc))) == 0
    assert ints(dec) == [H(e["d"]) for e in enc]
    da = fx["decrypt_any"]
    c = limbs([H(e["c"]) for e in da], 2 * ln)
    dec = np.zeros((len(da), ln), dtype=np.uint32)
    assert coracle.orc_raw_decrypt(P(N), ln, P(limbs([p], lp)), P(limbs([q], lp)), lp, P(c), P(dec), ctypes.c_long(len(da))) == 0
    assert ints(dec) == [H(e["d"]) for e in da]
    add = fx["add"]
    a, b = limbs([H(e["a"]) for e in add], 2 * ln), limbs([H(e["b"]) for e in add], 2 * ln)
    s = np.zeros((len(add), 2 * ln), dtype=np.uint32)
    assert coracle.orc_raw_add(P(N), ln, P(a), P(b), P(s), ctypes.c_long(len(add))) == 0
    assert ints(s) == [H(e["s"]) for e in add]
    mul = fx["mul"][: 20 if kb >= 2048 else 200]
    cc, kk = limbs([H(e["c"]) for e in mul], 2 * ln), limbs([H(e["k"]) for e in mul], ln)
    o = np.zeros((len(mul), 2 * ln), dtype=np.uint32)
    st = np.zeros(len(mul), dtype=np.int32)
    assert coracle.orc_raw_mul(P(N), ln, P(cc), P(kk), P(o), P(st), ctypes.c_long(len(mul))) == 0
    for g, sflag, e in zip(ints(o), st.tolist(), mul):
        if "error" in e:
            assert sflag == 1
        else:
            assert sflag == 0 and g == H(e["o"])


def test_c_oracle_seam(coracle):
    kat = load_golden("kat_reference_tests.json")
    for a, b, c, o in kat["powmod"]:
        out = np.zeros((1, 2), dtype=np.uint32)
        assert coracle.orc_powmod(P(limbs([a], 2)), 2, P(limbs([b], 2)), 2, P(limbs([c], 2)), 2, P(out)) == 0
        assert ints(out) == [o]
    for a in range(1, 101):
        out = np.zeros((1, 2), dtype=np.uint32)
        assert coracle.orc_invert(P(limbs([a], 2)), 2, P(limbs([101], 2)), 2, P(out)) == 0
        assert ints(out) == [kat["invert_mod_101"][a - 1]]
