"""CPU tests of the host logic: the phe-compatible Python layer + the C-ABI orchestration, run on the
TEST-ONLY host simulation of the kernels (tests/hostsim), against the reference-generated fixtures.
The product never loads the simulation library; these tests install it explicitly."""
import math
import os
import random

import pytest

from oracle.golden import H, load_golden


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    from importlib import import_module
    engine_mod = import_module("python-paillier_b200.engine")
    eng = pkg.Engine(ge.build_hostsim())
    engine_mod._set_engine_for_tests(eng)
    yield eng
    engine_mod._set_engine_for_tests(None)
    import_module("python-paillier_b200.util")._ctx_cache.clear()


def test_reference_kat_through_api(pkg, sim):
    """phe/tests/paillier_test.py:128-155"""
    pk = pkg.PaillierPublicKey(126869)
    sk = pkg.PaillierPrivateKey(pk, 293, 433)
    assert pk.raw_encrypt(10100, 74384) == 935906717
    assert sk.raw_decrypt(935906717) == 10100
    enc = pk.encrypt(10100, r_value=74384)
    assert enc.ciphertext(be_secure=False) == 935906717
    assert pk.encrypt(1, r_value=1).ciphertext(False) == 126870
    assert pk.encrypt(1).ciphertext(False) != 126870 or pk.encrypt(1).ciphertext(False) != 126870
    assert (sk.psquare, sk.qsquare, sk.p_inverse, sk.hp, sk.hq) == (85849, 187489, 300, 203, 133)
    with pytest.raises(TypeError):
        pk.raw_encrypt("123")
    with pytest.raises(TypeError):
        sk.raw_decrypt("123")
    # wrap-around of large plaintexts (phe/tests/paillier_test.py:114-126)
    for m in (pk.n - 1, pk.n, pk.n + 1):
        assert sk.raw_decrypt(pk.raw_encrypt(m)) == m % pk.n


def test_private_key_checks(pkg, sim):
    pk = pkg.PaillierPublicKey(293 * 433)
    with pytest.raises(ValueError):
        pkg.PaillierPrivateKey(pk, 293, 431)
    with pytest.raises(ValueError):
        pkg.PaillierPrivateKey(pkg.PaillierPublicKey(293 * 293), 293, 293)
    a, b = pkg.PaillierPrivateKey(pk, 293, 433), pkg.PaillierPrivateKey(pk, 433, 293)
    assert a == b and hash(a) == hash(b) and a.p == 293
    tot = pkg.PaillierPrivateKey.from_totient(pk, 292 * 432)
    assert tot == a
    with pytest.raises(ValueError):
        pkg.PaillierPrivateKey.from_totient(pk, 292 * 432 + 1)


def test_util_seam(pkg, sim):
    """phe/tests/util_test.py:29-58"""
    u = pkg.util
    assert u.powmod(5, 3, 3) == 2 and u.powmod(2, 10, 1000) == 24
    for a in range(1, 101):
        assert u.invert(a, 101) * a % 101 == 1
    assert u.invert(1, 4) == 1 and u.invert(3, 4) == 3
    with pytest.raises(ZeroDivisionError):
        u.invert(2, 4)
    fx = load_golden("vectors_256.json")
    for e in fx["seam"]["powmod"]:
        assert u.powmod(H(e["a"]), H(e["b"]), H(e["c"])) == H(e["o"])
    for e in fx["seam"]["mulmod"]:
        assert u.mulmod(H(e["a"]), H(e["b"]), H(e["c"])) == H(e["o"])
    for e in fx["seam"]["invert"]:
        if "error" in e:
            with pytest.raises(ZeroDivisionError):
                u.invert(H(e["a"]), H(e["b"]))
        else:
            assert u.invert(H(e["a"]), H(e["b"])) == H(e["o"])
    assert u.base64_to_int(u.int_to_base64(123456789 ** 7)) == 123456789 ** 7
    assert u.isqrt(10 ** 40 + 12345) == math.isqrt(10 ** 40 + 12345)
    p = u.getprimeover(64)
    assert p.bit_length() == 64 and u.is_prime(p)


def test_encoding_matches_reference(pkg, sim):
    fx = load_golden("api_1024.json")
    pk = pkg.PaillierPublicKey.__new__(pkg.PaillierPublicKey)
    pk.g, pk.n, pk.nsquare, pk.max_int, pk._ctx = H(fx["n"]) + 1, H(fx["n"]), H(fx["n"]) ** 2, H(fx["n"]) // 3 - 1, None
    for rec in fx["encode"]:
        v = eval(rec["v"])
        enc = pkg.EncodedNumber.encode(pk, v)
        assert (enc.encoding, enc.exponent) == (H(rec["encoding"]), rec["exponent"])
        assert repr(enc.decode()) == rec["decoded"]
        for key, kw in (("prec_1e-6", {"precision": 1e-6}), ("maxexp_-20", {"max_exponent": -20})):
            if key in rec:
                if rec[key] == "ValueError":
                    with pytest.raises(ValueError):
                        pkg.EncodedNumber.encode(pk, v, **kw)
                else:
                    e2 = pkg.EncodedNumber.encode(pk, v, **kw)
                    assert [e2.encoding, e2.exponent] == [H(rec[key][0]), rec[key][1]]
    with pytest.raises(ValueError):
        pkg.EncodedNumber.encode(pk, pk.max_int + 1)
    with pytest.raises(OverflowError):
        pkg.EncodedNumber(pk, pk.max_int + 5, 0).decode()
    with pytest.raises(ValueError):
        pkg.EncodedNumber(pk, pk.n, 0).decode()
    with pytest.raises(ValueError):
        pkg.EncodedNumber.encode(pk, 1.0).decrease_exponent_to(5)
    assert pkg.EncodedNumber(pk, 1, -2000).decode() == 0.0       # phe/tests/paillier_test.py:1088-1095


def test_operator_semantics_small_key(pkg, sim):
    """EncryptedNumber operators incl. exponent alignment, lazy obfuscation and errors
    (phe/tests/paillier_test.py:430-1058) on a 256-bit key so that the simulation is quick."""
    fx = load_golden("vectors_256.json")
    pk = pkg.PaillierPublicKey(H(fx["n"]))
    sk = pkg.PaillierPrivateKey(pk, H(fx["p"]), H(fx["q"]))
    rng = random.Random(3)
    for a, b in [(1.5, 2.25), (3, 4), (-7, 2.5), (0.1, 0.2), (1e-5, 123456), (-1.25, -3.5), (2 ** 40, -0.375)]:
        ea, eb = pk.encrypt(a), pk.encrypt(b, r_value=rng.randrange(1, pk.n))
        assert sk.decrypt(ea + eb) == pytest.approx(a + b, rel=1e-12)
        assert sk.decrypt(ea + b) == pytest.approx(a + b, rel=1e-12)
        assert sk.decrypt(b + ea) == pytest.approx(a + b, rel=1e-12)
        assert sk.decrypt(ea - eb) == pytest.approx(a - b, rel=1e-12)
        assert sk.decrypt(b - ea) == pytest.approx(b - a, rel=1e-12)
        assert sk.decrypt(ea * b) == pytest.approx(a * b, rel=1e-12)
        assert sk.decrypt(b * ea) == pytest.approx(a * b, rel=1e-12)
        assert sk.decrypt(ea / 4) == pytest.approx(a / 4, rel=1e-12)
        assert sk.decrypt(ea + pkg.EncodedNumber.encode(pk, b)) == pytest.approx(a + b, rel=1e-12)
    e = pk.encrypt(5, r_value=7)
    assert (e * 1).ciphertext(False) == e.ciphertext(False)                      # phe/tests/paillier_test.py:893-899
    assert sk.decrypt(e * 0) == 0 and sk.decrypt(e * -1) == -5
    with pytest.raises(NotImplementedError):
        e * e
    with pytest.raises(ValueError):
        e._raw_mul(pk.n)
    with pytest.raises(TypeError):
        e._raw_mul(1.5)
    with pytest.raises(ValueError):
        e.decrease_exponent_to(3)
    other = pkg.PaillierPublicKey(293 * 433)
    with pytest.raises(ValueError):
        e + other.encrypt(1)
    with pytest.raises(ValueError):
        sk.decrypt(other.encrypt(1))
    with pytest.raises(TypeError):
        sk.decrypt(12)
    with pytest.raises(TypeError):
        pkg.EncryptedNumber(12, 5)
    # lazy obfuscation state machine (phe/tests/paillier_test.py:1012-1049)
    s = e + 1
    c0 = s.ciphertext(be_secure=False)
    c1 = s.ciphertext()
    assert c0 != c1 and s.ciphertext() == c1 and sk.decrypt(s) == 6
    fresh = pk.encrypt(9)
    assert fresh.ciphertext(False) == fresh.ciphertext(True)
    # sum() / keyring
    vals = [pk.encrypt(x) for x in (1, 2.5, -3)]
    assert sk.decrypt(sum(vals)) == pytest.approx(0.5)
    ring = pkg.PaillierPrivateKeyring([sk])
    assert ring.decrypt(vals[0]) == 1 and len(ring) == 1 and ring[pk] == sk
    with pytest.raises(KeyError):
        ring.decrypt(other.encrypt(1))
    with pytest.raises(TypeError):
        ring.add("x")
    # overflow detection (phe/tests/paillier_test.py:608-620)
    big = pk.encrypt(pk.max_int)
    with pytest.raises(OverflowError):
        sk.decrypt(big + big)


@pytest.mark.parametrize("nrows", [12])
def test_config1_rows_and_api_fixture(pkg, sim, nrows):
    """BASELINE configs[0] plumbing fixture (1024-bit, int32 plaintexts, injected r): ciphertexts identical
    to the reference's, and the operator fixture with injected r (subset: the simulation is slow)."""
    c1 = load_golden("config1_1024.json")
    pk = pkg.PaillierPublicKey(H(c1["n"]))
    sk = pkg.PaillierPrivateKey(pk, H(c1["p"]), H(c1["q"]))
    for row in c1["rows"][:nrows]:
        e = pk.encrypt(row["x"], r_value=H(row["r"]))
        assert e.ciphertext(False) == H(row["c"]) and e.exponent == row["exponent"]
        assert sk.decrypt(e) == row["x"]
    api = load_golden("api_1024.json")
    for op in api["ops"][:3]:
        a, b = eval(op["a"]), eval(op["b"])
        ea, eb = pk.encrypt(a, r_value=H(op["ra"])), pk.encrypt(b, r_value=H(op["rb"]))
        assert [ea.ciphertext(False), ea.exponent] == [H(op["ea"][0]), op["ea"][1]]
        for name, val in (("add", ea + eb), ("add_scalar", ea + b), ("mul", ea * b), ("sub", ea - eb), ("div4", ea / 4)):
            assert [val.ciphertext(False), val.exponent] == [H(op[name][0]), op[name][1]], name
            assert repr(sk.decrypt(val)) == op[name][2], name
