"""Decimal wire format kernels (pai_radix.cuh) on the simulation engine against Python's own str() / int()."""
import importlib
import json
import random

import numpy as np
import pytest

from oracle.golden import H, load_golden


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    return pkg.Engine(ge.build_hostsim())


@pytest.mark.parametrize("limbs", [1, 2, 7, 16, 64])
def test_limbs_to_decimal_and_back(pkg, sim, limbs):
    eng = importlib.import_module("python-paillier_b200.engine")
    rng = random.Random(limbs)
    top = 2 ** (32 * limbs)
    vals = [0, 1, 999999999, 10 ** 9, top - 1, top // 2, 10 ** (len(str(top)) - 1)] + \
           [rng.getrandbits(rng.randrange(1, 32 * limbs + 1)) for _ in range(40)]
    vals = [v for v in vals if v < top]
    arr = pkg.ints_to_limbs(vals, limbs)
    width = eng.decimal_width(limbs, sim)
    assert width % 9 == 0 and width >= len(str(top - 1))
    text = np.zeros((len(vals), width), dtype=np.uint8)
    eng.limbs_to_decimal_dev(arr, limbs, text, len(vals), engine=sim)
    got = [bytes(row).decode() for row in text]
    assert got == [str(v).rjust(width, "0") for v in vals]
    # and back, through a width that is not a multiple of nine
    w2 = width + 4
    text2 = np.frombuffer(b"".join(str(v).rjust(w2, "0").encode() for v in vals), dtype=np.uint8).reshape(len(vals), w2).copy()
    out = np.full((len(vals), limbs), 0xdeadbeef, dtype=np.uint32)
    status = np.full((len(vals),), -1, dtype=np.int32)
    eng.decimal_to_limbs_dev(text2, w2, out, limbs, status, len(vals), engine=sim)
    assert not status.any() and pkg.limbs_to_ints(out) == vals


def test_decimal_errors(pkg, sim):
    eng = importlib.import_module("python-paillier_b200.engine")
    limbs = 2
    rows = [b"00000000000000000012", b"0000000000000000001x", str(2 ** 64).encode().rjust(20, b"0"),
            str(2 ** 64 - 1).encode().rjust(20, b"0"), b"-0000000000000000001"]
    text = np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), 20).copy()
    out = np.zeros((len(rows), limbs), dtype=np.uint32)
    status = np.zeros((len(rows),), dtype=np.int32)
    eng.decimal_to_limbs_dev(text, 20, out, limbs, status, len(rows), engine=sim)
    assert status.tolist() == [0, 1, 2, 0, 1]
    assert pkg.limbs_to_ints(out) == [12, 0, 0, 2 ** 64 - 1, 0]


def test_json_scheme_matches_reference_format(pkg, sim):
    engine_mod = importlib.import_module("python-paillier_b200.engine")
    engine_mod._set_engine_for_tests(sim)
    try:
        fx = load_golden("vectors_256.json")
        pk = pkg.PaillierPublicKey(H(fx["n"]))
        sk = pkg.PaillierPrivateKey(pk, H(fx["p"]), H(fx["q"]))
        v = pk.encrypt_batch([1.5, -2.25, 0.0, 7.0], r_values=[3, 5, 7, 11])
        js = v.to_json(be_secure=False)
        # exactly what docs/serialisation.rst:24-31 produces with json.dumps
        want = json.dumps({"public_key": {"n": pk.n},
                           "values": [(str(c), int(e)) for c, e in zip(v.ciphertexts(False), v.exponents)]})
        assert js == want
        back = pkg.EncryptedVector.from_json(js)
        assert back.ciphertexts(False) == v.ciphertexts(False) and sk.decrypt_batch(back) == [1.5, -2.25, 0.0, 7.0]
        # values at or above n^2 come back reduced; junk is rejected
        d = json.loads(js)
        d["values"][0][0] = str(int(d["values"][0][0]) + pk.nsquare)
        assert pkg.EncryptedVector.from_json(json.dumps(d)).ciphertexts(False) == v.ciphertexts(False)
        d["values"][1][0] = "12a4"
        with pytest.raises(ValueError):
            pkg.EncryptedVector.from_json(json.dumps(d))
    finally:
        engine_mod._set_engine_for_tests(None)
