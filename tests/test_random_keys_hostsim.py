"""Random keys of many sizes (8 ... 320 bits, like the key-length sweep of phe/tests/paillier_test.py:49)
through the C ABI on the simulation engine, against the oracle: exercises limb padding and the choice of
tile counts for moduli that do not fill their limbs."""
import importlib
import math
import random

import pytest

from oracle import paillier_oracle as orc


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    return pkg.Engine(ge.build_hostsim())


@pytest.mark.parametrize("bits", [8, 12, 16, 24, 32, 48, 64, 96, 128, 160, 200, 256, 320])
def test_random_key_roundtrip_vs_oracle(pkg, sim, bits):
    util = importlib.import_module("python-paillier_b200.util")
    rng = random.Random(bits)
    for trial in range(2):
        while True:
            p, q = util.getprimeover(bits // 2), util.getprimeover(bits - bits // 2)
            if p != q:
                break
        n = p * q
        pub, priv = pkg.PublicContext(n, engine=sim), pkg.PrivateContext(p, q, engine=sim)
        opub = orc.PublicConsts(n)
        opriv = orc.PrivateConsts(opub, p, q)
        assert (priv.p, priv.q, priv.p_inverse, priv.hp, priv.hq) == (opriv.p, opriv.q, opriv.p_inverse, opriv.hp, opriv.hq)
        ms = [0, 1, n - 1, n // 2] + [rng.randrange(n) for _ in range(5)]
        rs = [1, n - 1] + [rng.randrange(1, n) for _ in range(len(ms) - 2)]
        cs = pub.raw_encrypt(ms, rs)
        assert cs == [orc.raw_encrypt(opub, m, r) for m, r in zip(ms, rs)]
        dec = priv.raw_decrypt(cs)
        assert dec == [orc.raw_decrypt(opriv, c) for c in cs]
        assert all(d == m for d, m, r in zip(dec, ms, rs) if math.gcd(r, n) == 1)      # r must be a unit mod n
        assert priv.raw_decrypt([0, 1, n, n * n - 1, p, q]) == [orc.raw_decrypt(opriv, c) for c in (0, 1, n, n * n - 1, p, q)]
        assert pub.raw_add(cs, cs[::-1]) == [orc.raw_add(opub, a, b) for a, b in zip(cs, cs[::-1])]
        ks = [0, 1, n - 1, opub.max_int, n - opub.max_int] + [rng.randrange(n) for _ in range(len(ms) - 5)]
        ks = [k % n for k in ks]
        out, st = pub.raw_mul(cs, ks)
        for c, k, o, s in zip(cs, ks, out, st):
            try:
                exp = orc.raw_mul(opub, c, k)
                assert s == 0 and o == exp
            except ZeroDivisionError:
                assert s == 1
        pub.close(); priv.close()
