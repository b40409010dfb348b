"""Key sizes that straddle limb / tile boundaries (n of 255 ... 4090 bits) and unbalanced prime pairs through the
C ABI on the simulation engine against the oracle -- the reference accepts any p != q (phe/paillier.py:217-229),
not only the equal-length pairs its own keygen makes."""
import importlib
import random

import pytest

from oracle import paillier_oracle as orc


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    orc.BACKEND = "gmp" if orc.have_gmp() else "python"
    yield pkg.Engine(ge.build_hostsim())
    orc.BACKEND = "python"


CASES = [(255, 127), (256, 128), (257, 128), (258, 100), (300, 40), (511, 255), (512, 256), (513, 256), (513, 200),
         (514, 257), (520, 130), (600, 64), (768, 384), (769, 384), (770, 300), (1025, 512), (1030, 257), (2049, 1024),
         (2050, 700), (3100, 1100), (4090, 2044)]


@pytest.mark.parametrize("bits,pbits", CASES)
def test_edge_key(pkg, sim, bits, pbits):
    util = importlib.import_module("python-paillier_b200.util")
    rng = random.Random(bits * 10007 + pbits)

    def prime_bits(b):
        while True:
            c = rng.getrandbits(b) | (1 << (b - 1)) | 1
            if util.is_prime(c):
                return c
    while True:
        p, q = prime_bits(pbits), prime_bits(bits - pbits)
        n = p * q
        if p != q and n.bit_length() == bits:
            break
    pub, priv = pkg.PublicContext(n, engine=sim), pkg.PrivateContext(p, q, engine=sim)
    opub = orc.PublicConsts(n)
    opriv = orc.PrivateConsts(opub, p, q)
    assert (priv.p, priv.q, priv.p_inverse, priv.hp, priv.hq) == (opriv.p, opriv.q, opriv.p_inverse, opriv.hp, opriv.hq)
    ms = [0, 1, n - 1, n // 2] + [rng.randrange(n) for _ in range(2)]
    rs = [1, n - 1] + [rng.randrange(1, n) for _ in range(len(ms) - 2)]
    cs = pub.raw_encrypt(ms, rs)
    assert cs == [orc.raw_encrypt(opub, m, r) for m, r in zip(ms, rs)]
    xs = cs + [0, 1, n, n * n - 1, p, q, p * p, q * q]
    assert priv.raw_decrypt(xs) == [orc.raw_decrypt(opriv, c) for c in xs]
    ks = [0, 1, n - 1, opub.max_int, n - opub.max_int, rng.getrandbits(64)]
    out, st = pub.raw_mul(cs, ks)
    for c, k, o, s in zip(cs, ks, out, st):
        try:
            want = orc.raw_mul(opub, c, k)
            assert s == 0 and o == want
        except ZeroDivisionError:
            assert s == 1
    assert pub.raw_add(cs, cs[::-1]) == [orc.raw_add(opub, a, b) for a, b in zip(cs, cs[::-1])]
    pub.close(); priv.close()


def test_size_limits_are_reported(pkg, sim):
    """n above 4096 bits, or a prime above 2048 bits, is refused with a clear error, not computed wrongly."""
    engine_mod = importlib.import_module("python-paillier_b200.engine")
    with pytest.raises(engine_mod.EngineError, match="too large"):
        pkg.PublicContext((1 << 4100) + 1, engine=sim)
    with pytest.raises(engine_mod.EngineError, match="too large"):
        pkg.PrivateContext((1 << 1000) + 1, (1 << 2100) + 1, engine=sim)
