"""Property tests (hypothesis) of the C-ABI seam on the simulation engine against Python ints: arbitrary
operands, exponents and odd moduli of awkward sizes (the batched form of phe/util.py:38-103)."""
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    return pkg.Engine(ge.build_hostsim())


@pytest.fixture(params=["thread-per-ciphertext", "warp-per-ciphertext"])
def kernel_family(request, monkeypatch):
    """Both kernel families behind the same entry points (PAI_COOP_MAX routes small batches to pai_coop.cuh)."""
    monkeypatch.setenv("PAI_COOP_MAX", "0" if request.param.startswith("thread") else "1000000")


odd_moduli = st.integers(min_value=3, max_value=2 ** 1100).map(lambda x: x | 1)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(mod=odd_moduli, data=st.data())
def test_seam_matches_python(pkg, sim, kernel_family, mod, data):
    ctx = pkg.ModContext(mod, engine=sim)
    lim = 2 ** (32 * ctx.limbs)
    a = data.draw(st.lists(st.integers(min_value=0, max_value=lim - 1), min_size=1, max_size=3))
    b = [data.draw(st.integers(min_value=0, max_value=lim - 1)) for _ in a]
    assert ctx.mulmod(a, b) == [x * y % mod for x, y in zip(a, b)]
    e = data.draw(st.integers(min_value=0, max_value=2 ** 96))
    assert ctx.powmod(a, e) == [pow(x, e, mod) for x in a]
    es = [data.draw(st.integers(min_value=0, max_value=2 ** 70)) for _ in a]
    assert ctx.powmod(a, es) == [pow(x, y, mod) for x, y in zip(a, es)]
    wide = [x * lim + y for x, y in zip(a, b)]                       # double-width bases are reduced on the device
    assert ctx.powmod(wide, 65537) == [pow(x, 65537, mod) for x in wide]
    inv, status = ctx.invert(a)
    for x, i, s in zip(a, inv, status):
        try:
            expected = pow(x, -1, mod)
        except ValueError:
            expected = None
        assert (s, i) == ((1, 0) if expected is None else (0, expected))
    ctx.close()


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(bits=st.integers(min_value=10, max_value=600), seed=st.integers(min_value=0, max_value=2 ** 32), data=st.data())
def test_paillier_ops_match_python(pkg, sim, kernel_family, bits, seed, data):
    import importlib
    import random
    util = importlib.import_module("python-paillier_b200.util")
    rng = random.Random(seed)                                # deterministic keys: hypothesis must be able to replay

    def prime(nbits):
        c = rng.randrange(1 << (nbits - 1), 1 << nbits) | 1
        while not util.is_prime(c):
            c += 2
        return c
    p, q = prime(bits // 2 + 1), prime(bits - bits // 2 + 1)
    if p == q:
        return
    n = p * q
    nsq = n * n
    pub, priv = pkg.PublicContext(n, engine=sim), pkg.PrivateContext(p, q, engine=sim)
    m = data.draw(st.lists(st.integers(min_value=0, max_value=n - 1), min_size=1, max_size=4))
    r = [data.draw(st.integers(min_value=1, max_value=n - 1)) for _ in m]
    c = pub.raw_encrypt(m, r)
    assert c == [(1 + n * x) * pow(y, n, nsq) % nsq for x, y in zip(m, r)]
    k = [data.draw(st.integers(min_value=0, max_value=n // 3 - 2)) for _ in m]
    out, status = pub.raw_mul(c, k)
    assert status == [0] * len(m) and out == [pow(x, y, nsq) for x, y in zip(c, k)]
    assert pub.raw_add(c, out) == [x * y % nsq for x, y in zip(c, out)]
    import math
    dec = priv.raw_decrypt(c)
    assert all(d == x for d, x, y in zip(dec, m, r) if math.gcd(y, n) == 1)
    pub.close(); priv.close()
