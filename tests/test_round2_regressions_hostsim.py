"""Regression tests for the round-1 review findings, on the test-only host simulation:
  * one context shared by several host threads (the C ABI locks a context for the duration of a call);
  * EncryptedVector with an unbalanced prime pair (private rows wider than public rows);
  * numpy integer / array indexing of EncryptedVector; raw_mul status in decrease_exponent_to;
  * raw_encrypt_batch applies the scalar path's normalisation of r (phe/paillier.py:136-137);
  * the chunked Python-int pipeline of raw_encrypt / raw_decrypt gives the same list as one call.
"""
import importlib
import random
import threading

import numpy as np
import pytest

from oracle import paillier_oracle as orc
from oracle.golden import H, load_golden


@pytest.fixture(scope="module")
def env(pkg):
    import __graft_entry__ as ge
    engine_mod = importlib.import_module("python-paillier_b200.engine")
    engine_mod._set_engine_for_tests(pkg.Engine(ge.build_hostsim()))
    orc.BACKEND = "gmp" if orc.have_gmp() else "python"
    fx = load_golden("vectors_256.json")
    pk = pkg.PaillierPublicKey(H(fx["n"]))
    sk = pkg.PaillierPrivateKey(pk, H(fx["p"]), H(fx["q"]))
    yield pk, sk, engine_mod
    engine_mod._set_engine_for_tests(None)
    orc.BACKEND = "python"


def test_one_context_many_threads(pkg, env):
    """Reproduces the advisor's finding (4 threads on one PaillierPublicKey gave wrong ciphertexts): every thread's
    results must equal the oracle's."""
    pk, sk, _ = env
    opub = orc.PublicConsts(pk.n)
    util = importlib.import_module("python-paillier_b200.util")
    errors = []

    def work(seed):
        try:
            rng = random.Random(seed)
            for _ in range(25):
                m, r = rng.randrange(pk.n), rng.randrange(1, pk.n)
                c = pk.raw_encrypt(m, r)
                if c != orc.raw_encrypt(opub, m, r):
                    errors.append(("enc", seed))
                if sk.raw_decrypt(c) != m:
                    errors.append(("dec", seed))
                a, b = rng.randrange(pk.nsquare), rng.randrange(pk.nsquare)
                if util.mulmod(a | (1 << 1001), b, pk.nsquare) != (a | (1 << 1001)) * b % pk.nsquare:
                    errors.append(("mulmod", seed))
                k = rng.getrandbits(40)
                if (pkg.EncryptedNumber(pk, c) * k).ciphertext(False) != orc.raw_mul(opub, c, k):
                    errors.append(("mul", seed))
            ms = [rng.randrange(pk.n) for _ in range(7)]
            rs = [rng.randrange(1, pk.n) for _ in range(7)]
            if pk.raw_encrypt_batch(ms, rs) != [orc.raw_encrypt(opub, m, r) for m, r in zip(ms, rs)]:
                errors.append(("batch", seed))
            if sk.raw_decrypt_batch([orc.raw_encrypt(opub, m, r) for m, r in zip(ms, rs)]) != ms:
                errors.append(("dbatch", seed))
        except Exception as e:        # noqa: BLE001
            errors.append((repr(e), seed))
    threads = [threading.Thread(target=work, args=(s,)) for s in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]


def test_ctx_cache_eviction_keeps_live_contexts(pkg, env):
    util = importlib.import_module("python-paillier_b200.util")
    rng = random.Random(3)
    mods = [(rng.getrandbits(1100) | (1 << 1099) | 1) for _ in range(util._CTX_CACHE_SIZE + 4)]
    first = util._mod_ctx(mods[0])
    for m in mods[1:]:
        util._mod_ctx(m)
    assert mods[0] not in util._ctx_cache and first.h             # evicted from the cache, not destroyed under us
    a, b = rng.getrandbits(1090), rng.getrandbits(1090)
    assert first.mulmod([a], [b]) == [a * b % mods[0]]


def test_vector_with_unbalanced_primes(pkg, env):
    """p of 100 bits and q of 412 bits: public rows are 16/32 limbs, private rows 32/64 (advisor finding)."""
    util = importlib.import_module("python-paillier_b200.util")
    rng = random.Random(11)

    def prime_bits(b):
        while True:
            c = rng.getrandbits(b) | (1 << (b - 1)) | 1
            if util.is_prime(c):
                return c
    p, q = prime_bits(100), prime_bits(412)
    pk = pkg.PaillierPublicKey(p * q)
    sk = pkg.PaillierPrivateKey(pk, p, q)
    assert sk.engine_context().c_limbs != pk.engine_context().c_limbs
    vals = [0.5, -1.25, 3.0, 12345678, -7]
    v = pk.encrypt_batch(vals)
    assert sk.decrypt_batch(v) == vals
    assert [e.decode() for e in v.decrypt_encoded(sk)] == vals
    assert sk.decrypt(v[2]) == 3.0


def test_vector_indexing(pkg, env):
    pk, sk, _ = env
    vals = [1.5, -2.0, 3.25, 4.0, -5.5, 6.0]
    v = pk.encrypt_batch(vals)
    assert sk.decrypt(v[np.int64(3)]) == 4.0 and sk.decrypt(v[-1]) == 6.0 and sk.decrypt(v[np.int32(0)]) == 1.5
    with pytest.raises(IndexError):
        v[6]
    assert sk.decrypt_batch(v[1:4]) == vals[1:4]
    assert sk.decrypt_batch(v[np.array([4, 0, 2])]) == [vals[4], vals[0], vals[2]]
    w = v[np.array([True, False, True, False, False, True])]
    assert len(w) == 3 and sk.decrypt_batch(w) == [vals[0], vals[2], vals[5]]


def test_raw_encrypt_batch_normalises_r(pkg, env):
    pk, sk, _ = env
    opub = orc.PublicConsts(pk.n)
    n, nsq = pk.n, pk.nsquare
    rng = random.Random(5)
    ms = [rng.randrange(n) for _ in range(6)]
    wide = nsq - 5                                   # legal for the reference (powmod reduces), wider than an engine row
    rs = [rng.randrange(1, n), 1, nsq + 7, wide, n + 3, None]
    out = pk.raw_encrypt_batch(ms, rs)
    for m, r, c in zip(ms[:5], rs[:5], out[:5]):
        assert c == orc.raw_encrypt(opub, m, r % nsq if r >= nsq else r)
        assert c == pk.raw_encrypt(m, r)
    assert sk.raw_decrypt(out[5]) == ms[5]           # falsy r: a fresh obfuscator was drawn
    with pytest.raises(ValueError):
        pk.raw_encrypt_batch(ms, rs[:3])
    with pytest.raises(TypeError):
        pk.raw_encrypt_batch([1.5], [3])


def test_python_int_pipeline_equals_single_call(pkg, env, monkeypatch):
    pk, sk, engine_mod = env
    ctx, pctx = pk.engine_context(), sk.engine_context()
    assert ctx.wave() >= 1 and pctx.wave() >= 1
    rng = random.Random(8)
    ms = [rng.randrange(-5, pk.n + 5) for _ in range(23)]
    rs = [rng.randrange(1, pk.n) for _ in range(23)]
    whole = ctx.raw_encrypt(ms, rs)
    monkeypatch.setattr(engine_mod, "_PIPE_MIN", 4)
    monkeypatch.setattr(engine_mod, "_chunk_ranges", lambda count, wave, target=0: [(lo, min(count, lo + 5)) for lo in range(0, count, 5)])
    assert ctx.raw_encrypt(ms, rs) == whole
    assert pctx.raw_decrypt(whole) == [m % pk.n for m in ms]
    assert engine_mod.limbs_to_ints(engine_mod.ints_to_limbs([], 4)) == []
    with pytest.raises(ValueError):
        engine_mod.ints_to_limbs([1 << 128], 4)
    with pytest.raises(ValueError):
        engine_mod.ints_to_limbs([-1], 4)
