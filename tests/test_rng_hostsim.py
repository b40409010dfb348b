"""Device-side obfuscator generator (pai_rng.cuh) on the simulation engine: keystream equal to an independent
ChaCha20 (the `cryptography` package), range, determinism."""
import numpy as np
import pytest

from oracle.golden import H, load_golden


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    return pkg.Engine(ge.build_hostsim())


def _keystream(seed, nonce, counter, nbytes):
    algorithms = pytest.importorskip("cryptography.hazmat.primitives.ciphers.algorithms")
    from cryptography.hazmat.primitives.ciphers import Cipher
    # the library's 16-byte "nonce" is the last four state words: 64-bit counter, 64-bit nonce (little endian)
    iv = counter.to_bytes(8, "little") + nonce.to_bytes(8, "little")
    enc = Cipher(algorithms.ChaCha20(seed, iv), mode=None).encryptor()
    return enc.update(b"\x00" * nbytes)


def test_keystream_range_and_determinism(pkg, sim):
    seed = bytes(range(32))
    # n with all bits set in its top limb: the first attempt is accepted (almost surely), so r IS the keystream
    n = 2 ** 512 - 569
    pub = pkg.PublicContext(n, engine=sim)
    out = np.zeros((5, pub.n_limbs), dtype=np.uint32)
    pub.random_lt_n_dev(out, 5, seed=seed, nonce=0x1122334455667788)
    for g in range(5):
        ks = _keystream(seed, 0x1122334455667788, g << 12, 64)
        assert out[g, :16].tobytes() == ks and not out[g, 16:].any()
    vals = pkg.limbs_to_ints(out)
    assert all(1 <= v < n for v in vals) and len(set(vals)) == 5
    out2 = np.zeros_like(out)
    pub.random_lt_n_dev(out2, 5, seed=seed, nonce=0x1122334455667788)
    assert (out == out2).all()
    pub.random_lt_n_dev(out2, 5, seed=seed, nonce=1)
    assert not (out == out2).any(axis=1).all()
    # a modulus just above a power of two: about half of the attempts are rejected, results stay in range
    fx = load_golden("vectors_256.json")
    pub = pkg.PublicContext(H(fx["n"]), engine=sim)
    out = np.zeros((400, pub.n_limbs), dtype=np.uint32)
    pub.random_lt_n_dev(out, 400)
    vals = pkg.limbs_to_ints(out)
    assert all(1 <= v < pub.n for v in vals) and len(set(vals)) == 400
    mean = sum(v / pub.n for v in vals) / 400                # uniform on [1, n): mean 1/2, sigma 0.0144
    assert 0.43 < mean < 0.57
    with pytest.raises(ValueError):
        pub.random_lt_n_dev(out, 4, seed=b"short")


def test_vector_encrypt_uses_device_rng(pkg, sim):
    import importlib
    engine_mod = importlib.import_module("python-paillier_b200.engine")
    engine_mod._set_engine_for_tests(sim)
    try:
        fx = load_golden("vectors_256.json")
        pk = pkg.PaillierPublicKey(H(fx["n"]))
        sk = pkg.PaillierPrivateKey(pk, H(fx["p"]), H(fx["q"]))
        v1, v2 = pk.encrypt_batch([1.5, -2.0, 3.0]), pk.encrypt_batch([1.5, -2.0, 3.0])
        assert v1.ciphertexts(False) != v2.ciphertexts(False)          # fresh r every time
        assert sk.decrypt_batch(v1) == sk.decrypt_batch(v2) == [1.5, -2.0, 3.0]
    finally:
        engine_mod._set_engine_for_tests(None)
