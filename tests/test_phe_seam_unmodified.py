"""The boundary exercised the way the reference would bind it: the UNMODIFIED /root/reference/phe package with its three
bigint seam functions (phe/util.py:38,53,85, imported by name at phe/paillier.py:29) rebound to
integration/phe_b200_backend.py -- the ctypes stub INTEGRATION.md section 1 quotes -- exactly as the reference's own tests
flip backends (phe/tests/util_test.py:64-75).  The reference's PaillierTestRawEncryption and PaillierTestEncryptedNumber
classes (phe/tests/paillier_test.py:106-164, 430-1058) and its util tests then run on the reference's own classes.
Build container only (needs /root/reference); the engine is the test-only host simulation of the device code."""
import importlib
import importlib.util
import os
import sys
import unittest

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "phe", "tests")), reason="reference tree not present")


@pytest.fixture(scope="module")
def real_phe():
    import __graft_entry__ as ge
    lib = ge.build_hostsim()
    saved_mods = {k: v for k, v in sys.modules.items() if k == "phe" or k.startswith("phe.")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        phe = importlib.import_module("phe")
        assert os.path.realpath(phe.__file__).startswith(REF), "must be the reference's own package"
        spec = importlib.util.spec_from_file_location("phe_b200_backend", os.path.join(ge.ROOT, "integration", "phe_b200_backend.py"))
        backend = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(backend)
        backend.install(phe, lib_path=lib)
        yield phe, backend
        backend.uninstall(phe)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "phe" or k.startswith("phe.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def _classes(mod, names):
    suite = unittest.TestSuite()
    for n in names:
        suite.addTests(unittest.defaultTestLoader.loadTestsFromTestCase(getattr(mod, n)))
    return suite


def test_seam_is_rebound_and_counts_calls(real_phe):
    phe, backend = real_phe
    import phe.paillier as pp
    import phe.util as pu
    assert pp.powmod is backend.powmod and pp.mulmod is backend.mulmod and pp.invert is backend.invert
    assert pu.powmod is backend.powmod
    # the reference's known answer (phe/tests/paillier_test.py:128-136) through the reference's own class
    pk = pp.PaillierPublicKey(126869)
    sk = pp.PaillierPrivateKey(pk, 293, 433)
    assert pk.raw_encrypt(10100, 74384) == 935906717 and sk.raw_decrypt(935906717) == 10100
    # a real-size key: every powmod of encrypt / decrypt goes through the engine (counted by wrapping the backend)
    calls = {"n": 0}
    orig = backend._lib.pai_mod_powmod_host

    class Counting:
        def __call__(self, *a):
            calls["n"] += 1
            return orig(*a)
    backend._lib_saved = backend._lib

    class LibProxy:
        def __getattr__(self, name):
            return Counting() if name == "pai_mod_powmod_host" else getattr(backend._lib_saved, name)
    backend._lib = LibProxy()
    try:
        pk2, sk2 = pp.generate_paillier_keypair(n_length=1024)
        c = pk2.encrypt(-123456.75)
        assert sk2.decrypt(c + 0.25) == -123456.5
        assert calls["n"] >= 5          # hp, hq (key constants), r^n, and the CRT pair
    finally:
        backend._lib = backend._lib_saved


def test_reference_test_classes_on_unmodified_phe(real_phe, monkeypatch):
    phe, backend = real_phe
    import phe.paillier as pp
    orig = pp.generate_paillier_keypair
    # the simulation is ~100x slower than the GPU: smaller default keys, nothing else changes
    monkeypatch.setattr(pp, "generate_paillier_keypair",
                        lambda private_keyring=None, n_length=None: orig(private_keyring, n_length=n_length or 1152))
    monkeypatch.setattr(phe, "generate_paillier_keypair", pp.generate_paillier_keypair, raising=False)
    spec = importlib.util.spec_from_file_location("ref_paillier_test_real", os.path.join(REF, "phe", "tests", "paillier_test.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert os.path.realpath(mod.paillier.__file__).startswith(REF)
    res = unittest.TextTestRunner(verbosity=0).run(_classes(mod, ["PaillierTestRawEncryption", "PaillierTestEncryptedNumber",
                                                                 "TestKeyring", "TestIssue62"]))
    assert res.testsRun >= 80
    assert not res.failures and not res.errors, (res.failures[:2], res.errors[:2])
    spec = importlib.util.spec_from_file_location("ref_util_test_real", os.path.join(REF, "phe", "tests", "util_test.py"))
    umod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(umod)
    res = unittest.TextTestRunner(verbosity=0).run(_classes(umod, ["PaillierUtilTest"]))
    assert res.testsRun >= 5 and not res.failures and not res.errors, (res.failures[:2], res.errors[:2])
    spec = importlib.util.spec_from_file_location("ref_math_test_real", os.path.join(REF, "phe", "tests", "math_test.py"))
    mmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mmod)
    res = unittest.TextTestRunner(verbosity=0).run(_classes(mmod, ["ArithmeticTest"]))      # np.mean / np.dot idioms
    assert res.testsRun >= 2 and not res.failures and not res.errors, (res.failures[:2], res.errors[:2])
