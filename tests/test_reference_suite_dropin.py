"""Run the REFERENCE'S OWN unit tests (phe/tests/paillier_test.py, util_test.py, math_test.py) against the
drop-in package, with `phe` aliased to python-paillier_b200 and the kernels on the test-only host simulation.
Only possible where /root/reference exists (the build container); skipped elsewhere.  Key sizes are reduced
(the simulation is ~100x slower than the GPU), nothing else is changed."""
import importlib
import os
import sys
import unittest

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "phe", "tests")), reason="reference tree not present")


@pytest.fixture(scope="module")
def aliased(pkg):
    import __graft_entry__ as ge
    engine_mod = importlib.import_module("python-paillier_b200.engine")
    engine_mod._set_engine_for_tests(pkg.Engine(ge.build_hostsim()))
    saved = {k: sys.modules.get(k) for k in ("phe", "phe.paillier", "phe.util", "phe.encoding")}
    sys.modules["phe"] = pkg
    sys.modules["phe.paillier"] = importlib.import_module("python-paillier_b200.paillier")
    sys.modules["phe.util"] = importlib.import_module("python-paillier_b200.util")
    sys.modules["phe.encoding"] = importlib.import_module("python-paillier_b200.encoding")
    pkg.paillier = sys.modules["phe.paillier"]
    pkg.encoding = sys.modules["phe.encoding"]
    pmod = sys.modules["phe.paillier"]
    orig = pmod.generate_paillier_keypair

    def small_keys(private_keyring=None, n_length=None):
        return orig(private_keyring, n_length=n_length or int(os.environ.get("PAI_REFTEST_KEYBITS", "1152")))
    pmod.generate_paillier_keypair = small_keys
    pkg.generate_paillier_keypair = small_keys
    yield pkg
    pmod.generate_paillier_keypair = orig
    pkg.generate_paillier_keypair = orig
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    engine_mod._set_engine_for_tests(None)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(mod, skip=()):
    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    for t in loader.loadTestsFromModule(mod):
        for case in t:
            if not any(s in case.id() for s in skip):
                suite.addTest(case)
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    return res


@pytest.fixture(params=["thread-per-ciphertext", "warp-per-ciphertext"])
def kernel_family(request, monkeypatch):
    """The scalar phe API is a batch of one: on the GPU it runs on the warp-per-ciphertext kernels (pai_coop.cuh);
    PAI_COOP_MAX=0 forces the throughput kernels.  The reference's tests must pass on both."""
    monkeypatch.setenv("PAI_COOP_MAX", "0" if request.param.startswith("thread") else "1000000")


def test_reference_paillier_tests(aliased, kernel_family):
    mod = _load(os.path.join(REF, "phe", "tests", "paillier_test.py"), "ref_paillier_test")
    # skipped: key-generation sweeps up to 4096 bits / 100 keys (out of the hot-path scope, hours in simulation)
    res = _run(mod, skip=("testKeyUniqueness", "testDefaultKeySize", "testStaticPrivateKeySize"))
    assert res.testsRun > 150
    assert not res.failures and not res.errors, (res.failures[:2], res.errors[:2])


def test_reference_util_and_math_tests(aliased, kernel_family):
    mod = _load(os.path.join(REF, "phe", "tests", "util_test.py"), "ref_util_test")
    res = _run(mod, skip=("Fallbacks",))        # the fallback class toggles phe.util.HAVE_GMP / HAVE_CRYPTO internals
    assert res.testsRun >= 5 and not res.failures and not res.errors, (res.failures[:2], res.errors[:2])
    mod = _load(os.path.join(REF, "phe", "tests", "math_test.py"), "ref_math_test")
    res = _run(mod)
    assert res.testsRun >= 2 and not res.failures and not res.errors, (res.failures[:2], res.errors[:2])
