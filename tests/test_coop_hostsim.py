"""Warp-per-ciphertext path (pai_coop.cuh, taken for small batches) on the simulation engine: same results as the
golden vectors / the oracle as the thread-per-ciphertext kernels, for every operand width K the layout supports."""
import importlib
import random

import numpy as np
import pytest

from oracle import paillier_oracle as orc
from oracle.golden import H, load_golden


@pytest.fixture(scope="module")
def sim(pkg):
    import __graft_entry__ as ge
    orc.BACKEND = "gmp" if orc.have_gmp() else "python"
    yield pkg.Engine(ge.build_hostsim())
    orc.BACKEND = "python"


@pytest.mark.parametrize("kb", [64, 256, 512, 1024, 2048])
def test_golden_vectors_through_the_warp_path(pkg, sim, monkeypatch, kb):
    monkeypatch.setenv("PAI_COOP_MAX", "100000")
    fx = load_golden("vectors_%d.json" % kb)
    n, p, q = H(fx["n"]), H(fx["p"]), H(fx["q"])
    pub, priv = pkg.PublicContext(n, engine=sim), pkg.PrivateContext(p, q, engine=sim)
    launches = sim.lib.pai_launch_count()
    enc = fx["encrypt"][:6] if kb >= 1024 else fx["encrypt"]
    cs = pub.raw_encrypt([H(e["m"]) for e in enc], [H(e["r"]) for e in enc])
    assert cs == [H(e["c"]) for e in enc]
    assert priv.raw_decrypt(cs) == [H(e["d"]) for e in enc]
    dec = fx["decrypt_any"][:6] if kb >= 1024 else fx["decrypt_any"]
    assert priv.raw_decrypt([H(d["c"]) for d in dec]) == [H(d["d"]) for d in dec]
    assert sim.lib.pai_launch_count() > launches
    # the same calls with the path disabled give the same answers (and both really ran different kernels)
    monkeypatch.setenv("PAI_COOP_MAX", "0")
    assert pub.raw_encrypt([H(e["m"]) for e in enc[:3]], [H(e["r"]) for e in enc[:3]]) == cs[:3]


@pytest.mark.parametrize("bits,pbits", [(255, 127), (258, 100), (513, 256), (600, 64), (769, 384), (1030, 257), (1152, 576),
                                        (1500, 750), (3072, 1536), (4090, 2044)])
def test_edge_keys_through_the_warp_path(pkg, sim, monkeypatch, bits, pbits):
    monkeypatch.setenv("PAI_COOP_MAX", "100000")
    util = importlib.import_module("python-paillier_b200.util")
    rng = random.Random(bits * 7919 + pbits)

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
    big = bits > 2000
    ms = [0, n - 1] + ([] if big else [1, n // 2, rng.randrange(n)])
    rs = [1, n - 1] + ([] if big else [rng.randrange(1, n) for _ in range(3)])
    ms.append((1 << (32 * pub.n_limbs)) - 1)            # unreduced inputs are accepted and reduced
    rs.append((1 << (32 * pub.n_limbs)) - 1)
    cs = pub.raw_encrypt(ms, rs)
    assert cs == [orc.raw_encrypt(opub, m % n, r) for m, r in zip(ms, rs)]
    xs = cs[:3] + [0, 1, n, n * n - 1, p, q * q] + ([] if big else [p * p, q, (1 << (32 * pub.c_limbs)) - 1])
    assert priv.raw_decrypt(xs) == [orc.raw_decrypt(opriv, c) for c in xs]
    pub.close(); priv.close()


def test_generic_powmod_through_the_warp_path(pkg, sim, monkeypatch):
    monkeypatch.setenv("PAI_COOP_MAX", "100000")
    rng = random.Random(5)
    for bits in (33, 250, 700, 1100, 2100):
        N = rng.getrandbits(bits) | (1 << (bits - 1)) | 1
        ctx = pkg.ModContext(N, engine=sim)
        assert ctx.powmod([N - 2], 77) == [pow(N - 2, 77, N)]           # first call on a fresh context, batch of one
        for e in (0, 1, 2, 15, 16, 65537, rng.getrandbits(bits), N - 1):
            bases = [0, 1, N - 1, rng.randrange(N)]
            assert ctx.powmod(bases, e) == [pow(b, e, N) for b in bases]
        wide = [N * N - 1, (1 << (64 * ctx.limbs)) - 1, N + 1]          # double-width bases are reduced first
        assert ctx.powmod(wide, 65537) == [pow(b, 65537, N) for b in wide]
        ctx.close()


def test_tail_of_a_batch_goes_to_the_warp_path(pkg, sim, monkeypatch):
    """Whole waves on the throughput kernel, the remainder on the warp kernels, one call (the simulation build has
    waves of a few ciphertexts, pai_pub_wave): results are in order and equal to the all-throughput run."""
    fx = load_golden("vectors_256.json")
    n, p, q = H(fx["n"]), H(fx["p"]), H(fx["q"])
    rng = random.Random(3)
    probe = pkg.PublicContext(n, engine=sim)
    wave = probe.wave()
    probe.close()
    for batch in (wave + 1, wave + 3, 2 * wave, 2 * wave + 3):
        ms = [rng.randrange(n) for _ in range(batch)]
        rs = [rng.randrange(1, n) for _ in range(batch)]
        outs = []
        for lim in ("0", "3"):
            monkeypatch.setenv("PAI_COOP_MAX", lim)
            pub, priv = pkg.PublicContext(n, engine=sim), pkg.PrivateContext(p, q, engine=sim)
            before = sim.lib.pai_launch_count()
            cs = pub.raw_encrypt(ms, rs)
            launches = sim.lib.pai_launch_count() - before
            outs.append((cs, priv.raw_decrypt(cs), launches))
            pub.close(); priv.close()
        assert outs[0][:2] == outs[1][:2] and outs[0][1] == ms
        # 0 < batch % wave <= 3 -> two launches more than the plain run (constants of the warp layout) or at least one
        assert (outs[1][2] > outs[0][2]) == (batch % wave != 0)
