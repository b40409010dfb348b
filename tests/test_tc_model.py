"""The arithmetic of the tensor-core reduction path (python-paillier_b200/csrc/pai_tc.cuh) restated on Python integers
with explicit base-256 column sums -- the exactness argument of its header, checked against plain integer division:
   m = T_lo * N' mod R from the column sums of GEMM 1 (all < 2^24);
   floor(m*n / R) = floor(S / R) + [s > l]   with S = byte columns D-4 .. 2D-2 of m*n propagated exactly
                                             (GEMM 2 = columns D-4 .. 2D-5, three scalar columns on top),
                                             s = the guard limb of S, l = the top limb of L = -T_lo mod R.
No engine involved: this pins the algorithm, tests/test_tc_hostsim.py pins the implementation."""
import random

import pytest


def _bytes(x, d):
    return list(x.to_bytes(d, "little"))


def _int(b):
    return int.from_bytes(bytes(b), "little")


def _propagate(cols, carry=0):
    out = []
    for c in cols:
        s = c + carry
        out.append(s & 0xff)
        carry = s >> 8
    return out, carry


def _redc_hi(mb, nb, D, l_top):
    cols = [sum(mb[k] * nb[j - k] for k in range(D) if 0 <= j - k < D) for j in range(D - 4, 2 * D - 4)]
    assert all(c < 2 ** 24 for c in cols)
    e0 = mb[D - 3] * nb[D - 1] + mb[D - 2] * nb[D - 2] + mb[D - 1] * nb[D - 3]
    e1 = mb[D - 2] * nb[D - 1] + mb[D - 1] * nb[D - 2]
    e2 = mb[D - 1] * nb[D - 1]
    digs, carry = _propagate(cols + [e0, e1, e2])
    assert carry < 256
    digs.append(carry)
    s = _int(digs[:4])
    assert _int(digs[4 + D:]) == 0                        # the high half fits D digits
    return _int(digs[4:4 + D]) + (1 if s > l_top else 0)


@pytest.mark.parametrize("D,trials", [(16, 200), (32, 60), (64, 20), (256, 2)])
def test_byte_gemm_montgomery_reduction_is_exact(D, trials):
    rng = random.Random(D)
    R = 256 ** D
    for t in range(trials):
        n = rng.getrandbits(8 * D) | 1 | (1 << (8 * D - 1))
        if t % 3 == 0:
            n = R - 1 - 2 * rng.getrandbits(16)               # n close to R: largest carries
        Np = (-pow(n, -1, R)) % R
        for kind in range(4):
            T = [rng.randrange(n * n), (n - 1) * (n - 1), rng.randrange(n) * R, rng.randrange(n) * R + (R - rng.getrandbits(20) - 1)][kind]
            T_lo, T_hi = T % R, T // R
            tb, npb = _bytes(T_lo, D), _bytes(Np, D)
            cols = [sum(tb[k] * npb[j - k] for k in range(j + 1)) for j in range(D)]
            assert all(c < 2 ** 24 for c in cols)
            mb, _ = _propagate(cols)
            m = _int(mb)
            assert m == T_lo * Np % R
            L = (-T_lo) % R
            top, lower = T_lo >> (8 * (D - 4)), T_lo & (256 ** (D - 4) - 1)
            l_top = (~top + (1 if lower == 0 else 0)) & 0xffffffff       # what the kernel derives while it stores T_lo
            assert l_top == L >> (8 * (D - 4))
            hi = _redc_hi(mb, _bytes(n, D), D, l_top)
            assert hi == m * n // R
            t_red = T_hi + hi + (1 if T_lo else 0)
            assert t_red == (T + m * n) // R and t_red < 2 * n
