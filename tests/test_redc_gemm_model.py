"""The GEMM formulation of Montgomery reduction planned for the tensor-core kernel (DESIGN.md section 8,
bench_micro/redc_gemm_model.py) reproduces REDC on Python integers, and its int32 column sums stay in range."""
import importlib.util
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("redc_gemm_model", os.path.join(ROOT, "bench_micro", "redc_gemm_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("bits", [64, 256, 1024, 2048])
def test_redc_by_gemm_matches_integers(bits):
    mdl = _model()
    rng = random.Random(bits)
    D = bits // 8
    R = 1 << bits
    N = rng.getrandbits(bits) | (1 << (bits - 1)) | 1
    ts = [0, N * R - 1, R - 1, R, (N - 1) * (N - 1)] + [rng.randrange(N * R) for _ in range(60)] + [rng.randrange(R) * R for _ in range(4)] + [rng.randrange(R) for _ in range(4)]
    ts += [255 * sum(1 << (8 * i) for i in range(2 * D)) % (N * R)]           # all digits 255: the column-sum worst case
    out, peak = mdl.redc_gemm(ts, N, D)
    Rinv = pow(R, -1, N)
    assert [u for u, _ in out] == [t * Rinv % N for t in ts]
    assert peak < 2 ** 31 and peak <= D * 255 * 255
    # 4 guard columns below column D plus the known low half give the exact high half of m*N in every row
    assert all(same for _, same in out)
