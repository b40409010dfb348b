"""Fixed key material for benchmarks and examples: the JSON fixtures under tests/golden (generated from the unmodified
reference by tests/golden/make_golden.py).  A fixed key keeps measurements comparable between runs; nothing here is
used by the engine itself."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def H(s):
    """hex string from the fixtures -> int (a leading '-' marks negative values)."""
    return -int(s[1:], 16) if s.startswith("-") else int(s, 16)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def fixed_key(key_bits):
    """(n, p, q) of the committed key of that size."""
    fx = load_golden("vectors_%d.json" % key_bits)
    return H(fx["n"]), H(fx["p"]), H(fx["q"])
