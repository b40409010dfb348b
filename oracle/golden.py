"""Fixture helpers (TEST INFRASTRUCTURE): load tests/golden/*.json generated from the reference."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def H(s):
    """hex string from the fixtures -> int (a leading '-' marks negative values)."""
    return -int(s[1:], 16) if s.startswith("-") else int(s, 16)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)
