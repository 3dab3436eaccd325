"""Oracle vs committed golden digests (tests/golden/golden.json, produced by the compiled reference with
tests/golden/make_golden.py).  Needs neither /root/reference nor a GPU."""
import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from cases import build_cases, digest  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))
CASES = build_cases()


def test_golden_covers_all_cases():
    assert sorted(GOLDEN) == sorted(name for name, _, _ in CASES)


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_oracle_matches_reference_digest(oracle, name):
    fn = next(o for n, _, o in CASES if n == name)
    assert digest(fn(oracle)) == GOLDEN[name]
