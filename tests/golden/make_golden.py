"""Generates tests/golden/golden.json: sha256 digests of the outputs of the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference by `make -C oracle ref`) on the seeded cases of cases.py.  Run here, commit the JSON; the digests
then pin the oracle on machines where the reference is not available (tests/test_golden.py)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from cases import build_cases, digest  # noqa: E402
from oracle_lib import Ref  # noqa: E402

ref = Ref()
out = {name: digest(ref_fn(ref)) for name, ref_fn, _ in build_cases()}
json.dump(out, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
print(f"wrote {len(out)} digests")
