#!/usr/bin/env python3
"""Generate tests/golden/golden_jlmini_random_v1.json: the seeded random problems of
tests/test_oracle_jlmini_random.py, executed through the reference's own source (jlmini), committed so that
the GPU box (no reference tree there) can check the CUDA path against them as well.
TEST INFRASTRUCTURE ONLY.    python oracle/jlmini/gen_golden_random.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import gen_golden as G  # noqa: E402
import refsolve as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "golden_jlmini_random_v1.json")


def main():
    if not R.available():
        print("reference tree not present; nothing generated")
        return 1
    import test_oracle_jlmini_random as T
    res = [G.run(T._random_case(k)) for k in range(T.N_CASES)]
    doc = {"generator": "oracle/jlmini/gen_golden_random.py (cases: tests/test_oracle_jlmini_random.py::_random_case)",
           "reference": "SciML/SimpleDiffEq.jl v1.16.3 at /root/reference", "cases": res}
    with open(OUT, "w") as fh:
        json.dump(doc, fh, indent=0, separators=(",", ":"))
    print("wrote %s: %d cases, %.1f KB" % (OUT, len(res), os.path.getsize(OUT) / 1e3))
    return 0


if __name__ == "__main__":
    sys.exit(main())
