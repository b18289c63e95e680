"""Generates tests/golden/store_vectors.json: lowered store cases with their expected
results, computed by the Python oracle (oracle/store_oracle.py, itself pinned on the
reference's own test expectations by tests/test_oracle_kats.py).  The reference cannot be
executed in this image (JavaScript, no engine), so these vectors are oracle outputs, not
reference outputs; they freeze the oracle's behaviour so that a later change to either
oracle or to the CUDA path is caught against committed data.

    python tests/golden/make_golden.py        # rewrites store_vectors.json
"""
import itertools
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle.store_oracle import OracleStore  # noqa: E402


def enc(values):
    """floats -> JSON-safe list (NaN/inf as strings, -0.0 kept)"""
    out = []
    for v in np.asarray(values, dtype=np.float64).tolist():
        if v != v:
            out.append("nan")
        elif math.isinf(v):
            out.append("inf" if v > 0 else "-inf")
        elif v == 0 and math.copysign(1, v) < 0:
            out.append("-0")
        else:
            out.append(v)
    return out


def jsonable(case):
    d = {}
    for k, v in case.items():
        if k in ("data", "my_data", "his_data"):
            d[k] = enc(v)
        elif k in ("default", "my_default", "his_default"):
            d[k] = "nan" if v != v else 0
        elif isinstance(v, np.ndarray):
            d[k] = v.tolist()
        elif isinstance(v, list):
            d[k] = [x.tolist() if isinstance(x, np.ndarray) else x for x in v]
        else:
            d[k] = v
    return d


def main():
    picked = []
    gens = [cases.drillup_cases(), cases.drilldown_cases(), cases.dice_cases(), cases.reorder_cases(), cases.load_cases()]
    for gen in gens:
        small = [c for c in gen if len(c.get("data", c.get("my_data"))) <= 400]
        picked += small[:: max(1, len(small) // 24)]
    vectors = []
    for case in picked:
        want = cases.run_case(case, OracleStore)
        entry = jsonable(case)
        entry["expected"] = enc(want)
        vectors.append(entry)
    path = os.path.join(HERE, "store_vectors.json")
    json.dump({"generator": "tests/golden/make_golden.py", "oracle": "oracle/store_oracle.py", "vectors": vectors},
              open(path, "w"), separators=(",", ":"))
    print(f"{len(vectors)} vectors -> {path} ({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    main()
