"""Loads tests/golden/store_vectors.json back into the case dicts of tests/cases.py."""
import json
import math
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "store_vectors.json")


def dec(values):
    table = {"nan": math.nan, "inf": math.inf, "-inf": -math.inf, "-0": -0.0}
    return np.asarray([table[v] if isinstance(v, str) else v for v in values], dtype=np.float64)


def load():
    out = []
    for entry in json.load(open(PATH))["vectors"]:
        case = dict(entry)
        expected = dec(case.pop("expected"))
        for k in ("data", "my_data", "his_data"):
            if k in case:
                case[k] = dec(case[k]).astype(np.float32)
        for k in ("default", "my_default", "his_default"):
            if k in case:
                case[k] = math.nan if case[k] == "nan" else 0.0
        for k in ("maps", "keep"):
            if k in case:
                case[k] = [np.asarray(m, dtype=np.int32) for m in case[k]]
        out.append((case, expected))
    return out
