"""Generates tests/golden/serialized_test_cube.b64: the reference's shared test fixture
(test/helpers/create-test-cube.js:3-58, with one cell unset) serialized by THIS repository's
restatement of the wire format (olap_in_memory_b200/serialization.py) on the Python oracle store.

These are NOT bytes written by Node (no JS engine in the image): they freeze the restatement so
that any later change to the format code is caught, and give a maintainer with Node one string
to feed to `Cube.deserializeFromBase64String` of the real package as a cross-check.

    python tests/golden/make_serialized.py        # rewrites serialized_test_cube.b64
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kats  # noqa: E402
from oracle.store_oracle import OracleStore  # noqa: E402


def fixture():
    cube = kats.create_test_cube(OracleStore)
    cube.setSingleData("antennas", {"location": "toledo", "period": "winter"}, 0)
    return cube


if __name__ == "__main__":
    with open(os.path.join(HERE, "serialized_test_cube.b64"), "w") as f:
        f.write(fixture().serializeToBase64String() + "\n")
