"""CPU-side checks of the drop-in boundary: libolapgpu.so loads without a GPU, exports
every symbol include/olap_gpu.h declares, reports argument errors with the reference's
texts, and fails loudly (no CPU fallback) when a store is requested without a device."""
import os
import re

import pytest

from olap_in_memory_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "olap_gpu.h")

pytestmark = pytest.mark.skipif(not os.path.exists(_native.LIB_PATH), reason="libolapgpu.so not built (run build())")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(olap_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load_library()
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in olap_gpu.h but not exported"
    assert sorted(_native.SIGNATURES) == names, "ctypes signature table out of sync with the header"


def test_abi_version_and_method_names():
    import ctypes as C

    lib = _native.load_library()
    assert lib.olap_abi_version() == 1
    code = C.c_int(-1)
    for k, name in enumerate(["sum", "average", "highest", "lowest", "first", "last", "product"]):
        assert lib.olap_method_from_name(name.encode(), C.byref(code)) == 0 and code.value == k
    assert lib.olap_method_from_name(b"median", C.byref(code)) == _native.E_INVALID
    # in-memory.js:294-296
    assert lib.olap_last_error().decode() == "Unsupported aggregation method: median"


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from olap_in_memory_b200 import Cube, GenericDimension

    cube = Cube([GenericDimension("d", "root", ["a", "b"])])
    with pytest.raises(_native.OlapError, match="no CUDA device|CUDA"):
        cube.createStoredMeasure("m1")


def test_host_side_argument_errors_use_reference_texts():
    from olap_in_memory_b200.store import _default_kind, _method_code

    with pytest.raises(ValueError, match="Invalid default value, only NaN and 0 are supported"):
        _default_kind(1)
    with pytest.raises(ValueError, match="Unsupported aggregation method: median"):
        _method_code("median")


def test_napi_shim_type_checks_against_the_header():
    """addon/olap_napi.cc cannot be built here (no Node.js headers), but it can be type-checked:
    a stub node_api.h with the published N-API signatures (addon/stub/) lets g++ verify every
    olap_* call of the shim against include/olap_gpu.h, and that every `native.<name>` the JS
    facade uses is a property the shim defines."""
    import re
    import shutil
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    proc = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "addon", "stub"),
                           os.path.join(root, "addon", "olap_napi.cc")], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-3000:]
    shim = open(os.path.join(root, "addon", "olap_napi.cc")).read()
    defined = set(re.findall(r'\{"(\w+)", nullptr, \w+, nullptr', shim))
    used = set(re.findall(r"native\.(\w+)", open(os.path.join(root, "js", "gpu-store.js")).read()))
    assert used <= defined, sorted(used - defined)
