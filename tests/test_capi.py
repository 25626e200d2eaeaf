"""CPU: the C-ABI library loads and exports every symbol include/airice_b200.h declares; without a GPU the entry
points fail loudly (there is no CPU path to fall back to)."""
import ctypes as C
import os
import re

import pytest

from conftest import ATMOSPHERE, ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "airice_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(airice_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from airiceraytracing_b200 import _capi
    assert declared_functions() == sorted(_capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    from airiceraytracing_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(_capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_library_contains_sm100a_code_only():
    import subprocess
    from airiceraytracing_b200 import _capi
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from airiceraytracing_b200 import AirIceError, AirIceSolver, _capi
    with pytest.raises(AirIceError):
        AirIceSolver(ATMOSPHERE)
    lib = _capi.load()
    h = C.c_void_p()
    rc = lib.airice_create(ATMOSPHERE.encode(), 0, 0, C.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.airice_last_error()


def test_product_does_not_reference_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/ or tests/hostsim."""
    pkg = os.path.join(ROOT, "airiceraytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "ORACLE_PATH_IN_COMMENT") or f == "build.py" or \
                    "import oracle" not in text and "from oracle" not in text, f
                assert "hostsim" not in text, f
                assert "liboracle" not in text and "libmultiray_ref" not in text, f
