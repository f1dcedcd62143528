"""The C-ABI library loads without a GPU and exports every symbol include/m3p2i_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from m3p2i_b200 import _abi as A
from m3p2i_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return B.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "m3p2i_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(m3p2i_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(libpath):
    lib = C.CDLL(libpath)
    declared = _declared_symbols()
    assert len(declared) >= 35
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/m3p2i_b200.h but not exported"
        assert sym in A.PROTOTYPES, f"{sym} has no ctypes prototype in _abi.py"
    assert set(A.PROTOTYPES) <= set(declared)


def test_struct_layouts_match(libpath):
    lib = C.CDLL(libpath)
    lib.m3p2i_abi_sizeof.argtypes = [C.c_char_p]
    for name, struct in A.STRUCTS.items():
        assert lib.m3p2i_abi_sizeof(name.encode()) == C.sizeof(struct), name
    assert lib.m3p2i_abi_sizeof(b"nope") == -1


def test_no_cpu_fallback(libpath):
    """Without a CUDA device creation must fail loudly with ERR_NO_DEVICE (there is no CPU path in the product)."""
    lib = C.CDLL(libpath)
    fn = A.bind(lib)
    if fn["m3p2i_device_count"]() > 0:
        pytest.skip("a GPU is visible")
    from m3p2i_b200 import scene as S
    cfg = S.build_config(S.make_cfg("point_env", "navigation", [1.0, 1.0], 32, 12))
    h = A.vp()
    rc = fn["m3p2i_create"](C.byref(cfg), 0, C.byref(h))
    assert rc == -5 and not h.value
    assert b"no CUDA device" in fn["m3p2i_last_error"]()


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "m3p2i-aip_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle_py" not in text and "liboracle" not in text and "orc_" not in text, os.path.join(dp, f)
