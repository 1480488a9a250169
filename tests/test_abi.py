"""CPU tier: the C-ABI shared library loads and exports every symbol include/mmpc.h declares; the
ctypes mirror has the same struct layout; compute entry points refuse to run without a GPU
(no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    _lib.build_library()
    return _lib.lib()


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmpc_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(L):
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(_lib.EXPORTS)


def test_struct_layout_and_defaults(L):
    a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
    assert L.mmpc_struct_sizes(C.byref(a), C.byref(b), C.byref(c)) == 0
    assert (a.value, b.value, c.value) == (C.sizeof(_abi.MmpcConfig), C.sizeof(_abi.MmpcBatchIn), C.sizeof(_abi.MmpcBatchOut))
    cfg = _abi.MmpcConfig()
    L.mmpc_default_config(C.byref(cfg))
    py = _abi.default_config()
    assert bytes(cfg) == bytes(py)        # C defaults == Python defaults, field for field
    # reference defaults, controllers/mpc_wholebody_qref.py:11-22
    assert list(cfg.Qd) == [25, 25, 0, 0, 0, 5, 5, 5, 5] and cfg.S == 1e5 and cfg.max_iter == 2000
    assert cfg.xlim[0][2] == -np.inf and cfg.dulim[1][0] == np.inf and cfg.ulim[1][1] == np.pi


def test_version_and_error_strings(L):
    assert L.mmpc_version() >= 100
    assert b"CPU fallback" in L.mmpc_error_string(_abi.ERR_NO_DEVICE)


def test_no_cpu_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    cfg = _abi.default_config()
    assert L.mmpc_create(C.byref(cfg), 4, 0, C.byref(h)) == _abi.ERR_NO_DEVICE
    from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
    with pytest.raises(_lib.MmpcError):
        BatchSolver(B_max=1)
    t = C.c_double()
    assert L.mmpc_bench_fp64(0, C.byref(t)) == _abi.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mobile_manipulator_mpc_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("oracle/mmpc_oracle.c", "").replace("oracle/", "").lower(), f
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "refshim" not in txt and "import casadi" not in txt, f
