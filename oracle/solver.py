"""ctypes front-end of the CPU oracle solver (oracle/mmpc_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mmpc_oracle.c.  PARITY UNPINNED (no CasADi/IPOPT
in this image): the oracle is pinned by oracle/model.py, finite differences and SciPy solves.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from mobile_manipulator_mpc_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libmmpc_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.mmpc_oracle_solve.argtypes = [C.POINTER(_abi.MmpcConfig), C.c_int32, C.POINTER(_abi.MmpcBatchIn),
                                           C.POINTER(_abi.MmpcBatchOut), C.c_int32]
        _LIB.mmpc_oracle_solve.restype = C.c_int
    return _LIB


def config_from_batch(batch, mode=_abi.MODE_REFERENCE, **over):
    cfg = _abi.default_config(N=batch["N"], dt=batch["dt"], n_obs=batch["n_obs"], n_pl=batch["n_pl"], mode=mode)
    cfg.obs_per_stage = int(batch.get("obs_per_stage", 0))
    if "Qd" in batch:
        cfg.Qd[:] = list(batch["Qd"]); cfg.Pd[:] = list(batch.get("Pd", batch["Qd"]))
    if "Rd" in batch:
        cfg.Rd[:] = list(batch["Rd"])
    if "Wd" in batch:
        cfg.Wd[:] = list(batch["Wd"])
    if "S" in batch:
        cfg.S = float(batch["S"])
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


def _solve_chunk(cfg, batch, lo, hi, u_guess):
    B = hi - lo
    N = cfg.N
    f = lambda a: None if a is None else np.ascontiguousarray(a[lo:hi], dtype=np.float64)
    arrs = dict(x_init=f(batch["x_init"]), x_ref=f(batch["x_ref"]), u_ref=f(batch["u_ref"]), u_last=f(batch["u_last"]),
                u_guess=f(u_guess), circles=f(batch.get("circles")), planes=f(batch.get("planes")))
    npl = batch.get("n_pl_inst")
    npl = None if npl is None else np.ascontiguousarray(npl[lo:hi], dtype=np.int32)
    flags = batch.get("flags")
    flags = None if flags is None else np.ascontiguousarray(flags[lo:hi], dtype=np.uint8)
    xg = f(batch.get("x_guess"))
    bi = _abi.MmpcBatchIn(*[_abi.ptr(arrs[k]) for k in ("x_init", "x_ref", "u_ref", "u_last", "u_guess", "circles", "planes")],
                          _abi.ptr(npl), _abi.ptr(flags), _abi.ptr(xg))
    out = dict(U=np.zeros((B, N, 5)), X=np.zeros((B, N + 1, 9)), s=np.zeros((B, N + 1)), cost=np.zeros(B),
               kkt=np.zeros(B), iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32))
    bo = _abi.MmpcBatchOut(*[_abi.ptr(out[k]) for k in ("U", "X", "s", "cost", "kkt", "iters", "status")])
    rc = lib().mmpc_oracle_solve(C.byref(cfg), B, C.byref(bi), C.byref(bo), 1)
    assert rc == 0
    return out


def solve(batch, cfg=None, mode=_abi.MODE_REFERENCE, threads=1, u_guess=None):
    """Solve every instance of ``batch`` (dict in the layout of scenarios.make_batch) on the CPU.
    ``threads`` > 1 splits the batch over a thread pool (ctypes releases the GIL)."""
    cfg = cfg or config_from_batch(batch, mode)
    B = batch["x_init"].shape[0]
    threads = max(1, min(threads, B))
    if threads == 1:
        return _solve_chunk(cfg, batch, 0, B, u_guess)
    step = max(1, min(64, (B + 4 * threads - 1) // (4 * threads)))
    bounds = [(i, min(B, i + step)) for i in range(0, B, step)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda b: _solve_chunk(cfg, batch, b[0], b[1], u_guess), bounds))
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}
