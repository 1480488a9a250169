"""TEST INFRASTRUCTURE: ctypes front-end of the CPU lane emulator of the solver kernel."""
import ctypes as C
import os
import subprocess

import numpy as np

from mobile_manipulator_mpc_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = C.CDLL(os.path.join(_HERE, "_build", "libmmpc_emu_staged.so"))
        _LIB.staged = _LIB
        _LIB.staged.mmpc_emu_staged_solve.argtypes = [C.POINTER(_abi.MmpcConfig), C.c_int32, C.POINTER(_abi.MmpcBatchIn),
                                                      C.POINTER(_abi.MmpcBatchOut), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_int32]
        _LIB.staged.mmpc_emu_ik.argtypes = [C.c_int32] + [C.c_void_p] * 4
        _LIB.staged.mmpc_emu_episode_update.argtypes = [C.c_int32] * 4 + [C.POINTER(_abi.MmpcEpisodeIO)]
    return _LIB


def ik(q_guess, target):
    """csrc/mmpc_episode.cuh::ik_solve on the CPU; q_guess [B,3], target [B,3] -> (q [B,3], status [B])."""
    q_guess = np.ascontiguousarray(q_guess, dtype=np.float64); target = np.ascontiguousarray(target, dtype=np.float64)
    B = q_guess.shape[0]
    q = np.zeros((B, 3)); st = np.zeros(B, np.int32)
    assert lib().staged.mmpc_emu_ik(B, _abi.ptr(q_guess), _abi.ptr(target), _abi.ptr(q), _abi.ptr(st)) == 0
    return q, st


def episode_update(N, M, n_manip, io):
    """csrc/mmpc_episode.cuh::episode_update on the CPU; ``io``: dict of NumPy arrays named like MmpcEpisodeIO (in place)."""
    e = _abi.MmpcEpisodeIO()
    for name, _ in _abi.MmpcEpisodeIO._fields_:
        a = io.get(name)
        setattr(e, name, None if a is None else a.ctypes.data)
    assert lib().staged.mmpc_emu_episode_update(N, io["x"].shape[0], M, n_manip, C.byref(e)) == 0


def solve(batch, cfg, kernel="staged"):
    B = batch["x_init"].shape[0]
    N = cfg.N
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    arrs = {k: f(batch.get(k)) for k in ("x_init", "x_ref", "u_ref", "u_last", "u_guess", "circles", "planes")}
    npl = batch.get("n_pl_inst")
    npl = None if npl is None else np.ascontiguousarray(npl, dtype=np.int32)
    flags = batch.get("flags")
    flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    bi = _abi.MmpcBatchIn(*[_abi.ptr(arrs[k]) for k in ("x_init", "x_ref", "u_ref", "u_last", "u_guess", "circles", "planes")],
                          _abi.ptr(npl), _abi.ptr(flags), _abi.ptr(f(batch.get("x_guess"))))
    out = dict(U=np.zeros((B, N, 5)), X=np.zeros((B, N + 1, 9)), s=np.zeros((B, N + 1)), cost=np.zeros(B),
               kkt=np.zeros(B), iters=np.zeros(B, np.int32), status=np.zeros(B, np.int32))
    bo = _abi.MmpcBatchOut(*[_abi.ptr(out[k]) for k in ("U", "X", "s", "cost", "kkt", "iters", "status")])
    # staged: product default (team Riccati, fused trial+evaluation, warp-specialised parts);
    # staged_thread: one-thread Riccati; staged_fat: one thread per item; staged_unfused: separate eval
    assert kernel.startswith("staged"), kernel
    rounds = C.c_int32(0)
    team = 1 if kernel in ("staged", "staged_noparts") else 0
    fused = 0 if kernel == "staged_unfused" else 1
    parts = 1 if kernel in ("staged", "staged_thread") else 0
    assert lib().staged.mmpc_emu_staged_solve(C.byref(cfg), B, C.byref(bi), C.byref(bo), C.byref(rounds), team, fused, parts) == 0
    out["rounds"] = rounds.value
    return out
