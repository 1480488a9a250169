"""ctypes mirror of include/mmpc.h (the C ABI of the batched whole-body MPC solver).

Field order and types must match ``include/mmpc.h`` exactly; ``tests/test_abi.py`` checks the
struct sizes against the values the shared library reports.
"""
import ctypes as C

import numpy as np

NX, NU = 9, 5
MAX_PLANES = 4

OK, ERR_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED = 0, 1, 2, 3, 4
STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_LINESEARCH, STATUS_FACTOR, STATUS_NAN, STATUS_ACCEPTABLE = range(6)
STATUS_NAMES = ("converged", "max_iter", "linesearch", "factor", "nan", "acceptable")
MODE_REFERENCE, MODE_CLEAN = 0, 1
MODEL_WHOLEBODY, MODEL_BASE, MODEL_POSEREF = 0, 1, 2
KERNEL_AUTO, KERNEL_STAGED, KERNEL_STAGED_THREAD, KERNEL_STAGED_UNFUSED, KERNEL_STAGED_FAT, KERNEL_STAGED_HOSTLOOP, KERNEL_RESIDENT = 0, 3, 4, 5, 6, 7, 8

_d = C.c_double


class MmpcConfig(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("n_obs", C.c_int32), ("n_pl", C.c_int32), ("mode", C.c_int32),
        ("obs_per_stage", C.c_int32), ("max_iter", C.c_int32), ("terminal_rows_on_sN", C.c_int32), ("model", C.c_int32),
        ("dt", _d),
        ("Qd", _d * 9), ("Pd", _d * 9), ("Rd", _d * 5), ("Wd", _d * 5), ("S", _d),
        ("ulim", (_d * 5) * 2), ("xlim", (_d * 9) * 2), ("dulim", (_d * 5) * 2),
        ("base_radius", _d), ("self_collision_radius", _d), ("obstacle_expand_dist", _d),
        ("tol", _d), ("mu_init", _d), ("acceptable_tol", _d),
    ]


class MmpcBatchIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("x_init", "x_ref", "u_ref", "u_last", "u_guess", "circles", "planes", "n_pl_inst", "flags", "x_guess")]


class MmpcBatchOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("U", "X", "s", "cost", "kkt", "iters", "status")]


TASK_MOVE, TASK_APPROACH, TASK_ROTATE, TASK_MOVE_FINISH, TASK_MANIPULATE, TASK_FINISHED, TASK_IK_FAILED = range(7)
TASK_NAMES = ("move", "approach", "rotate", "move finish", "manipulate", "manipulate finish", "ik failed")


class MmpcEpisodeIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("x", "pose_target", "traj", "traj_len", "task", "flags", "wset", "active", "x_ref", "u_ref",
                 "local_pose_target", "ik_status")]


def default_config(N=20, dt=0.1, n_obs=3, n_pl=3, mode=MODE_REFERENCE):
    """Reference defaults: controllers/mpc_wholebody_qref.py:11-22,43-44,280-285;
    N, dt from demo_wholebody_qref.py:10-11.  (Pure-Python twin of mmpc_default_config.)"""
    c = MmpcConfig()
    c.N, c.n_obs, c.n_pl, c.mode, c.obs_per_stage, c.max_iter = N, n_obs, n_pl, mode, 0, 2000
    c.dt = dt
    q = [25.0, 25.0, 0.0, 0.0, 0.0, 5.0, 5.0, 5.0, 5.0]
    c.Qd[:] = q
    c.Pd[:] = q
    c.Rd[:] = [0.1, 0.1, 0.0, 0.0, 0.0]
    c.Wd[:] = [0.0, 0.0, 0.1, 0.1, 0.1]
    c.S = 1e5
    pi, inf = np.pi, np.inf
    set_limits(c,
               ulim=np.array([[-2, -pi, -1, -1, -1], [2, pi, 1, 1, 1]], dtype=float),
               xlim=np.array([[-100, -100, -inf, -2, -2, -pi, -pi / 2, -pi, 0],
                              [100, 100, inf, 2, 2, pi, pi / 2, 0, 3 * pi / 2]], dtype=float),
               dulim=np.array([[-inf, -inf, -0.5, -0.5, -0.5], [inf, inf, 0.5, 0.5, 0.5]], dtype=float))
    c.base_radius, c.self_collision_radius, c.obstacle_expand_dist = 0.4, 0.05, 0.03
    c.tol, c.mu_init, c.acceptable_tol = 1e-8, 0.1, 1e-8
    return c


def set_limits(c, ulim=None, xlim=None, dulim=None):
    for name, arr, n in (("ulim", ulim, 5), ("xlim", xlim, 9), ("dulim", dulim, 5)):
        if arr is None:
            continue
        arr = np.asarray(arr, dtype=float).reshape(2, n)
        for r in range(2):
            getattr(c, name)[r][:] = list(arr[r])


def ptr(a):
    """void* of a NumPy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)
