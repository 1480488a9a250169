"""CPU tier: every analytic derivative of the C oracle (oracle/mmpc_oracle.c) against exact
complex-step derivatives of the expression-level restatement (oracle/model.py)."""
import ctypes as C

import numpy as np

from oracle import model as M
from oracle import solver

POSE = [0, 1, 2, 6, 7, 8]


def _cs_grad_hess(fun, x):
    """gradient by complex step, Hessian by central differences of the complex-step gradient"""
    n = x.size
    def grad(xx):
        g = np.empty(n)
        for i in range(n):
            xc = xx.astype(complex); xc[i] += 1e-30j
            g[i] = fun(xc).imag / 1e-30
        return g
    g = grad(x)
    H = np.empty((n, n)); h = 1e-6
    for j in range(n):
        e = np.zeros(n); e[j] = h
        H[:, j] = (grad(x + e) - grad(x - e)) / (2 * h)
    return g, H


def _row(lib, kind, idx, x, par):
    h = C.c_double(); g = np.zeros(6); H = np.zeros((6, 6))
    par = np.ascontiguousarray(par, dtype=float)
    lib.mmpc_oracle_row(kind, idx, x.ctypes.data_as(C.c_void_p), par.ctypes.data_as(C.c_void_p),
                        C.c_double(0.4), C.c_double(0.05), C.c_double(0.03), C.byref(h),
                        g.ctypes.data_as(C.c_void_p), H.ctypes.data_as(C.c_void_p))
    return h.value, g, H


def test_row_derivatives(oracle_lib):
    rng = np.random.default_rng(3)
    plane = np.array([4.577, 5.0, 1.209, 0.3, -0.5, 0.81])
    for trial in range(6):
        x = rng.uniform(-1.5, 1.5, 9)
        cases = [(0, 0, np.array([2.5, 1.0, 0.6]), lambda xx: M.circle_rows(xx, [(2.5, 1.0, 0.6)])[0])]
        for m in range(4):
            cases.append((1, m, np.zeros(3), lambda xx, m=m: M.self_collision_rows(xx)[m]))
        for i in range(6):
            cases.append((2, i, plane, lambda xx, i=i: -M.plane_margins(xx, [(plane[:3], plane[3:])])[i][0]))
        for kind, idx, par, fun in cases:
            h, g, H = _row(oracle_lib, kind, idx, x, par)
            gr, Hr = _cs_grad_hess(fun, x)
            assert abs(h - fun(x)) <= 1e-13 * max(1, abs(h))
            assert np.allclose(g, gr[POSE], rtol=1e-11, atol=1e-12), (kind, idx)
            assert np.allclose(gr[[3, 4, 5]], 0)
            assert np.allclose(H, Hr[np.ix_(POSE, POSE)], rtol=1e-6, atol=1e-7), (kind, idx)


def test_dynamics_jacobians(oracle_lib):
    rng = np.random.default_rng(4)
    for _ in range(5):
        x = rng.uniform(-1, 1, 9); u = rng.uniform(-1, 1, 5); dt = 0.1
        xn = np.zeros(9); A = np.zeros((9, 9)); B = np.zeros((9, 5))
        oracle_lib.mmpc_oracle_dynamics(x.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), C.c_double(dt),
                                        xn.ctypes.data_as(C.c_void_p), A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p))
        assert np.allclose(xn, M.f_kinematics(x, u, dt), atol=1e-15)
        for i in range(9):
            xc = x.astype(complex); xc[i] += 1e-30j
            assert np.allclose(A[:, i], M.f_kinematics(xc, u.astype(complex), dt).imag / 1e-30, atol=1e-14)
        for j in range(5):
            uc = u.astype(complex); uc[j] += 1e-30j
            assert np.allclose(B[:, j], M.f_kinematics(x.astype(complex), uc, dt).imag / 1e-30, atol=1e-14)
