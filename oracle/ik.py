"""CPU restatement of ManipulatorPanda3DoF.inverse_transformation (robot_models/manipulator_3DoF.py:79-133).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference hands the 3-variable NLP
    min (x(q) - xt)^2 + (z(q) - zt)^2   s.t.  q1 in [-pi/2, pi/2], q2 in [-3pi/4, 0], q3 in [0, 3pi/2]   (:118-123)
to IPOPT (absent here).  Two equations in three unknowns: the minimisers form a curve, IPOPT returns the point its
path ends at, and the only recorded answer (:223, for target (0.6, 0, 0.1), start unknown) is one point of it.
What can be checked, and is: the restated NLP itself (`residual`, `LO`, `HI`, against oracle/model.arm_fk which
follows the reference's long trig expressions :101-102), an independent bounded solve (`solve_scipy`), and the
recurrence the device kernel runs (`solve_lm`, plain Python floats, same order of operations as
csrc/mmpc_episode.cuh::ik_solve)."""
import math

import numpy as np

from . import model as M

LO = np.array([-math.pi / 2, -math.pi * 3 / 4, 0.0])     # lbg :123
HI = np.array([math.pi / 2, 0.0, math.pi * 3 / 2])       # ubg :123


def residual(q, target):
    """(x(q) - xt, z(q) - zt) with the reference's expressions (:101-102 == forward_tranformation :58-75)."""
    e = M.arm_fk(np.asarray(q, float))[0]
    return np.array([float(e[0]) - target[0], float(e[1]) - target[2]])


def solve_scipy(q0, target):
    """Independent solve of the same bounded least-squares problem (SciPy trust-region reflective)."""
    from scipy.optimize import least_squares
    q0 = np.clip(np.asarray(q0, float), LO + 1e-12, HI - 1e-12)
    r = least_squares(lambda q: residual(q, target), q0, bounds=(LO, HI), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    return r.x, float(r.fun @ r.fun)


def _segments(q):
    t1 = q[0]; t2 = q[0] - q[1]; t3 = q[0] - q[1] - q[2]
    s1, c1, s2, c2, s3, c3 = math.sin(t1), math.cos(t1), math.sin(t2), math.cos(t2), math.sin(t3), math.cos(t3)
    vr = (M.A2 * s1 + M.A3 * c1, -M.A3 * c2 + M.A5 * s2, M.A6 * c3 - M.A7 * s3)
    vh = (M.A2 * c1 - M.A3 * s1, M.A3 * s2 + M.A5 * c2, -M.A6 * s3 - M.A7 * c3)
    return vr, vh


def solve_lm(q0, target):
    """Projected Levenberg-Marquardt, the recurrence of ik_solve (csrc/mmpc_episode.cuh).  Returns (q, status)."""
    xt, zt = float(target[0]), float(target[2])
    lo, hi = [float(v) for v in LO], [float(v) for v in HI]
    q = [min(max(float(q0[i]), lo[i]), hi[i]) for i in range(3)]
    lam = 1e-3
    vr, vh = _segments(q)
    r0, r1 = (vr[0] + vr[1]) + vr[2] - xt, (vh[0] + vh[1]) + vh[2] - zt
    for _ in range(200):
        J0 = [(vh[0] + vh[1]) + vh[2], -(vh[1] + vh[2]), -vh[2]]
        J1 = [-((vr[0] + vr[1]) + vr[2]), vr[1] + vr[2], vr[2]]
        g = [J0[i] * r0 + J1[i] * r1 for i in range(3)]
        fr = [not ((q[i] <= lo[i] and g[i] > 0) or (q[i] >= hi[i] and g[i] < 0)) for i in range(3)]
        gn2 = 0.0
        for i in range(3):
            if fr[i]:
                gn2 += g[i] * g[i]
        if math.sqrt(gn2) < 1e-14:
            break
        H = [[(J0[i] * J0[j] + J1[i] * J1[j] + (lam if i == j else 0.0)) if (fr[i] and fr[j]) else (1.0 if i == j else 0.0)
              for j in range(3)] for i in range(3)]
        b = [-g[i] if fr[i] else 0.0 for i in range(3)]
        d0 = H[0][0]; l10 = H[1][0] / d0; l20 = H[2][0] / d0
        d1 = H[1][1] - l10 * H[1][0]; l21 = (H[2][1] - l20 * H[1][0]) / d1
        d2 = H[2][2] - l20 * H[2][0] - l21 * l21 * d1
        y0 = b[0]; y1 = b[1] - l10 * y0; y2 = b[2] - l20 * y0 - l21 * y1
        s2 = y2 / d2; s1 = y1 / d1 - l21 * s2; s0 = y0 / d0 - l10 * s1 - l20 * s2
        qn = [min(max(q[i] + st, lo[i]), hi[i]) for i, st in enumerate((s0, s1, s2))]
        vrn, vhn = _segments(qn)
        n0, n1 = (vrn[0] + vrn[1]) + vrn[2] - xt, (vhn[0] + vhn[1]) + vhn[2] - zt
        if n0 * n0 + n1 * n1 < r0 * r0 + r1 * r1:
            q, vr, vh, r0, r1 = qn, vrn, vhn, n0, n1
            lam = max(lam / 3, 1e-12)
        else:
            lam *= 4
            if lam > 1e8:
                break
    return np.array(q), int(r0 * r0 + r1 * r1 > 1e-10)
