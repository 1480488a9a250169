import numpy as np

# classical-DH link lengths of the reduced Panda, robot_models/manipulator_3DoF.py:18-22
A2, A3, A5, A6, A7 = 0.316, 0.0825, 0.384, 0.088, 0.107


def _chain(q):
    """(endpoint, joint2, joint3) as (x, z) pairs in the arm plane; theta-chain form of the
    closed-form DH expressions at robot_models/manipulator_3DoF.py:30-70."""
    t1 = q[0]
    t2 = t1 - q[1]
    t3 = t2 - q[2]
    v1 = np.array([A2 * np.sin(t1) + A3 * np.cos(t1), A2 * np.cos(t1) - A3 * np.sin(t1)])
    v2 = np.array([-A3 * np.cos(t2) + A5 * np.sin(t2), A3 * np.sin(t2) + A5 * np.cos(t2)])
    v3 = np.array([A6 * np.cos(t3) - A7 * np.sin(t3), -A6 * np.sin(t3) - A7 * np.cos(t3)])
    return v1 + v2 + v3, v1, v1 + v2, (v1, v2, v3)


class ManipulatorPanda3DoF:
    def __init__(self, dt):
        self.dt = dt

    def forward_tranformation(self, q):
        """x_endpoint, x_joint_2, x_joint_3 as (1,3) row vectors [x, 0, z]
        (robot_models/manipulator_3DoF.py:10-77; the reference's spelling is kept)."""
        q = np.asarray(q, dtype=float).reshape(-1)
        e, j2, j3, _ = _chain(q)
        row = lambda p: np.array([[p[0], 0.0, p[1]]])
        return row(e), row(j2), row(j3)

    def inverse_transformation(self, q_initial_guess, x_target):
        """Joint angles reaching the planar target (robot_models/manipulator_3DoF.py:79-133):
        min (x(q)-xt)^2 + (z(q)-zt)^2  s.t. q1 in [-pi/2, pi/2], q2 in [-3pi/4, 0], q3 in [0, 3pi/2] (:123),
        started from q_initial_guess.  The reference calls IPOPT; this is a projected
        Levenberg-Marquardt iteration on the same 3-variable problem."""
        x_target = np.asarray(x_target, dtype=float).squeeze()
        q = np.asarray(q_initial_guess, dtype=float).squeeze().copy()
        if x_target.shape[0] != 3:
            raise ValueError("Wrong target ")
        assert x_target[1] == 0.0, "y should always be 0"
        lo = np.array([-np.pi / 2, -np.pi * 3 / 4, 0.0])
        hi = np.array([np.pi / 2, 0.0, np.pi * 3 / 2])
        q = np.clip(q, lo, hi)
        tgt = np.array([x_target[0], x_target[2]])
        T = np.array([[1, 0, 0], [1, -1, 0], [1, -1, -1]], dtype=float)  # dtheta/dq
        lam = 1e-3

        def resid(qq):
            return _chain(qq)[0] - tgt

        r = resid(q)
        for _ in range(200):
            _, _, _, (v1, v2, v3) = _chain(q)
            Jth = np.array([[v1[1], v2[1], v3[1]], [-v1[0], -v2[0], -v3[0]]])  # d(x,z)/dtheta
            J = Jth @ T
            g = J.T @ r
            free = ~(((q <= lo) & (g > 0)) | ((q >= hi) & (g < 0)))
            if np.linalg.norm(g[free]) < 1e-14:
                break
            H = J.T @ J + lam * np.eye(3)
            step = np.zeros(3)
            idx = np.nonzero(free)[0]
            step[idx] = -np.linalg.solve(H[np.ix_(idx, idx)], g[idx])
            qn = np.clip(q + step, lo, hi)
            rn = resid(qn)
            if rn @ rn < r @ r:
                q, r, lam = qn, rn, max(lam / 3, 1e-12)
            else:
                lam *= 4
                if lam > 1e8:
                    break
        if r @ r > 1e-10:
            raise ValueError(f"No solution in joint space found for given target {x_target} in cartesian space")
        return q

    def f_kinematics(self, q, q_dot):
        """robot_models/manipulator_3DoF.py:189-191 (the reference updates q in place; so do we)."""
        q += np.asarray(q_dot, dtype=float) * self.dt
        return q
