import numpy as np


class Base:
    """Unicycle base with acceleration inputs (reference: robot_models/base.py)."""

    def __init__(self, dt):
        self.dt = dt
        self.base_length = 2 * (0.7 / 2 + 0.157)
        self.base_width = 0.52

    def base_radius(self):
        return 0.4  # robot_models/base.py:15

    def f_kinematics(self, x, u, limited_yaw=False):
        """Explicit-Euler step, robot_models/base.py:17-31; x = [x y psi dx dy dpsi], u = [dV dw]."""
        x = np.asarray(x, dtype=float).reshape(-1)
        u = np.asarray(u, dtype=float).reshape(-1)
        dt = self.dt
        nxt = np.array([
            x[0] + dt * x[3],
            x[1] + dt * x[4],
            x[2] + dt * x[5],
            x[3] + dt * (u[0] * np.cos(x[2]) - x[4] * x[5]),
            x[4] + dt * (u[0] * np.sin(x[2]) + x[3] * x[5]),
            x[5] + dt * u[1],
        ])
        if limited_yaw:
            nxt[2] = np.fmod(nxt[2] + np.pi, 2 * np.pi) - np.pi
        return nxt.reshape(1, 6)
