import numpy as np

from .base import Base
from .manipulator_3DoF import ManipulatorPanda3DoF


class MobileManipulator:
    """Base + 3-DoF arm (reference: robot_models/mobile_manipulator.py:9-75)."""

    def __init__(self, dt):
        self.dt = dt
        self.base = Base(dt)
        self.manipulator = ManipulatorPanda3DoF(dt)
        self.baselink2joint1_x = -0.007          # :14
        self.baselink2joint1_z = 0.606 + 0.333   # :15

    def forward_tranformation(self, state):
        """pose_endpoint (1,4) [x y z psi], pos_joint_2 (1,3), pos_joint_3 (1,3) in the world frame (:17-55)."""
        state = np.asarray(state, dtype=float).reshape(-1)
        x, q = state[:6], state[6:]
        e, j2, j3 = self.manipulator.forward_tranformation(q)
        c, s = np.cos(x[2]), np.sin(x[2])
        bx, bz = self.baselink2joint1_x, self.baselink2joint1_z
        lift = lambda p: [x[0] + (p[0, 0] + bx) * c, x[1] + (p[0, 0] + bx) * s, 0 + p[0, 2] + bz]
        return (np.array([lift(e) + [x[2]]]), np.array([lift(j2)]), np.array([lift(j3)]))

    def f_kinematics(self, x, u):
        """One model step (:57-75); returns a (1,9) row like the reference's horzcat."""
        x = np.asarray(x, dtype=float).reshape(-1)
        u = np.asarray(u, dtype=float).reshape(-1)
        q_next = self.manipulator.f_kinematics(x[6:], u[2:])   # in place on x[6:], as the reference does
        x_base_next = self.base.f_kinematics(x[:6], u[:2])
        return np.hstack([x_base_next.reshape(-1), np.asarray(q_next).reshape(-1)]).reshape(1, 9)
