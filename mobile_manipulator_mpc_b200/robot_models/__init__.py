"""Numeric (NumPy) mirrors of the reference's robot_models package -- same class and method names,
no CasADi.  Only what MPCWholeBody and its caller touch is provided."""
from .base import Base
from .manipulator_3DoF import ManipulatorPanda3DoF
from .mobile_manipulator import MobileManipulator
from .obstacles import Obstacles

__all__ = ["Base", "ManipulatorPanda3DoF", "MobileManipulator", "Obstacles"]
