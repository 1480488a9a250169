"""Stand-in for the reference's pybullet layer (simulation/albert_robot.py needs gymnasium/urdfenvs, absent here).
interface_wholebody_qref.py:9 imports it at module top; with physical_sim=False nothing in it is ever called.  TEST INFRASTRUCTURE."""
