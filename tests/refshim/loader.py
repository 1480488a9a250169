"""Runs the reference's OWN caller code -- interface_wholebody_qref.py and demo_wholebody_qref.py, unmodified, loaded
from /root/reference -- on top of this repo's drop-in classes.  TEST INFRASTRUCTURE ONLY.

What is swapped: the packages ``controllers.mpc_wholebody_qref`` and ``robot_models`` resolve
to the drop-in modules; ``casadi``, ``matplotlib`` and ``simulation`` resolve to the stand-ins next to this file.  The one
intervention besides the import swap: the demo asks for ``physical_sim=True`` (pybullet, absent here and out of scope),
so ``Interface.__init__`` is wrapped to run with ``physical_sim=False`` -- the reference's own pybullet-free branch
(interface_wholebody_qref.py:76, :142-143).
"""
import contextlib
import importlib
import io
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MMPC_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF, "interface_wholebody_qref.py"))


@contextlib.contextmanager
def reference_caller(controller_module):
    """Context in which ``import interface_wholebody_qref`` / running the demo use the reference's files with
    ``controllers.mpc_wholebody_qref`` = ``controller_module`` (a module that defines MPCWholeBody)."""
    import mobile_manipulator_mpc_b200.robot_models as rm
    saved_path, saved = list(sys.path), dict(sys.modules)
    sys.path[:0] = [HERE, REF]
    for name in [m for m in sys.modules if m.split(".")[0] in ("casadi", "matplotlib", "simulation", "controllers", "robot_models",
                                                               "interface_wholebody_qref")]:
        del sys.modules[name]
    sys.modules["robot_models"] = rm
    for sub in ("base", "manipulator_3DoF", "mobile_manipulator", "obstacles"):
        sys.modules["robot_models." + sub] = importlib.import_module("mobile_manipulator_mpc_b200.robot_models." + sub)
    pkg = types.ModuleType("controllers")
    pkg.__path__ = []
    pkg.mpc_wholebody_qref = controller_module
    sys.modules["controllers"] = pkg
    sys.modules["controllers.mpc_wholebody_qref"] = controller_module
    try:
        iface = importlib.import_module("interface_wholebody_qref")
        assert os.path.samefile(iface.__file__, os.path.join(REF, "interface_wholebody_qref.py"))
        real_init = iface.Interface.__init__

        def init_without_pybullet(self, *a, **k):
            k["physical_sim"] = False
            if len(a) > 6:
                a = a[:6]
            real_init(self, *a, **k)

        iface.Interface.__init__ = init_without_pybullet
        yield iface
    finally:
        sys.path[:] = saved_path
        for name in list(sys.modules):
            if name not in saved:
                del sys.modules[name]
        sys.modules.update(saved)


def run_demo(controller_module, quiet=True):
    """runpy of the reference's demo_wholebody_qref.py (scenario 1, N = 20, dt = 0.1, :10-14); returns its globals
    (``world`` is the Interface after run() and plot3D(), ``mpc_controller`` the controller)."""
    with reference_caller(controller_module):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf if quiet else sys.stdout):
            g = runpy.run_path(os.path.join(REF, "demo_wholebody_qref.py"), run_name="__main__")
    g["__stdout__"] = buf.getvalue()
    return g


FLAGS = ("move", "approach", "rotate", "move finish", "manipulate", "manipulate finish")


def flags_from_stdout(text):
    """timerCallback prints ``<step>: `` + the task flag at the top of every MPC step (interface_wholebody_qref.py:102-107)."""
    out = []
    for line in text.splitlines():
        head, sep, tail = line.partition(": ")
        if sep and head.isdigit() and tail in FLAGS:
            out.append(tail)
    return out
