"""Golden FK vectors from the reference's own symbolic DH derivation.

Imports /root/reference/utils/dh_to_kinematics.py (pure sympy, runs in the authoring container),
takes the translation columns of "T frame 3/5/7 to frame 0" (= joint 2, joint 3, endpoint, the
expressions hard-coded at robot_models/manipulator_3DoF.py:30-70), substitutes the link lengths
of robot_models/manipulator_3DoF.py:18-22 and evaluates them at seeded random joint angles.
Output: tests/golden/fk_dh_golden.npz   (q[256,3], endpoint[256,3], joint2[256,3], joint3[256,3])
"""
import contextlib
import importlib.util
import io
import os

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/dh_to_kinematics.py"


def main():
    spec = importlib.util.spec_from_file_location("dh_to_kinematics", REF)
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    Ts = mod.Ts
    consts = {mod.a2: 0.316, mod.a3: 0.0825, mod.a5: 0.384, mod.a6: 0.088, mod.a7: 0.107}
    fns = {}
    for name, idx in (("joint2", 3), ("joint3", 5), ("endpoint", 7)):
        col = sp.simplify(Ts[idx][:3, 3].subs(consts))
        fns[name] = sp.lambdify((mod.q1, mod.q2, mod.q3), list(col), "numpy")
    rng = np.random.default_rng(20240611)
    q = np.column_stack([rng.uniform(-np.pi / 2, np.pi / 2, 256), rng.uniform(-np.pi, 0, 256),
                         rng.uniform(0, 1.5 * np.pi, 256)])
    out = {"q": q}
    for name, f in fns.items():
        vals = f(q[:, 0], q[:, 1], q[:, 2])
        out[name] = np.column_stack([np.broadcast_to(np.asarray(v, float), (256,)) for v in vals])
    np.savez(os.path.join(HERE, "fk_dh_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
