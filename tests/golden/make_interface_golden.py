#!/usr/bin/env python
"""Golden of the reference's own closed loop: /root/reference/demo_wholebody_qref.py (scenario 1, N = 20, dt = 0.1) run
UNMODIFIED through /root/reference/interface_wholebody_qref.py::Interface on this repo's drop-in MPCWholeBody
(tests/refshim/loader.py: import swap of INTEGRATION.md section 1, physical_sim forced to False), every solve done by
the CPU oracle (tests/refshim/oracle_controller.py).

    python tests/golden/make_interface_golden.py      -> tests/golden/interface_demo1.npz

Contents: the Interface's x_log / u_log, the task flag printed at every step, and the trace of every
controller.solve() call the Interface made (inputs, weights, terminal-equality flag, U*, cost) -- the GPU tier replays
that trace through the CUDA class, the CPU tier checks oracle/episode.py (the restated Interface) against the logs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "refshim")]
import loader              # noqa: E402
import oracle_controller   # noqa: E402


def run():
    oracle_controller.MPCWholeBody.TRACE.clear()
    g = loader.run_demo(oracle_controller)
    w = g["world"]
    tr = list(oracle_controller.MPCWholeBody.TRACE)
    flags = loader.flags_from_stdout(g["__stdout__"])
    out = dict(x_log=np.asarray(w.x_log, float), u_log=np.asarray(w.u_log, float),
               flags=np.array([loader.FLAGS.index(f) for f in flags], np.int32), final_flag=loader.FLAGS.index(w.task_flag),
               local_pose_target=np.asarray(w.local_pose_target, float), q_target=np.asarray(w.traj_ref[-1, 6:], float),
               N=g["N"], dt=g["dt"], t_move=g["t_move"], t_manipulate=g["t_manipulate"])
    for k in ("x_init", "x_ref", "u_ref", "u_last", "Qd", "Pd", "U"):
        out["call_" + k] = np.stack([t[k] for t in tr])
    for k in ("flag", "status", "iters"):
        out["call_" + k] = np.array([t[k] for t in tr], np.int32)
    out["call_cost"] = np.array([t["cost"] for t in tr])
    return out


if __name__ == "__main__":
    o = run()
    np.savez_compressed(os.path.join(HERE, "interface_demo1.npz"), **o)
    first = {loader.FLAGS[f]: int(np.argmax(o["flags"] == f)) + 1 for f in sorted(set(o["flags"].tolist()))}
    print("steps", len(o["x_log"]), "solves", len(o["call_cost"]), "first step of each flag", first, "final", loader.FLAGS[int(o["final_flag"])])
