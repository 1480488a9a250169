"""Golden of one whole "push the button" episode (demo scenario 1, demo_wholebody_qref.py:10-44) from the CPU
restatement of Interface (oracle/episode.py): the step at which every task flag is first seen, the IK answer and the
final state.  Regression anchor for the state machine; the numbers come from the oracle solver (parity unpinned).
usage: python tests/golden/make_episode_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mobile_manipulator_mpc_b200 import scenarios  # noqa: E402
from oracle.episode import Episode                 # noqa: E402


def run(rows_on_sN):
    x_start, tgt, planes = scenarios.demo_scenario(1)
    ep = Episode(0.1, 5, 2, x_start, tgt, scenarios.DEMO_CIRCLES, planes, N=20, terminal_rows_on_sN=rows_on_sN)
    first = {}
    while ep.active and ep.steps < 400:
        ep.step()
        first.setdefault(ep.flag, ep.steps)
    return dict(first_step=first, steps=ep.steps, final_state=[float(v) for v in ep.state],
                local_pose_target=[float(v) for v in ep.local_pose_target], q_target=[float(v) for v in ep.traj_ref[-1, 6:]])


if __name__ == "__main__":
    out = dict(literal=run(0), rows_on_sN=run(1))
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "episode_demo1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
