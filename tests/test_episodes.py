"""the callers either side of the solve -- batched inverse kinematics
(robot_models/manipulator_3DoF.py:79-133) and the task state machine (interface_wholebody_qref.py:146-228).

CPU tier: the restatements (oracle/ik.py, oracle/episode.py) against the reference's recorded numbers and a committed
golden, and the device sources (csrc/mmpc_episode.cuh) executed on the CPU by tests/emu against those restatements.
GPU tier: the kernels through the C ABI against the restatements, whole episodes free-running."""
import json
import os

import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, scenarios
from oracle import ik as IK
from oracle import model as M
from oracle.episode import Episode

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _targets(n, seed=11):
    """reachable (x, 0, z) targets: FK of random feasible joint angles; starts: other random feasible angles"""
    rng = np.random.default_rng(seed)
    q_true = rng.uniform(IK.LO, IK.HI, (n, 3))
    tg = np.zeros((n, 3))
    for i in range(n):
        e = M.arm_fk(q_true[i])[0]
        tg[i] = [float(e[0]), 0.0, float(e[1])]
    return rng.uniform(IK.LO, IK.HI, (n, 3)), tg


# ---- IK -------------------------------------------------------------------------------------------------------------
def test_ik_restatement_and_recorded_answer():
    # the one number the reference records (:223): IK answer for the target (0.6, 0, 0.1)
    r = IK.residual([0.42323673, -1.39921683, 1.15256477], [0.6, 0.0, 0.1])
    assert np.abs(r).max() < 1e-8
    # restated long-form expressions (:101-102) == the theta-chain form the kernels use
    for q in np.random.default_rng(0).uniform(-3, 3, (50, 3)):
        e1, e2 = M.arm_fk(q)[0], M.arm_fk_compact(q)[0]
        assert abs(e1[0] - e2[0]) < 1e-14 and abs(e1[1] - e2[1]) < 1e-14


def test_ik_lm_reaches_targets_within_bounds_like_scipy():
    q0, tg = _targets(40)
    for i in range(40):
        q, st = IK.solve_lm(q0[i], tg[i])
        qs, fs = IK.solve_scipy(q0[i], tg[i])
        if st == 0:
            assert np.abs(IK.residual(q, tg[i])).max() < 1e-5
            assert (q >= IK.LO - 1e-15).all() and (q <= IK.HI + 1e-15).all()
        else:  # a local stall of the projected iteration: SciPy from the same start must not do better than 1e-10 either ... or it is a miss
            assert fs > 1e-12, (q0[i], tg[i])
    q, st = IK.solve_lm([0, 0, 0], [2.0, 0.0, 0.1])   # out of reach (arm length < 1 m)
    assert st == 1


def test_ik_host_class_and_device_source_match_restatement():
    from tests.emu import emu
    from mobile_manipulator_mpc_b200.robot_models.manipulator_3DoF import ManipulatorPanda3DoF
    arm = ManipulatorPanda3DoF(0.1)
    q0, tg = _targets(64, seed=12)
    qe, ste = emu.ik(q0, tg)
    for i in range(64):
        q, st = IK.solve_lm(q0[i], tg[i])
        assert st == ste[i] and np.abs(q - qe[i]).max() < 1e-12
        if st == 0:
            assert np.abs(arm.inverse_transformation(q0[i], tg[i]) - q).max() < 1e-9
    with pytest.raises(ValueError):
        arm.inverse_transformation(np.zeros(3), np.array([2.0, 0.0, 0.1]))


# ---- state machine --------------------------------------------------------------------------------------------------
def _emu_machine(xs, gps, n_move, n_manip, N):
    B = xs.shape[0]
    Mrows = max(n_move, n_manip) + 1
    x_target = np.stack([gps[:, 0] - 0.6 * np.cos(gps[:, 3]), gps[:, 1] - 0.6 * np.sin(gps[:, 3]), gps[:, 3], np.zeros(B), np.zeros(B),
                         np.zeros(B), xs[:, 6], xs[:, 7], xs[:, 8]], axis=1)
    traj = np.zeros((B, Mrows, 9))
    for b in range(B):
        traj[b, :n_move + 1] = np.linspace(xs[b], x_target[b], n_move + 1)
    return dict(x=xs.copy(), pose_target=gps.copy(), traj=traj, traj_len=np.full(B, n_move + 1, np.int32), task=np.zeros(B, np.int32),
                flags=np.zeros(B, np.uint8), wset=np.zeros(B, np.int32), active=np.ones(B, np.int32), x_ref=np.zeros((B, N + 1, 9)),
                u_ref=np.ones((B, N, 5)), local_pose_target=np.zeros((B, 3)), ik_status=np.zeros(B, np.int32)), Mrows


def test_episode_restatement_matches_golden():
    with open(os.path.join(GOLD, "episode_demo1.json")) as f:
        gold = json.load(f)["literal"]
    x_start, tgt, planes = scenarios.demo_scenario(1)
    ep = Episode(0.1, 5, 2, x_start, tgt, scenarios.DEMO_CIRCLES, planes, N=20)
    first = {}
    while ep.active and ep.steps < 400:
        ep.step()
        first.setdefault(ep.flag, ep.steps)
    assert first == gold["first_step"] and ep.flag == "manipulate finish"
    assert np.abs(ep.state - np.array(gold["final_state"])).max() < 1e-6
    e = np.asarray(M.forward_transformation(ep.state)[0], float)[:3]
    assert np.linalg.norm(e - tgt[:3]) <= 0.01                      # the button is reached (:222)
    assert ep.term_eq == 1                                            # :167 fired on the way


def test_device_state_machine_source_follows_restatement():
    """csrc/mmpc_episode.cuh::episode_update (run on the CPU by tests/emu) driven along the restated episode: at every step
    the same task flag, weight set, terminal-equality flag and local reference as Interface.stateMachineUpdate."""
    from tests.emu import emu
    from oracle.episode import Q_DEFAULT, Q_MANIPULATE, Q_ROTATE
    names = {v: k for k, v in enumerate(_abi.TASK_NAMES)}
    for scen in (1, 2):
        x_start, tgt, planes = scenarios.demo_scenario(scen)
        ep = Episode(0.1, 5, 2, x_start, tgt, scenarios.DEMO_CIRCLES, planes, N=20)
        io, Mrows = _emu_machine(x_start[None], tgt[None], 50, 20, 20)
        seen = set()
        while ep.active and ep.steps < 400:
            io["x"][0] = ep.state                     # the machine reads the state the restatement is in
            ep.step()
            emu.episode_update(20, Mrows, 20, io)
            assert io["task"][0] == names[ep.flag], (ep.steps, ep.flag, io["task"][0])
            assert io["active"][0] == int(ep.active)
            seen.add(ep.flag)
            if ep.active:
                assert np.abs(io["x_ref"][0] - ep.local_ref).max() < 1e-12, (ep.steps, ep.flag)
                assert (io["u_ref"][0] == 0).all()
                assert io["flags"][0] == ep.term_eq
                want = {tuple(Q_DEFAULT): 0, tuple(Q_ROTATE): 1, tuple(Q_MANIPULATE): 2}[tuple(ep.Qd)]
                assert io["wset"][0] == want
                if ep.flag == "manipulate":
                    n = io["traj_len"][0]
                    assert n == ep.traj_ref.shape[0] and np.abs(io["traj"][0, :n] - ep.traj_ref).max() < 1e-12
        assert seen >= {"move", "approach", "rotate", "manipulate", "manipulate finish"}


def test_angle_diff_device_source():
    # angleDiff enters the machine through calcLocalRefPose (:407-410): checked through the 'approach' reference above;
    # here the corner cases of the fold itself (controllers/mpc_wholebody_qref.py:92-117)
    for a, b in ((3.14, -3.14), (-3.14, 3.14), (0.1, -0.1), (4.0, 1.0), (-4.0, 2.5), (7.0, -7.0)):
        d = M.angle_diff(a, b)
        assert -np.pi <= d <= np.pi + 1e-12
        assert abs(np.sin(d) - np.sin(a - b)) < 1e-12 and abs(np.cos(d) - np.cos(a - b)) < 1e-12


# ---- GPU tier -------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_ik_matches_restatement():
    import torch
    from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
    S = BatchSolver(B_max=1)
    q0, tg = _targets(4096, seed=13)
    tg[::97, 0] = 2.0  # unreachable ones in between
    dev = torch.device("cuda", 0)
    q, st = S.ik(torch.from_numpy(q0).to(dev), torch.from_numpy(tg).to(dev))
    q, st = q.cpu().numpy(), st.cpu().numpy()
    assert (st[::97] == 1).all()
    ok = st == 0
    assert ok.mean() > 0.9
    res = np.array([IK.residual(q[i], tg[i]) for i in np.nonzero(ok)[0]])
    assert np.abs(res).max() < 1e-5
    assert (q >= IK.LO - 1e-15).all() and (q <= IK.HI + 1e-15).all()
    for i in range(0, 4096, 16):   # the same recurrence on the CPU
        qo, so = IK.solve_lm(q0[i], tg[i])
        # unreachable targets end on a flat stall of the projected iteration: libm vs device sincos rounding shows there
        assert so == st[i] and np.abs(qo - q[i]).max() < (1e-9 if so == 0 else 1e-6), i
    S.close()


@pytest.mark.gpu
def test_gpu_episodes_free_running_against_restatement():
    """B = 6 whole episodes on the device (demo scenarios 1 and 2 and perturbed starts); episodes 0 and 1 are followed by the
    CPU restatement free-running (no re-seeding): same task flag at every step, states within 1e-4 after ~200 closed-loop
    steps.  Both solve the reference's NLP to the letter (terminal self-collision rows on s[N-1])."""
    from mobile_manipulator_mpc_b200.episodes import BatchedInterface
    B = 6
    xs, gps, circ, pls, npl = scenarios.episode_batch(B)
    T = BatchedInterface(0.1, 5, 2, xs, gps, circ, pls, npl, N=20, mode=_abi.MODE_REFERENCE)
    eps = [Episode(0.1, 5, 2, xs[b], gps[b], circ[b], pls[b][:npl[b]], N=20) for b in range(2)]
    worst = 0.0
    while T.steps < 600:
        n = T.step()
        task, x = T.task.cpu().numpy(), T.x.cpu().numpy()
        for b, ep in enumerate(eps):
            if ep.active:
                ep.step()
            assert _abi.TASK_NAMES[task[b]] == ep.flag, (T.steps, b)
            worst = max(worst, float(np.abs(x[b] - ep.state).max()))
        if n == 0:
            break
    assert worst < 1e-4, worst
    task, x = T.task.cpu().numpy(), T.x.cpu().numpy()
    assert (task == _abi.TASK_FINISHED).all(), task
    for b in range(B):   # every robot pushed its button (:222)
        e = np.asarray(M.forward_transformation(x[b])[0], float)[:3]
        assert np.linalg.norm(e - gps[b, :3]) <= 0.01
    assert (T.flags.cpu().numpy() == 1).all()
    T.close()
