"""CPU tier: the CUDA solver *sources* (csrc/mmpc_staged.cuh + mmpc_team.cuh + mmpc_parts.cuh) executed by the CPU
emulator of tests/emu against the independent dense oracle
(oracle/mmpc_oracle.c).  This is how the kernels are debugged on the GPU-less authoring box; the
real parity tests are the -m gpu ones."""
import numpy as np
import pytest

from mobile_manipulator_mpc_b200 import _abi, scenarios
from oracle import solver
from tests.emu import emu


def _compare(batch, kernel, tol_cost=1e-6, tol_u=1e-4):
    cfg = solver.config_from_batch(batch, mode=_abi.MODE_CLEAN)
    o = solver.solve(batch, cfg=cfg, threads=4)
    e = emu.solve(batch, cfg, kernel=kernel)
    both = (o["status"] == 0) & (e["status"] == 0)
    assert both.mean() >= 0.9
    rel = np.abs(o["cost"] - e["cost"])[both] / np.abs(o["cost"][both])
    du = np.abs(o["U"][:, 0] - e["U"][:, 0]).max(axis=1)[both]
    assert (rel < tol_cost).mean() >= 0.95 and (du < tol_u).mean() >= 0.95
    return o, e


# staged = phase bodies + 16-lane team Riccati (the product default); staged_thread = phase bodies +
# one-thread-per-instance Riccati
@pytest.mark.parametrize("kernel", ["staged", "staged_thread"])
def test_config1_bit_level_agreement(kernel):
    b = scenarios.make_batch(1, 1)
    o, e = _compare(b, kernel)
    assert o["iters"][0] == e["iters"][0]
    assert np.abs(o["X"] - e["X"]).max() < 1e-9 and np.abs(o["U"] - e["U"]).max() < 1e-9


@pytest.mark.parametrize("kernel,B3,B5", [("staged", 4, 2), ("staged_thread", 12, 3)])
def test_config3_and_moving_obstacles(kernel, B3, B5):
    _compare(scenarios.make_batch(3, B3), kernel)
    _compare(scenarios.make_batch(5, B5), kernel)


def test_staged_rounds_and_list_compaction():
    """The staged solver finishes in max(iterations + line-search retries) + 1 rounds, every instance
    leaves the lists exactly once, and the thread and team Riccati agree to rounding."""
    b = scenarios.make_batch(2, 6)
    cfg = solver.config_from_batch(b, mode=_abi.MODE_CLEAN)
    a = emu.solve(b, cfg, kernel="staged_thread")
    t = emu.solve(b, cfg, kernel="staged")
    assert (a["status"] == 0).all() and (t["status"] == 0).all()
    assert a["rounds"] >= a["iters"].max() + 1 and a["rounds"] <= a["iters"].max() + 60
    assert (a["iters"] == t["iters"]).all()
    assert np.abs(a["U"] - t["U"]).max() < 1e-8 and np.abs(a["cost"] - t["cost"]).max() < 1e-8 * np.abs(a["cost"]).max()


@pytest.mark.parametrize("kernel", ["staged", "staged_thread", "staged_fat"])
def test_terminal_xy_equality_flag(kernel):
    """flags bit 0 = opti.subject_to(X[N,:2] == X_ref[N,:2]) (interface_wholebody_qref.py:167)."""
    b = scenarios.approach_instance()
    cfg = solver.config_from_batch(b, mode=_abi.MODE_CLEAN)
    o = solver.solve(b, cfg=cfg, threads=1)
    e = emu.solve(b, cfg, kernel=kernel)
    assert o["status"][0] == 0 and e["status"][0] == 0
    assert np.abs(e["X"][0, -1, :2] - b["x_ref"][0, -1, :2]).max() < 1e-8          # the equality holds
    assert abs(e["cost"][0] - o["cost"][0]) <= 1e-9 * abs(o["cost"][0])
    assert np.abs(e["U"][0, 0] - o["U"][0, 0]).max() < 1e-7
    free = dict(b); free["flags"] = None                                            # and it binds: without it x_N differs
    f = emu.solve(free, cfg, kernel=kernel)
    assert np.abs(f["X"][0, -1, :2] - b["x_ref"][0, -1, :2]).max() > 1e-3


def _terminal_rows_instance():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_terminal_rows_instance.npz"))
    Q = np.array([25.0, 25, 0, 0, 0, 5, 5, 5, 5])
    return dict(N=20, dt=0.1, n_obs=3, n_pl=3, obs_per_stage=0, x_init=g["x_init"][None], x_ref=g["x_ref"][None], u_ref=np.zeros((1, 20, 5)),
                u_last=g["u_last"][None], circles=g["circles"][None], planes=g["planes"][None], n_pl_inst=np.array([int(g["n_pl"])], np.int32),
                flags=np.zeros(1, np.uint8), Qd=Q, Pd=Q)


def test_reference_mode_terminal_rows_both_variants():
    """SURVEY.md 8(a) row 9 on a path-sensitive instance (found in a closed loop next to demo scenario 2's aerial obstacle):
    the reference-mode NLP with the terminal self-collision rows bounded by s[N-1] (the reference to the letter,
    cfg.terminal_rows_on_sN = 0) and by s[N] (= 1) end in different local optima of the non-smooth NLP although none of those
    rows is active.  The kernel source implements both and must land where the oracle lands in each; both optima must be
    feasible points of the dense restated NLP."""
    from oracle import nlp
    b = _terminal_rows_instance()
    cost = {0: 457.41613461, 1: 475.98075895}
    for v in (0, 1):
        cfg = solver.config_from_batch(b, _abi.MODE_REFERENCE, terminal_rows_on_sN=v)
        ora = solver.solve(b, cfg=cfg)
        ker = emu.solve(b, cfg, kernel="staged")
        assert ora["status"][0] == 0 and ker["status"][0] == 0
        assert abs(ora["cost"][0] - cost[v]) < 1e-5
        assert abs(ker["cost"][0] - ora["cost"][0]) < 1e-5 * ora["cost"][0]
        assert np.abs(ker["U"][0, 0] - ora["U"][0, 0]).max() < 1e-4
        assert abs(int(ker["iters"][0]) - int(ora["iters"][0])) <= 4
        P = nlp.from_batch(b, 0, "reference")   # the dense restatement has the rows on s[N-1]; both points are feasible for it
        assert P.violation(P.pack(ora["X"][0], ora["U"][0], ora["s"][0])) < 1e-6


def test_reference_mode_literal_terminal_rows_kernel_vs_oracle():
    """The literal reference NLP (terminal rows on s[N-1], the default) in the kernel source against the oracle on seeded
    batches of configs 2, 3 and 5: same optimum within the north-star tolerances."""
    for cid, B in ((2, 4), (3, 6), (5, 2)):
        b = scenarios.make_batch(cid, B)
        cfg = solver.config_from_batch(b, _abi.MODE_REFERENCE)
        assert cfg.terminal_rows_on_sN == 0
        ora = solver.solve(b, cfg=cfg, threads=4)
        ker = emu.solve(b, cfg, kernel="staged")
        both = (ora["status"] == 0) & (ker["status"] == 0)
        assert both.sum() >= B - 1
        rel = np.abs(ora["cost"] - ker["cost"])[both] / np.abs(ora["cost"][both])
        du0 = np.abs(ora["U"][:, 0] - ker["U"][:, 0]).max(axis=1)[both]
        assert (rel < 1e-5).all() and (du0 < 1e-4).all(), (cid, rel.max(), du0.max())
