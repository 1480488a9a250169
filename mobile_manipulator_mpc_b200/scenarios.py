"""Workload definitions: the reference demo's scenario tables and the synthetic batches of
BASELINE.json's configs.  Host-side NumPy only; every array is float64
and instance-major, exactly the layout include/mmpc.h documents.

Reference sources (relative to the reference checkout):
  scenario tables      demo_wholebody_qref.py:18-44
  base goal / x_target interface_wholebody_qref.py:23-32
  global reference     interface_wholebody_qref.py:247-266 (globalPlan2D)
  local window         interface_wholebody_qref.py:353-396 (calcLocalRefTraj)
"""
import numpy as np

PI = np.pi
WORKING_RADIUS = 0.6  # interface_wholebody_qref.py:23

# demo_wholebody_qref.py:40-44
DEMO_CIRCLES = np.array([[2.5, 3.0, 0.6], [2.5, 1.0, 0.6], [5 - 0.6, 5.0, 0.1]])
# demo_wholebody_qref.py:20,29
DEMO_POSE_TARGET = np.array([5 - 0.6, 5, 0.606 + 0.333 + 0.5, -PI])


def demo_scenario(scenario):
    """(x_start, global_pose_target, planes[n_pl,6]) of demo_wholebody_qref.py:18-38."""
    if scenario == 1:
        x_start = np.array([0, 0, 0, 0, 0, 0, -PI / 4, -PI, PI])
        p = [5.007 - 0.43, 5, 0.27 + 0.606 + 0.333]
        planes = np.array([p + [0, 0, -1], p + [-1, 0, 0], p + [0, 1, 0]], dtype=float)
        return x_start, DEMO_POSE_TARGET.copy(), planes
    if scenario == 2:
        x_start = np.zeros(9)
        p = [2.5, 2, 0.35 + 0.606 + 0.333]
        r = 1 / np.sqrt(2)
        planes = np.array([p + [r, 0, r], p + [-r, 0, r]], dtype=float)
        return x_start, DEMO_POSE_TARGET.copy(), planes
    return np.zeros(9), np.array([-0.6, 0, 0.606 + 0.333 + 0.5, -PI]), np.zeros((0, 6))


def base_target(x_start, pose_target):
    """interface_wholebody_qref.py:24-32."""
    return np.array([pose_target[0] - WORKING_RADIUS * np.cos(pose_target[3]),
                     pose_target[1] - WORKING_RADIUS * np.sin(pose_target[3]),
                     pose_target[3], 0, 0, 0, x_start[6], x_start[7], x_start[8]])


def global_plan_2d(x_start, x_target, t_move, dt):
    """interface_wholebody_qref.py:247-266: linspace states, zero control reference."""
    n = int(t_move / dt)
    return np.linspace(x_start, x_target, n + 1), np.zeros((n, 5))


def local_window(traj_ref, u_ref, state, idx, N):
    """calcLocalRefTraj (interface_wholebody_qref.py:353-396): nearest global row over the state
    indices ``idx`` (first minimum), rows [i*, i*+N]; past the end the last row is repeated."""
    idx = np.asarray(idx)
    dist = np.linalg.norm(traj_ref[:, idx] - state[idx], axis=1)
    i0 = int(np.argmin(dist))
    rows = np.minimum(np.arange(i0, i0 + N + 1), traj_ref.shape[0] - 1)
    urows = np.minimum(np.arange(i0, i0 + N), u_ref.shape[0] - 1)
    return traj_ref[rows].copy(), u_ref[urows].copy()


def local_window_batch(traj_ref, u_ref, states, idx, N):
    """Vectorised local_window for states [B,9] sharing one global reference."""
    idx = np.asarray(idx)
    d = np.linalg.norm(traj_ref[None, :, idx] - states[:, None, idx], axis=2)
    i0 = np.argmin(d, axis=1)
    rows = np.minimum(i0[:, None] + np.arange(N + 1)[None, :], traj_ref.shape[0] - 1)
    urows = np.minimum(i0[:, None] + np.arange(N)[None, :], u_ref.shape[0] - 1)
    return traj_ref[rows], u_ref[urows]


XLIM = np.array([[-100, -100, -np.inf, -2, -2, -PI, -PI / 2, -PI, 0],
                 [100, 100, np.inf, 2, 2, PI, PI / 2, 0, 3 * PI / 2]])


def _perturbed_states(rng, ref, B):
    """progress index U{0..30}, then bounded perturbations, clipped like
    MPCWholeBody.solve (controllers/mpc_wholebody_qref.py:290-291)."""
    i0 = rng.integers(0, 31, size=B)
    d = np.empty((B, 9))
    d[:, 0:6] = rng.uniform(-0.3, 0.3, size=(B, 6))
    d[:, 6] = rng.uniform(-0.3, 0.3, size=B)
    d[:, 7] = rng.uniform(-0.6, 0.0, size=B)
    d[:, 8] = rng.uniform(0.0, 0.6, size=B)
    x = ref[i0] + d
    return np.clip(x, XLIM[0], XLIM[1])


def _random_circles(rng, xy, n, lo=0.5, hi=5.5, rmin=0.1, rmax=0.6, clearance=0.5):
    """n circles per instance: centres U([lo,hi]^2), radius U(rmin,rmax), redrawn while closer
    than r + 0.4 + 0.1 to the instance's start position."""
    B = xy.shape[0]
    c = np.empty((B, n, 3))
    todo = np.ones((B, n), dtype=bool)
    while todo.any():
        k = int(todo.sum())
        cand = np.stack([rng.uniform(lo, hi, k), rng.uniform(lo, hi, k), rng.uniform(rmin, rmax, k)], axis=1)
        c[todo] = cand
        d = np.linalg.norm(c[:, :, :2] - xy[:, None, :], axis=2)
        todo = d < c[:, :, 2] + clearance
    return c


def make_batch(config, B, N=None, seed=None):
    """Synthetic batch of BASELINE.json config ``config`` (1, 2, 3 or 5; config 4 is the closed
    loop built from config 3's generator, see closed_loop.py).  Returns a dict with the arrays of
    include/mmpc.h plus 'n_obs', 'n_pl', 'N', 'dt', 'obs_per_stage'."""
    dt = 0.1
    if config == 1:
        N = N or 20
        x_start, tgt, planes = demo_scenario(1)
        ref, uref = global_plan_2d(x_start, base_target(x_start, tgt), 5, dt)
        xr, ur = local_window(ref, uref, x_start, [0, 1], N)
        B = B or 1
        return dict(N=N, dt=dt, n_obs=3, n_pl=3, obs_per_stage=0,
                    x_init=np.tile(x_start, (B, 1)), x_ref=np.tile(xr, (B, 1, 1)), u_ref=np.tile(ur, (B, 1, 1)),
                    u_last=np.zeros((B, N, 5)), circles=np.tile(DEMO_CIRCLES, (B, 1, 1)),
                    planes=np.tile(planes, (B, 1, 1)), n_pl_inst=np.full(B, 3, np.int32))
    if config == 2:
        N = N or 20
        rng = np.random.default_rng(2 if seed is None else seed)
        x_start, tgt, planes = demo_scenario(2)
        ref, uref = global_plan_2d(x_start, base_target(x_start, tgt), 5, dt)
        x0 = _perturbed_states(rng, ref, B)
        xr, ur = local_window_batch(ref, uref, x0, [0, 1], N)
        return dict(N=N, dt=dt, n_obs=3, n_pl=2, obs_per_stage=0, x_init=x0, x_ref=np.ascontiguousarray(xr),
                    u_ref=np.ascontiguousarray(ur), u_last=np.zeros((B, N, 5)),
                    circles=np.tile(DEMO_CIRCLES, (B, 1, 1)), planes=np.tile(planes, (B, 1, 1)),
                    n_pl_inst=np.full(B, 2, np.int32))
    if config in (3, 5):
        N = N or (20 if config == 3 else 40)
        rng = np.random.default_rng(config if seed is None else seed)
        x0 = np.empty((B, 9)); xr = np.empty((B, N + 1, 9)); ur = np.zeros((B, N, 5))
        planes = np.zeros((B, 3, 6)); npl = np.empty(B, np.int32)
        for sc, sel in ((1, np.arange(B) % 2 == 0), (2, np.arange(B) % 2 == 1)):
            nb = int(sel.sum())
            if nb == 0:
                continue
            x_start, tgt, pl = demo_scenario(sc)
            ref, uref = global_plan_2d(x_start, base_target(x_start, tgt), 5, dt)
            xs = _perturbed_states(rng, ref, nb)
            a, b = local_window_batch(ref, uref, xs, [0, 1], N)
            x0[sel], xr[sel], ur[sel] = xs, a, b
            planes[sel, :pl.shape[0]] = pl
            if pl.shape[0] < 3:  # pad by duplicating the last plane (max-invariant); n_pl_inst keeps the true count
                planes[sel, pl.shape[0]:] = pl[-1]
            npl[sel] = pl.shape[0]
        circ = _random_circles(rng, x0[:, :2], 16)
        out = dict(N=N, dt=dt, n_obs=16, n_pl=3, obs_per_stage=0, x_init=x0, x_ref=xr, u_ref=ur,
                   u_last=np.zeros((B, N, 5)), circles=circ, planes=planes, n_pl_inst=npl)
        if config == 5:
            # moving circles: centre at stage k = o + k*dt*v, |v| ~ U(0,0.5) m/s, heading U(0,2pi)
            # (the reference's moving_obs branch is absent, README.md:85-88; semantics are ours)
            speed = rng.uniform(0, 0.5, (B, 16)); head = rng.uniform(0, 2 * PI, (B, 16))
            v = np.stack([speed * np.cos(head), speed * np.sin(head)], axis=2)
            k = np.arange(N + 1)[None, :, None, None] * dt
            mov = np.empty((B, N + 1, 16, 3))
            mov[..., :2] = circ[:, None, :, :2] + k * v[:, None, :, :]
            mov[..., 2] = circ[:, None, :, 2]
            out.update(circles=mov, obs_per_stage=1)
        return out
    raise ValueError(f"unknown config {config}")


def manipulate_instance(N=20, dt=0.1, q_target=(0.12261333, -1.36948989, 3.02634729)):
    """The manipulate-phase fixture of SURVEY.md 8(c)(iv): base parked at (5,5,-pi), weights
    diag(500,500,500,0,0,1,1,1,1) (interface_wholebody_qref.py:212-215), joint-space linspace
    reference to the IK answer for the local target (0.607, 0, 0.5) (:204-211,284-293)."""
    x = np.array([5, 5, -PI, 0, 0, 0, -PI / 4, -PI, PI])
    x_t = np.hstack([x[:6], np.asarray(q_target)])
    ref = np.linspace(x, x_t, int(2 / dt) + 1)
    uref = np.zeros((ref.shape[0] - 1, 5))
    xr, ur = local_window(ref, uref, x, [6, 7, 8], N)
    _, _, planes = demo_scenario(1)
    return dict(N=N, dt=dt, n_obs=3, n_pl=3, obs_per_stage=0, x_init=x[None], x_ref=xr[None], u_ref=ur[None],
                u_last=np.zeros((1, N, 5)), circles=DEMO_CIRCLES[None].copy(), planes=planes[None].copy(),
                n_pl_inst=np.full(1, 3, np.int32),
                Qd=np.array([500, 500, 500, 0, 0, 1, 1, 1, 1.0]))


def approach_instance(N=20, dt=0.1):
    """An 'approach'-phase instance (interface_wholebody_qref.py:155-167): scenario 1, the base within
    2 m of its goal (5, 5) and still moving, reference window = the tail of the global plan (last row
    repeated), and the one-time terminal equality X[N,:2] == X_ref[N,:2] switched on (flags bit 0)."""
    x_start, tgt, planes = demo_scenario(1)
    ref, uref = global_plan_2d(x_start, base_target(x_start, tgt), 5, dt)
    x = np.array([3.7, 3.9, 0.6, 0.9, 0.8, 0.1, -PI / 4, -PI, PI])
    xr, ur = local_window(ref, uref, x, [0, 1], N)
    return dict(N=N, dt=dt, n_obs=3, n_pl=3, obs_per_stage=0, x_init=x[None], x_ref=xr[None], u_ref=ur[None],
                u_last=np.zeros((1, N, 5)), circles=DEMO_CIRCLES[None].copy(), planes=planes[None].copy(),
                n_pl_inst=np.full(1, 3, np.int32), flags=np.ones(1, np.uint8))


def episode_batch(B, seed=7):
    """Starts and targets of B "push the button" episodes: episode 0 is demo scenario 1 and
    episode 1 demo scenario 2 exactly (demo_wholebody_qref.py:17-44); the others perturb the start pose and joint
    angles of those two.  Returns x_start [B,9], global_pose_target [B,4], circles [B,3,3], planes [B,3,6] (scenario 2
    padded by repeating its last plane, which leaves the max over the planes unchanged), n_pl_inst [B]."""
    rng = np.random.default_rng(seed)
    xs = np.empty((B, 9)); gps = np.empty((B, 4)); planes = np.zeros((B, 3, 6)); npl = np.empty(B, np.int32)
    for b in range(B):
        x_start, tgt, pl = demo_scenario(1 + (b % 2))
        x = x_start.copy()
        if b >= 2:
            x[0:2] += rng.uniform(-0.5, 0.5, 2)
            x[2] += rng.uniform(-0.5, 0.5)
            x[6] += rng.uniform(0.0, 0.3)
            x[7] = min(0.0, x[7] + rng.uniform(0.0, 0.4))
            x[8] = max(0.0, x[8] - rng.uniform(0.0, 0.4))
        xs[b], gps[b] = x, tgt
        planes[b, :pl.shape[0]] = pl
        planes[b, pl.shape[0]:] = pl[-1]
        npl[b] = pl.shape[0]
    return xs, gps, np.tile(DEMO_CIRCLES, (B, 1, 1)), planes, npl


def make_base_batch(B, N=10, seed=6, n_obs=3, dt=0.1):
    """Instances of the base-only controller MPCBase (controllers/mpc_base.py), NATIVE layout:
    x_init [B,6], x_ref [B,N+1,6], u_ref [B,N,2], circles [B,n_obs,3].  Start = a point of the straight-line plan from the
    origin to (5, 5) plus bounded perturbations (as config 2, base part), window by calcLocalRefTraj semantics; circles: the
    demo's three, or n_obs random ones as in config 3."""
    rng = np.random.default_rng(seed)
    x_start = np.zeros(6)
    x_target = np.array([5.0, 5.0, -PI, 0, 0, 0])
    ref = np.linspace(x_start, x_target, 51)
    uref = np.zeros((50, 2))
    i0 = rng.integers(0, 31, size=B)
    x0 = ref[i0] + rng.uniform(-0.3, 0.3, size=(B, 6))
    lim = np.array([[-100, -100, -np.inf, -2, -2, -PI], [100, 100, np.inf, 2, 2, PI]])
    x0 = np.clip(x0, lim[0], lim[1])
    d = np.linalg.norm(ref[None, :, :2] - x0[:, None, :2], axis=2)
    j0 = np.argmin(d, axis=1)
    rows = np.minimum(j0[:, None] + np.arange(N + 1)[None, :], 50)
    urows = np.minimum(j0[:, None] + np.arange(N)[None, :], 49)
    circ = np.tile(DEMO_CIRCLES, (B, 1, 1)) if n_obs == 3 else _random_circles(rng, x0[:, :2], n_obs)
    return dict(N=N, dt=dt, n_obs=n_obs, x_init=x0, x_ref=ref[rows].copy(), u_ref=uref[urows].copy(), circles=circ)


POSE_XLIM = np.array([[-100, -100, -np.inf, -2, -2, -PI, -PI / 2, -PI * 3 / 4, 0],
                      [100, 100, np.inf, 2, 2, PI, PI / 2, 0, PI]])            # controllers/mpc_wholebody.py:17-20


def make_pose_batch(B, N=10, seed=8, n_obs=3, dt=0.1):
    """Instances of the pose-reference whole-body controller (controllers/mpc_wholebody.py): x_init [B,9],
    x_ref [B,N+1,9] with the END-POINT pose reference (x, y, z, psi) in the first four columns (the rest zero), u_ref [B,N,5],
    u_last [B,N,5], circles [B,n_obs,3].  The pose reference is the end-point pose along a straight-line plan of the state from a
    perturbed start towards a goal near the demo's table, so it is reachable; circles: the demo's three or n_obs random ones."""
    from .robot_models import MobileManipulator
    rng = np.random.default_rng(seed)
    x_start = np.array([0, 0, 0, 0, 0, 0, -PI / 4, -PI / 2, PI / 2])
    x_goal = np.array([4.0, 4.5, 0.8, 0, 0, 0, 0.2, -1.2, 0.9])
    plan = np.linspace(x_start, x_goal, 51)
    i0 = rng.integers(0, 31, size=B)
    x0 = plan[i0] + rng.uniform(-0.2, 0.2, size=(B, 9)) * np.array([1, 1, 1, 0.5, 0.5, 0.5, 0.3, 0.3, 0.3])
    x0 = np.clip(x0, POSE_XLIM[0] + 1e-3, POSE_XLIM[1] - 1e-3)
    rows = np.minimum(i0[:, None] + 2 + np.arange(N + 1)[None, :], 50)
    robot = MobileManipulator(dt)
    pose = np.array([np.asarray(robot.forward_tranformation(p)[0], float).reshape(4) for p in plan])     # [51, 4]
    x_ref = np.zeros((B, N + 1, 9)); x_ref[:, :, :4] = pose[rows]
    circ = np.tile(DEMO_CIRCLES, (B, 1, 1)) if n_obs == 3 else _random_circles(rng, x0[:, :2], n_obs)
    return dict(N=N, dt=dt, n_obs=n_obs, n_pl=0, obs_per_stage=0, x_init=x0, x_ref=x_ref, u_ref=np.zeros((B, N, 5)),
                u_last=np.zeros((B, N, 5)), circles=circ, planes=None, n_pl_inst=None,
                Qd=np.array([5., 5, 5, 5, 0, 0, 0, 0, 0]), Pd=np.array([50., 50, 50, 50, 0, 0, 0, 0, 0]))
