#!/usr/bin/env python
"""bench.py -- benchmarks of the batched whole-body MPC solver (BASELINE.json metric: "converged MPC solves/sec
(batch 65,536) at 1/2/4/8 B200; p50 solve latency").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5] [--nlp reference|clean] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

Default = the headline: BASELINE config 3 (batch 65,536 per GPU, mixed scenarios 1/2, 16 random circles, horizon 20, cold
start) solved as the REFERENCE'S NLP to the letter (MMPC_MODE_REFERENCE: stale plane columns, terminal self-collision rows on
s[N-1]; controllers/mpc_wholebody_qref.py:57-89, :257-265).  A "step" = one batched MPCWholeBody.solve over B independent
instances.  Weak scaling: every rank solves its own B instances (no data-path collective; NCCL only reduces the statistics).
--config 2 / 5: the same measurement on BASELINE configs 2 (B 4,096, scenario 2) and 5 (B 32,768, N 40, moving circles);
--config 4: the 500-step closed loop (B 16,384 in total, strong scaling); --config 1: one instance (latency).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every definition used here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged_mpc_solves_per_sec"
UNIT = "solves/s"
C_RIC = 4871                 # dense Riccati factor + solve of one stage, nx = 9, nu = 5 (Frison-Jorgensen count)
FP64_NOMINAL_TFLOPS = 37.2   # 148 SMs x 64 FP64 FMA/clk x 2 x 1.965 GHz
CONFIG_DEFAULT_BATCH = {1: 1, 2: 4096, 3: 65536, 4: 16384, 5: 32768}
STATUS_NAMES = ("converged", "max_iter", "linesearch", "factor", "nan", "acceptable")


def c_eval(n_obs, n_pl, nlp):
    """Algorithmic FP64 flops of one stage evaluation: 1,290 + 40 per circle + 122 per plane row; the
    reference's NLP carries 6 n_pl plane rows per stage (:76-89), the stage-separable one 6."""
    rows = 6 * n_pl if nlp == "reference" else 6 * (n_pl > 0)
    return 1290 + 40 * n_obs + 122 * rows


def flops_per_iteration(N, n_obs, n_pl, nlp):
    """F_iter(N, n_obs, n_pl) = N C_ric + (N + 1) C_eval: one interior-point iteration of one instance."""
    return N * C_RIC + (N + 1) * c_eval(n_obs, n_pl, nlp)


def bytes_per_solve(N, n_obs, n_pl):
    """Algorithmic HBM bytes of one solve (inputs read once, outputs written once): 6,216 B at N 20, 16 circles, 3 planes."""
    din = 9 + (N + 1) * 9 + N * 5 + N * 5 + n_obs * 3 + n_pl * 6
    dout = N * 5 + (N + 1) * 9 + (N + 1) + 3
    return 8 * (din + dout)


def phase_model(N, n_obs, n_pl, nlp):
    """Bytes of solver STATE that one instance streams through HBM in ONE launch of each phase kernel of the staged solver
    (DESIGN.md "Kernels"): every double a stage reads or writes, once, nothing cached between kernels -- the implementation's
    traffic model, next to the algorithmic 6,216 B per solve.  R = inequality rows with slack per stage."""
    R = n_obs + 4 + (6 if n_pl else 0) + (6 * (n_pl - 1) if nlp == "reference" and n_pl > 1 else 0)
    obs = 3 * n_obs + 6 * n_pl                       # static obstacle data: once per instance and launch
    st = N + 1
    ev_r = 24 + 18 + 28 + 2 * R + 19 + 10            # x u s lam | x+ lam+ | box z | t z | refs | u bounds
    ev_w = 8 + 9 + 78 + 8                            # FK cache | defect | stage QP | KKT partials
    so_r = 8 + 78 + 9 + 6 + 120 + 6                  # partials | stage QP | defect | dyn coefs | factors (roll-out) | slack column
    so_w = 120 + 24                                  # Riccati factors | dx du ds lam+
    sp_r = 15 + 15 + 28 + 2 * R + 8 + 19 + 10 + 9    # x u s | dx du ds | box z | t z | FK cache | refs | u bounds | defect
    sp_w = R + 6                                     # dt | step partials
    tr_r = 24 + 24 + 18 + 28 + 3 * R + 19 + 10       # x u s lam | step | x+ dx+ | box z | t z dt | refs | u bounds
    tr_w = 52 + 2 * R + 4 + ev_w                     # candidate iterate | merit partials | the fused evaluation's outputs
    ce = c_eval(n_obs, n_pl, nlp)
    return dict(eval=dict(bytes=8 * (st * (ev_r + ev_w) + obs), flops=st * ce),
                solve=dict(bytes=8 * st * (so_r + so_w), flops=N * C_RIC),
                step=dict(bytes=8 * (st * (sp_r + sp_w) + obs), flops=0),
                trial=dict(bytes=8 * (st * (tr_r + tr_w) + obs), flops=st * ce))   # fused: evaluates the accepted candidate


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_workload(cid, B, seed):
    from mobile_manipulator_mpc_b200 import scenarios
    return scenarios.make_batch(cid if cid in (1, 2, 3, 5) else 3, B, seed=seed)


def workload_name(cid, B, N, n_obs):
    return {1: "BASELINE config 1: demo scenario 1, single instance, N=%d" % N,
            2: "BASELINE config 2: batch %d per GPU, perturbed initial states, scenario 2 (3 circles, 2 planes), N=%d, cold start" % (B, N),
            3: "BASELINE config 3 shape: batch %d per GPU, mixed scenarios 1/2 (n_pl 3/2), %d random circles, N=%d, dt=0.1, cold start (u_last=0)" % (B, n_obs, N),
            5: "BASELINE config 5: batch %d per GPU, %d moving circles (time-varying rows), N=%d (2x default), cold start" % (B, n_obs, N)}[cid]


def mode_of(nlp):
    from mobile_manipulator_mpc_b200 import _abi
    return _abi.MODE_REFERENCE if nlp == "reference" else _abi.MODE_CLEAN


def cpu_baseline(batch, sample, threads, nlp):
    """Oracle port (oracle/mmpc_oracle.c) on the host cores, bounded sample of the same workload."""
    from oracle import solver as osolver
    sub = {k: (v[:sample] if isinstance(v, np.ndarray) else v) for k, v in batch.items()}
    osolver.lib()
    t = time.perf_counter()
    o = osolver.solve(sub, mode=mode_of(nlp), threads=threads)
    dt = time.perf_counter() - t
    conv = int((o["status"] == 0).sum())
    return conv / dt, conv, dt, o


def demo_episode_latency(mode, device, nlp, with_cpu=True):
    """BASELINE config 1 in its natural habitat: the 198 MPCWholeBody.solve calls the reference's own Interface made in its demo
    episode (scenario 1; recorded by tests/golden/make_interface_golden.py with the unmodified interface_wholebody_qref.py), each
    solved alone (B = 1) with the recorded inputs, weights and warm start.  Latency per call on the GPU (device-resident
    inputs, CUDA events) and through the host API (mmpc_solve_host), and of the CPU port on one core on the same calls."""
    import torch
    from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
    from mobile_manipulator_mpc_b200 import scenarios
    g = np.load(os.path.join(ROOT, "tests", "golden", "interface_demo1.npz"))
    _, _, planes = scenarios.demo_scenario(1)
    n = len(g["call_cost"])
    N, dt = int(g["N"]), float(g["dt"])
    S = BatchSolver(N=N, dt=dt, n_obs=3, n_pl=3, B_max=1, device=device, mode=mode)

    def call(i):
        return dict(N=N, dt=dt, n_obs=3, n_pl=3, obs_per_stage=0, x_init=g["call_x_init"][i][None], x_ref=g["call_x_ref"][i][None],
                    u_ref=g["call_u_ref"][i][None], u_last=g["call_u_last"][i][None], circles=scenarios.DEMO_CIRCLES[None].copy(),
                    planes=planes[None].copy(), n_pl_inst=np.full(1, 3, np.int32), flags=np.full(1, int(g["call_flag"][i]), np.uint8))
    dev_ms, host_ms, iters, ok = [], [], [], 0
    out = None
    for rep in range(2):   # first pass warms up
        dev_ms, host_ms, iters, ok = [], [], [], 0
        for i in range(n):
            if i == 0 or not np.array_equal(g["call_Qd"][i], g["call_Qd"][i - 1]):
                S.set_weights(Q=np.diag(g["call_Qd"][i]), P=np.diag(g["call_Pd"][i]))
            b = call(i)
            d = S.to_device(b)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = S.solve_device(d, out=out); e.record(); torch.cuda.synchronize()
            dev_ms.append(a.elapsed_time(e))
            t = time.perf_counter(); oh = S.solve_host(b); host_ms.append((time.perf_counter() - t) * 1e3)
            iters.append(int(oh["iters"][0])); ok += int(oh["status"][0] == 0)
            assert abs(float(oh["cost"][0]) - float(g["call_cost"][i])) <= 1e-5 * max(1e-12, abs(float(g["call_cost"][i]))), i
    res = dict(calls=n, converged=ok, solver=S.last_solver(), mean_iterations=float(np.mean(iters)), median_iterations=float(np.median(iters)),
               device_p50_ms=float(np.median(dev_ms)), device_p99_ms=float(np.percentile(dev_ms, 99)), device_mean_ms=float(np.mean(dev_ms)),
               host_api_p50_ms=float(np.median(host_ms)), host_api_p99_ms=float(np.percentile(host_ms, 99)), host_api_mean_ms=float(np.mean(host_ms)),
               ms_per_iteration=float(np.sum(dev_ms) / np.sum(iters)),
               note="every call's optimal cost equals the recorded one to 1e-5 relative (asserted)")
    S.close()
    if with_cpu:
        from oracle import solver as osolver
        osolver.lib()
        cpu_ms = []
        for i in range(n):
            b = call(i)
            cfg = osolver.config_from_batch(b, mode)
            cfg.Qd[:] = list(g["call_Qd"][i]); cfg.Pd[:] = list(g["call_Pd"][i])
            t = time.perf_counter(); osolver.solve(b, cfg=cfg, mode=mode, threads=1); cpu_ms.append((time.perf_counter() - t) * 1e3)
        res.update(cpu_port_p50_ms=float(np.median(cpu_ms)), cpu_port_p99_ms=float(np.percentile(cpu_ms, 99)), cpu_port_mean_ms=float(np.mean(cpu_ms)))
    return res


def run_reference(args):
    """--impl reference: the reference path on the host cores.  CasADi/IPOPT cannot be installed in this image (no wheel, no
    network), so this arm times the oracle port -- the same interior-point algorithm on the same NLP -- with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cid = args.config if args.config in (1, 2, 3, 5) else 3
    sample = args.ref_sample if cid != 1 else 1
    batch = make_workload(cid, max(sample, 64) if cid != 1 else 1, seed=3)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, conv, dt, _ = cpu_baseline(batch, sample, cores, args.nlp)
        if i >= args.warmup:
            vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=float(np.mean(times) * 1e3), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload=workload_name(cid, sample, batch["N"], batch["n_obs"]) + " (bounded sample of the GPU arm's batch shape)",
                            batch_per_step=sample, horizon=batch["N"], n_obs=batch["n_obs"], nlp=args.nlp),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port",
                                  sample=f"{sample} instances per step, oracle/mmpc_oracle.c over {cores} threads; "
                                         "CasADi/IPOPT not installable in this image"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    _emit(line)


def _emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL's version banner ...)
    has been sent to stderr by _guard_stdout()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def _guard_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _dist_setup():
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local, dist


def pinned_copy(batch):
    """The step's inputs in PINNED host memory (the contract's e2e source); NumPy views of torch pinned tensors."""
    import torch
    out, keep = {}, []
    for k, v in batch.items():
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            keep.append(t)
            out[k] = t.numpy()
        else:
            out[k] = v
    out["_keep"] = keep
    return out


def run_batched(args):
    """configs 2, 3, 5 (and 1 as a batch of one): cold batched solves."""
    import ctypes as C
    import torch
    from mobile_manipulator_mpc_b200._lib import lib, check
    from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
    from mobile_manipulator_mpc_b200 import scenarios

    rank, world, local, dist = _dist_setup()
    cid = args.config
    B = args.batch or CONFIG_DEFAULT_BATCH[cid]
    T = max(1, args.contexts if B > 1 else 1)
    nlp = args.nlp
    mode = mode_of(nlp)
    # T solver contexts (handle + stream + workspace + its own batch each): the thin tail of slow instances of one batch
    # overlaps the bulk of another; every context solves DIFFERENT instances (seed per rank and context)
    ctx = []
    for t in range(T):
        batch = make_workload(cid, B, seed=3 + 1000 * rank + 17 * t)
        S = BatchSolver(N=batch["N"], dt=batch["dt"], n_obs=batch["n_obs"], n_pl=batch["n_pl"], B_max=B, device=local, mode=mode,
                        obs_per_stage=batch["obs_per_stage"])
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            dev_in = S.to_device(batch)
            out = S.solve_device(dev_in)
        ctx.append(dict(S=S, stream=st, dev_in=dev_in, out=out, batch=batch, pinned=None, pinned_out=None, steps=0))
    torch.cuda.synchronize()
    batch0 = ctx[0]["batch"]
    N, n_obs, n_pl = batch0["N"], batch0["n_obs"], batch0["n_pl"]
    S = ctx[0]["S"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(nsteps, contexts, host=False):
        """nsteps solves spread over `contexts` concurrent solver contexts; a context takes the next step when it is free."""
        outs = [None] * contexts
        nxt = [0]
        lock = threading.Lock()
        for c in ctx:
            c["steps"] = 0

        def work(t):
            c = ctx[t]
            torch.cuda.set_device(local)
            with torch.cuda.stream(c["stream"]):
                while True:
                    with lock:
                        if nxt[0] >= nsteps:
                            break
                        nxt[0] += 1
                    if host:
                        outs[t] = c["S"].solve_host(c["pinned"], out=c["pinned_out"])
                    else:
                        outs[t] = c["S"].solve_device(c["dev_in"], out=c["out"])
                        c["stream"].synchronize()   # the context is free again when its solve has finished (one graph launch)
                    c["steps"] += 1
        if contexts == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(t,)) for t in range(contexts)]
            [x.start() for x in th]; [x.join() for x in th]
        return outs

    # ---- device-resident timing (value): K steps over T contexts ----
    run_steps(max(args.warmup, T), T)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l1 = sum(c["S"].launch_count() for c in ctx)
    t0 = time.perf_counter()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_all0.record()
    run_steps(args.steps, T)
    for c in ctx:
        torch.cuda.current_stream().wait_stream(c["stream"])
    e_all1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    launches = sum(c["S"].launch_count() for c in ctx) - l1
    total_ms = e_all0.elapsed_time(e_all1)
    # every context re-solves its own batch: converged instances and iterations per solve are per context
    conv_total, iters_total, solved_total = 0, 0, 0
    status_hist = np.zeros(6, np.int64)
    for c in ctx:
        stt = c["out"]["status"].cpu().numpy()
        c["conv"] = int((stt == 0).sum())
        c["iters_np"] = c["out"]["iters"].cpu().numpy()
        conv_total += c["conv"] * c["steps"]
        iters_total += int(c["iters_np"].sum()) * c["steps"]
        solved_total += B * c["steps"]
        status_hist += np.bincount(stt, minlength=6)[:6] * c["steps"]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    cc = torch.tensor([conv_total, iters_total, solved_total], dtype=torch.float64, device="cuda")
    hh = torch.from_numpy(status_hist.astype(np.float64)).cuda()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(cc, op=dist.ReduceOp.SUM); dist.all_reduce(hh, op=dist.ReduceOp.SUM)
    total_ms_max = float(t.item())
    conv_all, iters_all, solved_all = (float(v) for v in cc.tolist())
    value = conv_all / (total_ms_max * 1e-3)

    # ---- one context alone: latency of one batched solve through the graph driver ----
    lat_ms = []
    for i in range(min(args.steps, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx[0]["stream"]):
            a.record(); S.solve_device(ctx[0]["dev_in"], out=ctx[0]["out"]); b.record()
        torch.cuda.synchronize()
        lat_ms.append(a.elapsed_time(b))
    step_ms = float(np.median(lat_ms))
    solver_kind = S.last_solver()   # 'resident' (AUTO's choice for small batches) or 'staged'
    conv0, iters0_np = ctx[0]["conv"], ctx[0]["iters_np"]
    iters_sum = int(iters0_np.sum())
    # ---- the same solve with the host sequencing the rounds and CUDA events around every launch: the per-kernel durations
    #      the roofline is computed from (same kernels, same lists; only the driver differs) ----
    phase_ms, phase_ln, rounds, prof_ms = {}, {}, 0, []
    if rank == 0 or dist is None:
        S.set_profile(True)
        nprof = min(args.steps, 3)
        for i in range(nprof):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ctx[0]["stream"]):
                a.record(); S.solve_device(ctx[0]["dev_in"], out=ctx[0]["out"]); b.record()
            torch.cuda.synchronize()
            prof_ms.append(a.elapsed_time(b))
            pm, pl, rounds = S.phase_times()
            for k in pm:
                phase_ms[k] = phase_ms.get(k, 0.0) + pm[k] / nprof
                phase_ln[k] = pl[k]
        S.set_profile(False)
    # p50 latency of ONE instance (B = 1, BASELINE config 1: demo scenario 1), device-resident, same NLP
    lat1 = None
    if rank == 0:
        b1 = scenarios.make_batch(1, 1)
        S1 = BatchSolver(N=b1["N"], dt=b1["dt"], n_obs=b1["n_obs"], n_pl=b1["n_pl"], B_max=1, device=local, mode=mode)
        d1 = S1.to_device(b1); o1 = S1.solve_device(d1); torch.cuda.synchronize()
        l1 = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); S1.solve_device(d1, out=o1); b.record(); torch.cuda.synchronize()
            l1.append(a.elapsed_time(b))
        lat1 = dict(p50_ms=float(np.median(l1)), p99_ms=float(np.percentile(l1, 99)), iterations=int(o1["iters"][0]))
        S1.close()
    demo = None
    if rank == 0 and cid == 1:
        demo = demo_episode_latency(mode, local, nlp, with_cpu=not args.no_cpu_baseline)
    # side figure: the same batch as the other NLP variant, one context, two timed solves
    other = None
    if rank == 0 and not args.no_side:
        onlp = "clean" if nlp == "reference" else "reference"
        SR = BatchSolver(N=N, dt=batch0["dt"], n_obs=n_obs, n_pl=n_pl, B_max=B, device=local, mode=mode_of(onlp),
                         obs_per_stage=batch0["obs_per_stage"])
        dR = SR.to_device(batch0); oR = SR.solve_device(dR); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); SR.solve_device(dR, out=oR); SR.solve_device(dR, out=oR); b.record(); torch.cuda.synchronize()
        convR = int((oR["status"] == 0).sum().item())
        other = dict(nlp=onlp, single_context_value=convR * 2 / (a.elapsed_time(b) * 1e-3), ms_per_step=a.elapsed_time(b) / 2,
                     converged_fraction=convR / B, mean_iterations=float(oR["iters"].double().mean().item()))
        SR.close()
    barrier()

    # ---- end to end through the C ABI (mmpc_solve_host) with PINNED host buffers: H2D of the step's inputs, solve, D2H of
    #      every result array, per step ----
    for c in ctx:
        c["pinned"] = pinned_copy(c["batch"])
        c["pinned_out"] = c["S"].host_outputs(B, pinned=True)
    run_steps(T, T, host=True)
    barrier()
    t0 = time.perf_counter()
    outs_h = run_steps(args.steps, T, host=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    conv_h = sum(int((o["status"] == 0).sum()) * c["steps"] for o, c in zip(outs_h, ctx) if o is not None)
    out_h = next(o for o in outs_h if o is not None)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ce = torch.tensor([float(conv_h)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX); dist.all_reduce(ce, op=dist.ReduceOp.SUM)
    e2e_value = float(ce.item()) / float(te.item())
    h2d = sum(np.asarray(batch0[k]).nbytes for k in ("x_init", "x_ref", "u_ref", "u_last", "circles", "planes", "n_pl_inst"))
    d2h = sum(v.nbytes for k, v in out_h.items() if k != "_keep")

    # ---- NCCL gather of the per-instance result (u0, status): statistics only, outside the solve ----
    gathered = None
    if dist is not None:
        u0 = ctx[0]["out"]["U"][:, 0, :].contiguous()
        allu0 = torch.empty((world * B, 5), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allu0, u0)
        gathered = int(allu0.shape[0])

    if rank == 0:
        fp64 = C.c_double()
        check(lib().mmpc_bench_fp64(local, C.byref(fp64)))
        fp64_peak = float(fp64.value)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except OSError:
            hbm_src = "fallback 6650 GB/s"
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        model = phase_model(N, n_obs, n_pl, nlp)
        # units = instance passes one launch series processes: the starting point of every instance is evaluated once by the
        # stand-alone kernel; every iteration is one Riccati solve, one step and one (credited) trial + evaluation.  Extra
        # line-search trials, inertia retries and the final convergence-test pass are NOT credited.
        units = dict(eval=B, solve=iters_sum, step=iters_sum, trial=iters_sum)
        prof_step_ms = float(np.mean(prof_ms)) if prof_ms else step_ms
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except OSError:
            pass
        phases = {}
        for k in ("eval", "solve", "step", "trial"):
            ms = phase_ms.get(k, 0.0)
            gbs = model[k]["bytes"] * units[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            tfs = model[k]["flops"] * units[k] / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            phases[k] = dict(ms=ms, share=ms / prof_step_ms, launches=phase_ln.get(k, 0), units=units[k],
                             streamed_bytes_per_unit=model[k]["bytes"], flops_per_unit=model[k]["flops"],
                             hbm_gbs=gbs, hbm_frac=gbs / hbm_peak, fp64_tflops=tfs, fp64_frac=tfs / fp64_peak)
            assert phases[k]["hbm_frac"] <= 1.0 and phases[k]["fp64_frac"] <= 1.0, (k, phases[k])
        for k in ("compact", "ctrl_step", "ctrl_trial", "init", "pose"):
            phases[k] = dict(ms=phase_ms.get(k, 0.0), share=phase_ms.get(k, 0.0) / prof_step_ms, launches=phase_ln.get(k, 0))
        dom = max(("eval", "solve", "step", "trial"), key=lambda k: phases[k]["ms"])
        kname = {"eval": "mmpc::staged_eval_kernel", "solve": "mmpc::staged_solve_team_kernel",
                 "step": "mmpc::staged_step_kernel", "trial": "mmpc::staged_trial_kernel"}[dom]
        nl = max(1, phases[dom]["launches"])
        tr = traffic.get(nlp, traffic).get(dom, {}).get("dram_bytes_per_instance") if traffic else None
        work_flops = float(sum(flops_per_iteration(N, n_obs, int(p), nlp) * int(i) for p, i in zip(batch0["n_pl_inst"], iters0_np)))
        alg_b = bytes_per_solve(N, n_obs, n_pl)
        streamed_b = sum(model[k]["bytes"] * units[k] for k in units) / B
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload=workload_name(cid, B, N, n_obs),
                                batch_per_gpu=B, horizon=N, n_obs=n_obs, n_pl_max=n_pl, nlp=nlp,
                                solver=("resident: one persistent thread block per instance, solver state in shared memory, inputs by "
                                        "cp.async.bulk, atomic work queue (AUTO's choice for this batch size)") if solver_kind == "resident"
                                else "staged: one CUDA graph per solve, device-side WHILE loops over the rounds",
                                contexts=T,
                                l2_policy="solver state %.1f GB per context and inputs+outputs %.0f MB per step exceed the 126 MB L2; "
                                          "every context solves its own instances" % (S.workspace_bytes() / 1e9, (h2d + d2h) / 1e6),
                                converged_fraction=conv_all / solved_all, mean_iterations=iters_all / solved_all,
                                status_histogram={n: int(v) for n, v in zip(STATUS_NAMES, hh.tolist())},
                                rounds=rounds, single_context_ms_per_step=step_ms,
                                p50_batched_solve_latency_ms=step_ms, single_instance_latency=lat1, demo_episode_latency=demo, other_nlp=other,
                                single_context_value=conv0 / (step_ms * 1e-3), wall_ms_timed_region=wall_ms),
                    clocks=clocks, gpu_launches=int(launches),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             source="pinned host buffers -> mmpc_solve_host (H2D, solve, D2H of U X s cost kkt iters status)"),
                    roofline=dict(bound="fp64", achieved=phases[dom]["fp64_tflops"], peak=fp64_peak, unit="TFLOP/s",
                                  frac=phases[dom]["fp64_frac"],
                                  traffic=None if tr is None else tr * units[dom] / nl,
                                  kernel=kname, kernel_ms=phases[dom]["ms"] / nl, kernel_launches=nl,
                                  kernel_share_of_step=phases[dom]["share"],
                                  flops_per_unit=model[dom]["flops"], units_per_step=units[dom],
                                  peak_source="DFMA micro-benchmark in this run (mmpc_bench_fp64); nominal %.1f" % FP64_NOMINAL_TFLOPS,
                                  frac_of_nominal=phases[dom]["fp64_tflops"] / FP64_NOMINAL_TFLOPS,
                                  measured_in="single-context pass, host-sequenced rounds, CUDA events around every launch (%.1f ms per "
                                              "step against %.1f ms through the graph)" % (prof_step_ms, step_ms),
                                  whole_solve=dict(achieved=work_flops / (step_ms * 1e-3) / 1e12, unit="TFLOP/s",
                                                   frac=work_flops / (step_ms * 1e-3) / 1e12 / fp64_peak,
                                                   note="sum over instances of iterations x F_iter (SURVEY 8(d)) / graph-driven step"),
                                  hbm=dict(peak=hbm_peak, peak_source=hbm_src, unit="GB/s",
                                           algorithmic_bytes_per_solve=alg_b, streamed_state_bytes_per_solve=streamed_b,
                                           achieved_algorithmic=alg_b * B / (step_ms * 1e-3) / 1e9,
                                           achieved_streamed=streamed_b * B / (step_ms * 1e-3) / 1e9,
                                           frac_algorithmic=alg_b * B / (step_ms * 1e-3) / 1e9 / hbm_peak,
                                           frac_streamed=streamed_b * B / (step_ms * 1e-3) / 1e9 / hbm_peak,
                                           note="not the bound: the algorithmic traffic is inputs + outputs only"),
                                  phases=phases))
        if solver_kind == "resident":
            # the timed solves ran on ONE kernel: it is the dominant kernel; the per-phase figures above are the staged solver's
            # (profiling pass) and stay in the line for comparison only
            rf = line["roofline"]
            tfs = work_flops / (step_ms * 1e-3) / 1e12
            res_traffic = traffic.get("resident", {}).get("dram_bytes_per_instance") if traffic else None
            rf.update(achieved=tfs, frac=tfs / fp64_peak, frac_of_nominal=tfs / FP64_NOMINAL_TFLOPS, kernel="mmpc_res::resident_solve_kernel",
                      kernel_ms=step_ms, kernel_launches=1, kernel_share_of_step=1.0, flops_per_unit=work_flops / max(1, iters_sum),
                      units_per_step=iters_sum, traffic=None if res_traffic is None else res_traffic * B,
                      measured_in="CUDA events around the solve (set-up kernel, queue reset, resident kernel), one context")
            rf["staged_phases_for_comparison"] = rf.pop("phases")
        assert line["roofline"]["frac"] <= 1.0
        if gathered is not None:
            line["config"]["nccl_gathered_rows"] = gathered
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ns = min(args.cpu_sample, B)
            v, cconv, cdt, _ = cpu_baseline(batch0, ns, cores, nlp)
            v1, _, cdt1, _ = cpu_baseline(batch0, min(128, ns), 1, nlp)   # serial leg
            line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind="port",
                                        sample="first %d instances of rank 0's batch, %s NLP, oracle/mmpc_oracle.c on %d threads, %.1f s"
                                               % (ns, nlp, cores, cdt),
                                        serial_value=v1, serial_sample="first %d instances on 1 thread, %.1f s" % (min(128, ns), cdt1))
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_closed_loop(args):
    """BASELINE config 4: 500-step receding-horizon rollout on the device (window, solve, warm-start shift of the guess,
    plant), batch 16,384 in TOTAL split over the ranks (strong scaling), K interleaved sub-batches per GPU."""
    import torch
    from mobile_manipulator_mpc_b200 import closed_loop
    rank, world, local, dist = _dist_setup()
    Btot = args.batch or CONFIG_DEFAULT_BATCH[4]
    K = max(1, args.contexts)
    steps = args.cl_steps
    per_rank = Btot // world
    b, x_glob = closed_loop.config4(Btot)
    lo = rank * per_rank
    subs = []
    for i in range(K):
        a, e = lo + i * per_rank // K, lo + (i + 1) * per_rank // K
        bi = {k: (v[a:e] if isinstance(v, np.ndarray) else v) for k, v in b.items()}
        subs.append(closed_loop.ClosedLoop(bi, x_glob[a:e], device=local, shift_guess=True, mode=mode_of(args.nlp)))
    streams = [torch.cuda.Stream() for _ in subs]

    def drive(nsteps, record):
        """one host thread: every step of every sub-batch is a handful of asynchronous launches + one graph launch"""
        ev = []
        for s in range(nsteps):
            for L, st in zip(subs, streams):
                with torch.cuda.stream(st):
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(); L.step(count=True); e1.record()
                    if record:
                        ev.append((e0, e1))
        return ev

    drive(args.warmup, False)
    torch.cuda.synchronize()
    for L in subs:
        L.reset_counters()
    if dist is not None:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ev = drive(steps, True)
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    z.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = a.elapsed_time(z)
    lat = np.array([e0.elapsed_time(e1) for e0, e1 in ev]).reshape(steps, K)
    conv_steps = sum(L.conv_per_step().cpu().numpy() for L in subs)       # [steps] converged instances per step of this rank
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    c = torch.from_numpy(conv_steps.astype(np.float64)).cuda()
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    if rank == 0:
        frac = (c / (per_rank * world)).cpu().numpy()
        tot = float(c.sum().item())
        line = dict(metric="closed_loop_converged_instance_steps_per_sec", value=tot / (float(t.item()) * 1e-3), unit="instance-steps/s",
                    n_gpus=world, steps=steps, warmup=args.warmup, ms_per_step=float(t.item()) / steps, higher_is_better=True,
                    scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload="BASELINE config 4: closed-loop %d-step receding-horizon rollout, batch %d in total, 16 random circles, "
                                         "device-side window + solve + warm-start shift + plant" % (steps, per_rank * world),
                                batch_total=per_rank * world, batch_per_gpu=per_rank, sub_batches_per_gpu=K, nlp=args.nlp, horizon=b["N"],
                                p50_step_latency_ms=float(np.median(lat.max(axis=1))), p50_sub_batch_step_latency_ms=float(np.median(lat)),
                                converged_fraction=dict(mean=float(frac.mean()), min=float(frac.min()), first=float(frac[0]), last=float(frac[-1]))),
                    clocks=clocks, gpu_launches=int(sum(L.solver.launch_count() for L in subs)))
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (3 = the headline)")
    ap.add_argument("--nlp", default="reference", choices=["reference", "clean"],
                    help="reference = the reference's NLP to the letter (default); clean = the stage-separable variant")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU per step (default: the config's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contexts", type=int, default=3, help="concurrent solver contexts (streams) per GPU")
    ap.add_argument("--cl-steps", type=int, default=500, help="config 4: closed-loop steps")
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--ref-sample", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the other-NLP side figure")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == 4:
        return run_closed_loop(args)
    return run_batched(args)


if __name__ == "__main__":
    main()
