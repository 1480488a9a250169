#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched whole-body MPC solver (BASELINE.json metric:
"converged MPC solves/sec (batch 65,536)").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" = one batched MPCWholeBody.solve over B independent instances of BASELINE config 3's
shape (mixed scenarios 1/2, 16 random circles, horizon 20), cold start.  Weak scaling: every rank
solves its own B instances (no data-path collective; NCCL only reduces the statistics).
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


def flops_per_iteration(N, n_obs, n_pl):
    """Algorithmic FP64 flops of ONE interior-point iteration of one instance:
    N dense Riccati stages (nx=9, nu=5: 4,871 flop, Frison-Jorgensen count) + (N+1) stage
    evaluations (1,290 + 40 per circle + 122 per plane row, 6 plane rows in the clean NLP)."""
    c_ric = 4871
    c_eval = 1290 + 40 * n_obs + 122 * 6 * (n_pl > 0)
    return N * c_ric + (N + 1) * c_eval


def bytes_per_solve(N, n_obs, n_pl):
    """Algorithmic HBM bytes of one solve (inputs read once, outputs written once)."""
    din = 9 + (N + 1) * 9 + N * 5 + N * 5 + n_obs * 3 + n_pl * 6
    dout = N * 5 + (N + 1) * 9 + (N + 1) + 3
    return 8 * (din + dout)


def phase_model(N, n_obs, n_pl):
    """Algorithmic HBM bytes and FP64 flops that ONE instance costs ONE launch of each phase kernel of
    the staged solver (DESIGN.md "Kernels"): the doubles each stage must read and write, once, with
    nothing cached between kernels.  R = inequality rows per stage."""
    R = n_obs + 4 + (6 if n_pl else 0)
    obs = 3 * n_obs + 6 * n_pl                       # static obstacle data: once per instance and launch
    st = N + 1
    ev_r = 24 + 18 + 28 + 2 * R + 19 + 10            # x u s lam | x+ lam+ | box z | t z | refs | u bounds
    ev_w = 8 + 9 + 78 + 8                            # FK cache | defect | stage QP | KKT partials
    so_r = 8 + 78 + 9 + 6 + 120 + 6                  # partials | stage QP | defect | dyn coefs | factors (roll-out) | slack column
    so_w = 120 + 24                                  # Riccati factors | dx du ds lam+
    sp_r = 15 + 15 + 28 + 2 * R + 8 + 19 + 10 + 9    # x u s | dx du ds | box z | t z | FK cache | refs | u bounds | defect
    sp_w = R + 6                                     # dt | step partials
    tr_r = 24 + 24 + 18 + 28 + 3 * R + 19 + 10       # x u s lam | step | x+ dx+ | box z | t z dt | refs | u bounds
    tr_w = 52 + 2 * R + 4                            # candidate iterate | merit partials
    c_ric, c_eval = 4871, 1290 + 40 * n_obs + 122 * 6 * (n_pl > 0)
    return dict(eval=dict(bytes=8 * (st * (ev_r + ev_w) + obs), flops=st * c_eval),
                solve=dict(bytes=8 * st * (so_r + so_w), flops=N * c_ric),
                step=dict(bytes=8 * (st * (sp_r + sp_w) + obs), flops=0),
                trial=dict(bytes=8 * (st * (tr_r + tr_w) + obs), flops=0))


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


def make_workload(B, seed):
    from mobile_manipulator_mpc_b200 import scenarios
    return scenarios.make_batch(3, B, seed=seed)


def cpu_baseline(batch, sample, threads):
    """Oracle port (oracle/mmpc_oracle.c) on the host cores, bounded sample of the same workload."""
    from oracle import solver as osolver
    from mobile_manipulator_mpc_b200 import _abi
    sub = {k: (v[:sample] if isinstance(v, np.ndarray) else v) for k, v in batch.items()}
    osolver.lib()
    t = time.perf_counter()
    o = osolver.solve(sub, mode=_abi.MODE_CLEAN, threads=threads)
    dt = time.perf_counter() - t
    conv = int((o["status"] == 0).sum())
    return conv / dt, conv, dt, o


def run_reference(args):
    """--impl reference: the reference path on the host cores.  CasADi/IPOPT cannot be installed in
    this image (no wheel, no network), so this arm times the oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.ref_sample
    batch = make_workload(max(sample, 64), seed=3)
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, conv, dt, _ = cpu_baseline(batch, sample, cores)
        if i >= args.warmup:
            vals.append(v); times.append(dt)
    value = float(np.mean(vals))
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=float(np.mean(times) * 1e3), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="BASELINE config 3 shape: mixed scenarios 1/2, 16 circles, N=20, cold start",
                            batch_per_step=sample, horizon=20, n_obs=16, nlp="clean"),
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


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536, help="instances per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contexts", type=int, default=6, help="concurrent solver contexts (streams) per GPU")
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--ref-sample", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from mobile_manipulator_mpc_b200 import _abi
    from mobile_manipulator_mpc_b200._lib import lib, check
    from mobile_manipulator_mpc_b200.batch_solver import BatchSolver
    import ctypes as C

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
    B = args.batch
    T = max(1, args.contexts)
    batch = make_workload(B, seed=3 + 1000 * rank)
    N, n_obs, n_pl = batch["N"], batch["n_obs"], batch["n_pl"]
    # T solver contexts (handle + stream + workspace each), driven by one host thread each: the thin
    # tail of slow instances of one batch overlaps the bulk of the next batch
    ctx = []
    for t in range(T):
        S = BatchSolver(N=N, dt=batch["dt"], n_obs=n_obs, n_pl=n_pl, B_max=B, device=local)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            dev_in = S.to_device(batch)
            out = S.solve_device(dev_in)
        ctx.append(dict(S=S, stream=st, dev_in=dev_in, out=out))
    torch.cuda.synchronize()
    S = ctx[0]["S"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(nsteps, contexts, host=False):
        """nsteps solves of the batch spread round-robin over `contexts` concurrent solver contexts."""
        outs = [None] * contexts
        nxt = [0]
        lock = threading.Lock()

        def work(t):
            c = ctx[t]
            torch.cuda.set_device(local)
            with torch.cuda.stream(c["stream"]):
                while True:
                    with lock:   # a context takes the next step when it is free: exactly nsteps solves, balanced
                        if nxt[0] >= nsteps:
                            break
                        nxt[0] += 1
                    outs[t] = c["S"].solve_host(batch) if host else c["S"].solve_device(c["dev_in"], out=c["out"])
        if contexts == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(t,)) for t in range(contexts)]
            [x.start() for x in th]; [x.join() for x in th]
        return outs

    # ---- device-resident timing (value): K steps over T contexts ----
    run_steps(args.warmup, T)
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
    out = ctx[0]["out"]
    conv = int((out["status"] == 0).sum().item())
    iters_np = out["iters"].cpu().numpy()
    iters_sum = int(iters_np.sum())
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    c = torch.tensor([conv, iters_sum, B], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    total_ms_max = float(t.item())
    conv_all, iters_all, B_all = (float(v) for v in c.tolist())
    value = conv_all * args.steps / (total_ms_max * 1e-3)

    # ---- one context alone, CUDA events around every launch: latency of one batched solve and the
    #      per-kernel durations the roofline is computed from ----
    S.set_profile(True)
    lat_ms, phase_ms, phase_ln, rounds = [], {}, {}, 0
    for i in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx[0]["stream"]):
            a.record(); S.solve_device(ctx[0]["dev_in"], out=ctx[0]["out"]); b.record()
        torch.cuda.synchronize()
        lat_ms.append(a.elapsed_time(b))
        pm, pl, rounds = S.phase_times()
        for k in pm:
            phase_ms[k] = phase_ms.get(k, 0.0) + pm[k] / args.steps
            phase_ln[k] = pl[k]
    S.set_profile(False)
    # p50 latency of ONE instance (B = 1, BASELINE config 1: demo scenario 1), device-resident
    lat1 = None
    if rank == 0:
        from mobile_manipulator_mpc_b200 import scenarios
        b1 = scenarios.make_batch(1, 1)
        S1 = BatchSolver(N=b1["N"], dt=b1["dt"], n_obs=b1["n_obs"], n_pl=b1["n_pl"], B_max=1, device=local)
        d1 = S1.to_device(b1); o1 = S1.solve_device(d1); torch.cuda.synchronize()
        l1 = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); S1.solve_device(d1, out=o1); b.record(); torch.cuda.synchronize()
            l1.append(a.elapsed_time(b))
        lat1 = float(np.median(l1))
        S1.close()
    # the same batch as the reference's NLP to the letter (MMPC_MODE_REFERENCE: stale plane columns, terminal rows on s[N-1]),
    # one context, two timed solves: reported next to the headline, which is the stage-separable ("clean") NLP
    ref_nlp = None
    if rank == 0:
        from mobile_manipulator_mpc_b200 import _abi
        SR = BatchSolver(N=N, dt=batch["dt"], n_obs=n_obs, n_pl=n_pl, B_max=B, device=local, mode=_abi.MODE_REFERENCE)
        dR = SR.to_device(batch); oR = SR.solve_device(dR); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); SR.solve_device(dR, out=oR); SR.solve_device(dR, out=oR); b.record(); torch.cuda.synchronize()
        convR = int((oR["status"] == 0).sum().item())
        ref_nlp = dict(single_context_value=convR * 2 / (a.elapsed_time(b) * 1e-3), ms_per_step=a.elapsed_time(b) / 2,
                       converged_fraction=convR / B)
        SR.close()
    barrier()

    # ---- end to end through the C ABI with host buffers (pinned staging, H2D, solve, D2H per step) ----
    run_steps(min(T, args.steps), T, host=True)
    barrier()
    t0 = time.perf_counter()
    outs_h = run_steps(args.steps, T, host=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    out_h = next(o for o in outs_h if o is not None)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    ce = torch.tensor([float((out_h["status"] == 0).sum())], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX); dist.all_reduce(ce, op=dist.ReduceOp.SUM)
    e2e_value = float(ce.item()) * args.steps / float(te.item())
    h2d = sum(np.asarray(batch[k]).nbytes for k in ("x_init", "x_ref", "u_ref", "u_last", "circles", "planes", "n_pl_inst"))
    d2h = sum(v.nbytes for v in out_h.values())

    # ---- NCCL gather of the per-instance result (u0, status): statistics only, outside the solve ----
    gathered = None
    if dist is not None:
        u0 = out["U"][:, 0, :].contiguous()
        allu0 = torch.empty((world * B, 5), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allu0, u0)
        gathered = int(allu0.shape[0])

    if rank == 0:
        fp64 = C.c_double()
        check(lib().mmpc_bench_fp64(local, C.byref(fp64)))
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except OSError:
            peak_src = "fallback 6650 GB/s (of fallback)"
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # units = instances one launch series processed: every instance is evaluated and factorised
        # iters+1 times, stepped and tried iters times (extra line-search trials are not credited)
        model = phase_model(N, n_obs, n_pl)
        units = dict(eval=iters_sum + B, solve=iters_sum + B, step=iters_sum, trial=iters_sum)
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except OSError:
            pass
        step_ms = float(np.mean(lat_ms))
        phases = {}
        for k in ("eval", "solve", "step", "trial"):
            ms = phase_ms.get(k, 0.0)
            gbs = model[k]["bytes"] * units[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            tfs = model[k]["flops"] * units[k] / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            phases[k] = dict(ms=ms, share=ms / step_ms, launches=phase_ln.get(k, 0), bytes_per_instance=model[k]["bytes"],
                             hbm_gbs=gbs, hbm_frac=gbs / hbm_peak, fp64_tflops=tfs, fp64_frac=tfs / float(fp64.value))
        for k in ("compact", "ctrl_step", "ctrl_trial", "init"):
            phases[k] = dict(ms=phase_ms.get(k, 0.0), share=phase_ms.get(k, 0.0) / step_ms, launches=phase_ln.get(k, 0))
        dom = max(("eval", "solve", "step", "trial"), key=lambda k: phases[k]["ms"])
        kname = {"eval": "mmpc::staged_eval_kernel", "solve": "mmpc::staged_solve_team_kernel",
                 "step": "mmpc::staged_step_kernel", "trial": "mmpc::staged_trial_kernel"}[dom]
        nl = max(1, phases[dom]["launches"])
        tr = traffic.get(dom, {}).get("dram_bytes_per_instance")
        work_flops = float(sum(flops_per_iteration(N, n_obs, int(p)) * int(i) for p, i in zip(batch["n_pl_inst"], iters_np)))
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload="BASELINE config 3 shape: batch %d per GPU, mixed scenarios 1/2 (n_pl 3/2), "
                                         "16 random circles, N=20, dt=0.1, cold start (u_last=0)" % B,
                                batch_per_gpu=B, horizon=N, n_obs=n_obs, n_pl_max=n_pl, nlp="clean", solver="staged",
                                contexts=T,
                                l2_policy="solver state %.1f GB per context and inputs+outputs %.0f MB per step exceed the 126 MB L2"
                                          % (S.workspace_bytes() / 1e9, (h2d + d2h) / 1e6),
                                converged_fraction=conv_all / B_all, mean_iterations=iters_all / B_all,
                                rounds=rounds, single_context_ms_per_step=step_ms,
                                p50_batched_solve_latency_ms=float(np.median(lat_ms)), p50_single_instance_latency_ms=lat1, reference_nlp=ref_nlp,
                                single_context_value=conv / (step_ms * 1e-3), wall_ms_timed_region=wall_ms),
                    clocks=clocks, gpu_launches=int(launches),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h)),
                    roofline=dict(bound="hbm", achieved=phases[dom]["hbm_gbs"], peak=hbm_peak, unit="GB/s",
                                  frac=phases[dom]["hbm_frac"],
                                  traffic=None if tr is None else tr * units[dom] / nl,
                                  kernel=kname, kernel_ms=phases[dom]["ms"] / nl, kernel_launches=nl,
                                  kernel_share_of_step=phases[dom]["share"], peak_source=peak_src,
                                  algorithmic_bytes_per_launch=model[dom]["bytes"] * units[dom] / nl,
                                  measured_in="single-context pass of %d steps, CUDA events around every launch" % args.steps,
                                  fp64=dict(achieved=work_flops / (step_ms * 1e-3) / 1e12, peak=float(fp64.value), unit="TFLOP/s",
                                            frac=work_flops / (step_ms * 1e-3) / 1e12 / float(fp64.value),
                                            note="whole solve, SURVEY 8(d) flop count; peak = same-run DFMA micro-benchmark"),
                                  phases=phases))
        if gathered is not None:
            line["config"]["nccl_gathered_rows"] = gathered
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, cconv, cdt, _ = cpu_baseline(batch, args.cpu_sample, cores)
            v1, _, cdt1, _ = cpu_baseline(batch, min(256, args.cpu_sample), 1)   # serial leg
            line["cpu_baseline"] = dict(value=v, unit=UNIT, cores=cores, kind="port",
                                        sample="first %d instances of rank 0's batch, oracle/mmpc_oracle.c on %d threads, %.1f s"
                                               % (args.cpu_sample, cores, cdt),
                                        serial_value=v1, serial_sample="first %d instances on 1 thread, %.1f s" % (min(256, args.cpu_sample), cdt1))
        _emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
