"""BatchSolver: thin Python host over the C ABI (include/mmpc.h) for B independent instances of
the whole-body MPC NLP of MPCWholeBody (controllers/mpc_wholebody_qref.py:142-331 in the reference).

Two call paths, both into the same CUDA kernel:
  * ``solve_host(batch)``    NumPy in / NumPy out (mmpc_solve_host: pinned staging, H2D, solve, D2H)
  * ``solve_device(batch)``  torch CUDA tensors in / out, asynchronous on torch's current stream
torch is used for device memory and streams only.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._lib import MmpcError, check, lib

_IN_KEYS = ("x_init", "x_ref", "u_ref", "u_last", "u_guess", "circles", "planes")
_OUT_KEYS = ("U", "X", "s", "cost", "kkt", "iters", "status")


def _diag(M, n, name):
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 2:
        if M.shape != (n, n):
            raise ValueError(f"{name} must be {n}x{n}")
        if np.any(M - np.diag(np.diag(M)) != 0):
            raise NotImplementedError(f"{name}: only diagonal weight matrices are supported "
                                      "(every use in the reference is diagonal)")
        M = np.diag(M)
    M = np.ascontiguousarray(M.reshape(-1))
    if M.size != n:
        raise ValueError(f"{name} must have {n} diagonal entries")
    return M


class BatchSolver:
    def __init__(self, N=20, dt=0.1, n_obs=3, n_pl=3, B_max=1, mode=_abi.MODE_REFERENCE, device=0,
                 obs_per_stage=False, cfg=None, kernel=None, **overrides):
        if cfg is None:
            cfg = _abi.default_config(N=N, dt=dt, n_obs=n_obs, n_pl=n_pl, mode=mode)
            cfg.obs_per_stage = int(bool(obs_per_stage))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.B_max = int(B_max)
        self.device = int(device)
        self._h = C.c_void_p()
        check(lib().mmpc_create(C.byref(cfg), self.B_max, self.device, C.byref(self._h)))
        if kernel is not None:
            self.set_kernel(kernel)

    def set_kernel(self, kernel):
        """'auto' ('resident' for at most four instances per SM when the instance fits in shared memory, else 'staged') | 'staged' (phase kernels over active lists, one
        CUDA graph per solve) | 'resident' (one thread block per instance, state in shared memory) | A/B references:
        'staged_hostloop' (the host sequences the rounds), 'staged_thread', 'staged_unfused', 'staged_fat'."""
        k = {"auto": _abi.KERNEL_AUTO, "staged": _abi.KERNEL_STAGED, "staged_thread": _abi.KERNEL_STAGED_THREAD,
             "staged_unfused": _abi.KERNEL_STAGED_UNFUSED, "staged_fat": _abi.KERNEL_STAGED_FAT,
             "staged_hostloop": _abi.KERNEL_STAGED_HOSTLOOP, "resident": _abi.KERNEL_RESIDENT}.get(kernel, kernel)
        check(lib().mmpc_set_kernel(self._h, int(k)))

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().mmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def N(self):
        return self.cfg.N

    def occupancy(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib().mmpc_occupancy(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(sm_count=a.value, instances_per_sm=b.value, smem_bytes=c.value)

    PHASES = ("compact", "eval", "solve", "step", "ctrl_step", "trial", "ctrl_trial", "init", "pose")

    def set_profile(self, on=True):
        check(lib().mmpc_set_profile(self._h, int(bool(on))))

    def phase_times(self):
        """Device ms and launches per phase of the last solve (needs set_profile(True)), and its rounds."""
        ms = (C.c_double * 9)(); ln = (C.c_int64 * 9)(); rounds = C.c_int32()
        check(lib().mmpc_phase_times(self._h, ms, ln, C.byref(rounds)))
        return ({p: ms[i] for i, p in enumerate(self.PHASES)}, {p: int(ln[i]) for i, p in enumerate(self.PHASES)},
                int(rounds.value))

    def workspace_bytes(self):
        """Device bytes of the staged solver's state for B_max instances (allocated on first solve)."""
        v = C.c_int64()
        check(lib().mmpc_workspace_bytes(self._h, C.byref(v)))
        return int(v.value)

    def launch_count(self):
        return int(lib().mmpc_launch_count(self._h))

    def last_solver(self):
        """'resident' or 'staged': what the last solve ran on (AUTO picks by batch size); None before the first solve."""
        return {_abi.KERNEL_RESIDENT: "resident", _abi.KERNEL_STAGED: "staged"}.get(int(lib().mmpc_last_solver(self._h)))

    # -- weights: MPCWholeBody.setWeight (:119-139) ----------------------------------------------
    def set_weights(self, Q=None, R=None, P=None, S=None, W=None):
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        Qd = None if Q is None else _diag(Q, 9, "Q")
        Pd = None if P is None else _diag(P, 9, "P")
        Rd = None if R is None else _diag(R, 5, "R")
        Wd = None if W is None else _diag(W, 5, "W")
        Sv = float("nan") if S is None else float(np.asarray(S).reshape(-1)[0])
        check(lib().mmpc_set_weights(self._h, ptr(Qd), ptr(Pd), ptr(Rd), ptr(Wd), Sv))
        for name, v in (("Qd", Qd), ("Pd", Pd), ("Rd", Rd), ("Wd", Wd)):
            if v is not None:
                getattr(self.cfg, name)[:] = list(v)
        if S is not None:
            self.cfg.S = Sv

    # -- host path --------------------------------------------------------------------------------
    def _shapes(self, B):
        N, c = self.cfg.N, self.cfg
        circ = (B, N + 1, c.n_obs, 3) if c.obs_per_stage else (B, c.n_obs, 3)
        return dict(x_init=(B, 9), x_ref=(B, N + 1, 9), u_ref=(B, N, 5), u_last=(B, N, 5), u_guess=(B, N, 5),
                    circles=circ, planes=(B, c.n_pl, 6))

    def host_outputs(self, B, want=("X", "s", "cost", "kkt", "iters"), pinned=False):
        """Result arrays of solve_host; ``pinned``: page-locked (torch) so that mmpc_solve_host copies into them by DMA."""
        N = self.cfg.N
        shapes = dict(U=((B, N, 5), np.float64), status=((B,), np.int32), X=((B, N + 1, 9), np.float64), s=((B, N + 1), np.float64),
                      cost=((B,), np.float64), kkt=((B,), np.float64), iters=((B,), np.int32))
        out = {}
        for k in ("U", "status") + tuple(want):
            shp, dt = shapes[k]
            if pinned:
                import torch
                t = torch.empty(shp, dtype=torch.float64 if dt == np.float64 else torch.int32).pin_memory()
                out.setdefault("_keep", []).append(t)
                out[k] = t.numpy()
            else:
                out[k] = np.empty(shp, dt)
        return out

    def solve_host(self, batch, want=("X", "s", "cost", "kkt", "iters"), out=None):
        B = int(np.asarray(batch["x_init"]).shape[0])
        if B > self.B_max:
            raise ValueError(f"B={B} exceeds B_max={self.B_max}")
        shp = self._shapes(B)
        keep = []
        ptrs = []
        for k in _IN_KEYS:
            a = batch.get(k)
            if a is None or (k == "circles" and self.cfg.n_obs == 0) or (k == "planes" and self.cfg.n_pl == 0):
                ptrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != shp[k]:
                raise ValueError(f"{k}: expected shape {shp[k]}, got {a.shape}")
            keep.append(a)
            ptrs.append(a.ctypes.data_as(C.c_void_p))
        npl = batch.get("n_pl_inst")
        if npl is not None:
            npl = np.ascontiguousarray(npl, dtype=np.int32); keep.append(npl)
        flags = batch.get("flags")
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8); keep.append(flags)
        xg = batch.get("x_guess")
        if xg is not None:
            xg = np.ascontiguousarray(xg, dtype=np.float64); keep.append(xg)
            if xg.shape != (B, self.cfg.N + 1, 9):
                raise ValueError(f"x_guess: expected shape {(B, self.cfg.N + 1, 9)}, got {xg.shape}")
        bi = _abi.MmpcBatchIn(*ptrs, _abi.ptr(npl), _abi.ptr(flags), _abi.ptr(xg))
        if out is None:
            out = self.host_outputs(B, want)
        bo = _abi.MmpcBatchOut(*[_abi.ptr(out.get(k)) for k in _OUT_KEYS])
        check(lib().mmpc_solve_host(self._h, B, C.byref(bi), C.byref(bo)))
        return out

    # -- device path --------------------------------------------------------------------------------
    def solve_device(self, batch, out=None, want=("X", "s", "cost", "kkt", "iters")):
        """torch CUDA float64 tensors in ``batch``; returns dict of torch tensors (async)."""
        import torch
        x0 = batch["x_init"]
        B = int(x0.shape[0])
        if B > self.B_max:
            raise ValueError(f"B={B} exceeds B_max={self.B_max}")
        dev = x0.device
        if dev.type != "cuda" or dev.index != self.device:
            raise MmpcError(f"solve_device needs CUDA tensors on device {self.device}")
        shp = self._shapes(B)
        ptrs = []
        for k in _IN_KEYS:
            a = batch.get(k)
            if a is None or (k == "circles" and self.cfg.n_obs == 0) or (k == "planes" and self.cfg.n_pl == 0):
                ptrs.append(None)
                continue
            if a.dtype != torch.float64 or not a.is_contiguous() or tuple(a.shape) != shp[k]:
                raise ValueError(f"{k}: need contiguous float64 tensor of shape {shp[k]}")
            ptrs.append(C.c_void_p(a.data_ptr()))
        npl, flags = batch.get("n_pl_inst"), batch.get("flags")
        xg = batch.get("x_guess")
        bi = _abi.MmpcBatchIn(*ptrs, None if npl is None else C.c_void_p(npl.data_ptr()),
                              None if flags is None else C.c_void_p(flags.data_ptr()),
                              None if xg is None else C.c_void_p(xg.data_ptr()))
        N = self.cfg.N
        if out is None:
            out = dict(U=torch.empty((B, N, 5), dtype=torch.float64, device=dev),
                       status=torch.empty(B, dtype=torch.int32, device=dev))
            full = dict(X=(B, N + 1, 9), s=(B, N + 1), cost=(B,), kkt=(B,))
            for k in want:
                out[k] = (torch.empty(B, dtype=torch.int32, device=dev) if k == "iters"
                          else torch.empty(full[k], dtype=torch.float64, device=dev))
        bo = _abi.MmpcBatchOut(*[None if out.get(k) is None else C.c_void_p(out[k].data_ptr()) for k in _OUT_KEYS])
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib().mmpc_solve(self._h, B, C.byref(bi), C.byref(bo), stream))
        return out

    def to_device(self, batch):
        import torch
        dev = torch.device("cuda", self.device)
        out = {}
        for k, v in batch.items():
            if isinstance(v, np.ndarray) and (k in _IN_KEYS or k == "x_guess"):
                out[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
            elif k == "n_pl_inst" and v is not None:
                out[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.int32)).to(dev)
            elif k == "flags" and v is not None:
                out[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.uint8)).to(dev)
        return out

    # -- model evaluation, shift, plant (device tensors) ------------------------------------------------
    def eval_model(self, x, u=None, circles=None, planes=None):
        import torch
        M = int(x.shape[0])
        dev = x.device
        c = self.cfg
        f = torch.empty((M, 9), dtype=torch.float64, device=dev)
        fk = torch.empty((M, 10), dtype=torch.float64, device=dev)
        rows = None
        if (c.n_obs == 0 or circles is not None) and (c.n_pl == 0 or planes is not None):
            rows = torch.empty((M, c.n_obs + 4 + 6 * c.n_pl), dtype=torch.float64, device=dev)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(lib().mmpc_eval_model(self._h, M, p(x), p(u), p(circles), p(planes), p(f), p(fk), p(rows), stream))
        return f, fk, rows

    def shift(self, U):
        import torch
        ug = torch.empty_like(U)
        stream = C.c_void_p(torch.cuda.current_stream(U.device).cuda_stream)
        check(lib().mmpc_shift(self._h, int(U.shape[0]), C.c_void_p(U.data_ptr()), C.c_void_p(ug.data_ptr()), stream))
        return ug

    def window(self, x, x_glob, u_glob=None, idx=(0, 1), want_index=False):
        """calcLocalRefTraj on device: x [B,9], x_glob [M,9] (shared) or [B,M,9]; returns x_ref [B,N+1,9], u_ref [B,N,5]."""
        import torch
        B, N = int(x.shape[0]), self.cfg.N
        shared = x_glob.dim() == 2
        M = int(x_glob.shape[-2])
        xr = torch.empty((B, N + 1, 9), dtype=torch.float64, device=x.device)
        ur = torch.empty((B, N, 5), dtype=torch.float64, device=x.device)
        ist = torch.empty(B, dtype=torch.int32, device=x.device) if want_index else None
        mask = sum(1 << int(i) for i in idx)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        check(lib().mmpc_window(self._h, B, M, mask, int(shared), p(x), p(x_glob), p(u_glob), p(xr), p(ur), p(ist), stream))
        return (xr, ur, ist) if want_index else (xr, ur)

    def plant_step(self, x, u0):
        import torch
        xn = torch.empty_like(x)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        check(lib().mmpc_plant_step(self._h, int(x.shape[0]), C.c_void_p(x.data_ptr()), C.c_void_p(u0.data_ptr()),
                                    C.c_void_p(xn.data_ptr()), stream))
        return xn

    def ik(self, q_guess, target):
        """Batched inverse_transformation (robot_models/manipulator_3DoF.py:79-133) on device: q_guess [B,3],
        target [B,3] = (x, 0, z) in the arm frame; returns (q [B,3], status [B] int32: 0 reached, 1 unreachable)."""
        import torch
        B = int(q_guess.shape[0])
        q = torch.empty((B, 3), dtype=torch.float64, device=q_guess.device)
        st = torch.empty(B, dtype=torch.int32, device=q_guess.device)
        stream = C.c_void_p(torch.cuda.current_stream(q_guess.device).cuda_stream)
        check(lib().mmpc_ik(self._h, B, C.c_void_p(q_guess.contiguous().data_ptr()), C.c_void_p(target.contiguous().data_ptr()),
                            C.c_void_p(q.data_ptr()), C.c_void_p(st.data_ptr()), stream))
        return q, st

    def episode_update(self, io, B, M, n_manip):
        """Interface.stateMachineUpdate for B episodes (mmpc_episode_update); ``io``: dict of device tensors named like
        the fields of MmpcEpisodeIO (missing optional ones -> NULL)."""
        import torch
        e = _abi.MmpcEpisodeIO()
        for name, _ in _abi.MmpcEpisodeIO._fields_:
            t = io.get(name)
            setattr(e, name, None if t is None else t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(io["x"].device).cuda_stream)
        check(lib().mmpc_episode_update(self._h, int(B), int(M), int(n_manip), C.byref(e), stream))
