"""``controllers.mpc_wholebody_qref`` for the CPU tier: the drop-in MPCWholeBody with its one call into the CUDA
library replaced by the CPU oracle (oracle/mmpc_oracle.c), so the reference's Interface and demo can be driven in a
container without a GPU.  Every solve call is recorded (inputs, weights, flag, outputs) for the replay fixture of the
GPU tier.  TEST INFRASTRUCTURE ONLY -- the product class has no such backend."""
import numpy as np

from mobile_manipulator_mpc_b200 import _abi
from mobile_manipulator_mpc_b200.controllers import mpc_wholebody_qref as dropin
from oracle import solver as osolver


class MPCWholeBody(dropin.MPCWholeBody):
    TRACE = []          # one dict per solve call, in call order

    def _backend_solve(self, batch, B):
        cfg = _abi.MmpcConfig.from_buffer_copy(self._cfg)
        w = self.weights
        cfg.Qd[:] = list(w["Qd"]); cfg.Pd[:] = list(w["Pd"]); cfg.Rd[:] = list(w["Rd"]); cfg.Wd[:] = list(w["Wd"]); cfg.S = w["S"]
        b = {k: (None if v is None else np.array(v)) for k, v in batch.items()}
        b.update(N=cfg.N, dt=cfg.dt, n_obs=cfg.n_obs, n_pl=cfg.n_pl)
        out = osolver.solve(b, cfg=cfg)
        MPCWholeBody.TRACE.append(dict(x_init=b["x_init"][0].copy(), x_ref=b["x_ref"][0].copy(), u_ref=b["u_ref"][0].copy(),
                                       u_last=b["u_last"][0].copy(), flag=int(self.terminal_xy_eq), Qd=w["Qd"].copy(), Pd=w["Pd"].copy(),
                                       U=out["U"][0].copy(), cost=float(out["cost"][0]), status=int(out["status"][0]),
                                       iters=int(out["iters"][0])))
        return out


def _weighted_cfg(ctrl):
    cfg = _abi.MmpcConfig.from_buffer_copy(ctrl._cfg)
    w = ctrl.weights
    cfg.Qd[:] = list(w["Qd"]); cfg.Pd[:] = list(w["Pd"]); cfg.Rd[:] = list(w["Rd"]); cfg.Wd[:] = list(w["Wd"]); cfg.S = w["S"]
    return cfg


from mobile_manipulator_mpc_b200.controllers import mpc_base as dropin_base  # noqa: E402


class MPCBase(dropin_base.MPCBase):
    """the drop-in MPCBase with the CPU oracle as its solver; ``BACKEND`` may be replaced by the CPU
    emulation of the kernel sources (tests/emu)"""
    BACKEND = staticmethod(lambda b, cfg: osolver.solve(b, cfg=cfg, threads=4))

    def _backend_solve(self, batch, B):
        cfg = _weighted_cfg(self)
        b = {k: (None if v is None else np.array(v)) for k, v in batch.items()}
        b.update(N=cfg.N, dt=cfg.dt, n_obs=cfg.n_obs, n_pl=0)
        return type(self).BACKEND(b, cfg)
