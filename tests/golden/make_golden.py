"""Generate the golden NLP solutions under tests/golden/ with SciPy SLSQP.

Independent of every solver in this repo: the NLP is the expression-level restatement
oracle/nlp.py (built on oracle/model.py), derivatives are exact complex-step derivatives of
those expressions, and the optimiser is SciPy's SLSQP started from the reference's own initial
guess (controllers/mpc_wholebody_qref.py:302-304).  CasADi/IPOPT is not installable in this
image, so these are *substitute* goldens ("parity unpinned").

Run (takes minutes per instance):   python tests/golden/make_golden.py [name ...]
"""
import os
import sys
import time

import numpy as np
from scipy.optimize import minimize

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from mobile_manipulator_mpc_b200 import scenarios  # noqa: E402
from oracle import nlp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def instances():
    yield "cfg1_N20_reference", scenarios.make_batch(1, 1), "reference"
    yield "cfg1_N10_reference", scenarios.make_batch(1, 1, N=10), "reference"
    yield "manip_N20_reference", scenarios.manipulate_instance(), "reference"
    yield "manip_N20_clean", scenarios.manipulate_instance(), "clean"
    yield "approach_N20_clean", scenarios.approach_instance(), "clean"   # terminal xy equality (flags bit 0)
    b2 = scenarios.make_batch(2, 4)
    for i in range(2):
        one = {k: (v[i:i + 1] if isinstance(v, np.ndarray) else v) for k, v in b2.items()}
        yield f"cfg2_i{i}_reference", one, "reference"


def solve_slsqp(P, w0=None):
    w0 = P.initial_guess() if w0 is None else w0
    cons = [dict(type="eq", fun=P.eq, jac=lambda w: P.jac(P.eq, w)),
            dict(type="ineq", fun=lambda w: -P.ineq(w), jac=lambda w: -P.jac(P.ineq, w))]
    res = minimize(P.cost, w0, jac=P.cost_grad, constraints=cons, method="SLSQP",
                   options=dict(maxiter=400, ftol=1e-13))
    return res


def main(names):
    for name, batch, mode in instances():
        if names and name not in names:
            continue
        P = nlp.from_batch(batch, 0, mode)
        t = time.time()
        res = solve_slsqp(P)
        X, U, s = P.unpack(res.x)
        out = dict(name=name, mode=mode, cost=res.fun, X=X, U=U, s=s, status=res.status, nit=res.nit,
                   violation=P.violation(res.x), active=P.active_rows(res.x), seconds=time.time() - t)
        for k, v in batch.items():
            out["in_" + k] = v
        np.savez(os.path.join(HERE, name + ".npz"), **out)
        print(f"{name}: cost {res.fun:.10f} u0 {U[0]} status {res.status} nit {res.nit} "
              f"viol {out['violation']:.2e} {out['seconds']:.0f}s", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
