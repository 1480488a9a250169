#!/usr/bin/env python
"""Golden solutions of MPCBase instances (controllers/mpc_base.py) by SciPy SLSQP on the dense restatement
oracle/nlp_base.py -- independent of the interior-point solvers.   python tests/golden/make_golden_base.py"""
import os, sys
import numpy as np
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mobile_manipulator_mpc_b200 import scenarios   # noqa: E402
from oracle.nlp_base import NLPBase                 # noqa: E402

b = scenarios.make_base_batch(6, N=10, seed=6)
out = dict(x_init=b["x_init"], x_ref=b["x_ref"], u_ref=b["u_ref"], circles=b["circles"], N=b["N"], dt=b["dt"])
costs, U0, Xs = [], [], []
for i in range(6):
    P = NLPBase(b["N"], b["dt"], b["x_init"][i], b["x_ref"][i], b["u_ref"][i], b["circles"][i])
    w0 = P.initial_guess()
    cons = [dict(type="eq", fun=P.eq, jac=lambda w: P.jac(P.eq, w)), dict(type="ineq", fun=lambda w: -P.ineq(w), jac=lambda w: -P.jac(P.ineq, w))]
    r = minimize(P.cost, w0, jac=lambda w: P.jac(lambda v: np.atleast_1d(P.cost(v)), w)[0], constraints=cons, method="SLSQP",
                 options=dict(maxiter=500, ftol=1e-14))
    X, U, s = P.unpack(r.x)
    print(i, r.status, r.message, "cost %.10f" % r.fun, "viol %.2e" % P.violation(r.x), "u0", U[0])
    costs.append(r.fun); U0.append(U[0]); Xs.append(X)
out.update(cost=np.array(costs), u0=np.array(U0), X=np.array(Xs))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "base_N10_slsqp.npz"), **out)
