"""Numeric stand-in for the part of the ``casadi`` API that the reference's hot path touches.

TEST INFRASTRUCTURE ONLY.  CasADi 3.6.4 (requirements.txt:9) is not installable in this image, but the
reference's model and NLP code (/root/reference/robot_models/*.py, controllers/mpc_wholebody_qref.py,
interface_wholebody_qref.py, demo_wholebody_qref.py) only *builds expressions* with it.  With this module
first on ``sys.path`` those files import and run UNMODIFIED: every ``ca.*`` call is evaluated eagerly on
NumPy float64 arrays, so ``MPCWholeBody.reset()`` (:142-285) leaves behind the reference's own row list --
every ``opti.subject_to(...)`` in the order it was issued, stale ``self.constr`` columns and the leaked loop
variable ``k`` included -- evaluated at the numeric (X, U, s, parameters) that were fed to the ``Opti``.

Semantics mirrored (CasADi 3.6 documentation / observable behaviour):
  * everything is a 2-D matrix; ``x[i]`` / ``x[a:b]`` on a row (column) vector indexes its elements and keeps
    the orientation; ``x[i, j]`` keeps two dimensions; ``x[i, j] = v`` assigns in place;
  * binary operators broadcast a 1x1 operand; ``mtimes`` is the matrix product (list form = chained);
  * ``if_else(c, a, b)`` selects on the VALUE of c; ``mmax`` is the maximum entry; ``norm_2`` the Euclidean
    norm of a vector; ``fmod`` is C fmod (sign of the dividend);
  * comparison operators build a constraint object ``lhs (op) rhs``; in a boolean context it is its truth
    value (the Interface writes ``if ca.norm_2(...) <= threshold``).
An optional LEADING batch axis evaluates M points in one pass of the reference code: values have shape
``(M, r, c)`` or ``(r, c)`` and broadcast against each other.

Nothing here solves anything: ``Opti.solve`` raises unless a test installs a solve hook.
"""
import math

import numpy as np

pi = math.pi
inf = math.inf


def _raw(x):
    """float64 array with at least the two matrix axes."""
    if isinstance(x, NM):
        return x.v
    if isinstance(x, Constraint):
        return x.truth().astype(float)
    a = np.asarray(x, dtype=float)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(-1, 1)   # casadi: a 1-D NumPy array is a column vector
    return a


class NM:
    """Numeric matrix: ``v`` has shape (r, c) or (M, r, c)."""
    __array_priority__ = 1000

    def __init__(self, v):
        v = _raw(v) if not isinstance(v, np.ndarray) or v.ndim < 2 else v
        self.v = np.array(v, dtype=float) if not isinstance(v, np.ndarray) else v

    # -- shape ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.v.shape[-2:]

    def size1(self):
        return self.v.shape[-2]

    def size2(self):
        return self.v.shape[-1]

    @property
    def T(self):
        return NM(np.swapaxes(self.v, -1, -2))

    def __array__(self, dtype=None, copy=None):
        return np.array(self.v, dtype=dtype or float)

    def __float__(self):
        assert self.v.size == 1, self.v.shape
        return float(self.v.reshape(-1)[0])

    def __bool__(self):
        assert self.v.size == 1, self.v.shape
        return bool(self.v.reshape(-1)[0])

    def __repr__(self):
        return "NM(%r)" % (self.v,)

    def __len__(self):
        return self.v.shape[-2] * self.v.shape[-1] if 1 in self.shape else self.v.shape[-2]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    # -- indexing --------------------------------------------------------------------------------
    @staticmethod
    def _sl(k):
        if isinstance(k, slice):
            return k
        k = int(k)
        return slice(k, k + 1 if k != -1 else None)

    def _key(self, key):
        if isinstance(key, tuple):
            r, c = key
            return (Ellipsis, self._sl(r), self._sl(c))
        r_, c_ = self.shape
        if r_ == 1:
            return (Ellipsis, slice(None), self._sl(key))
        if c_ == 1:
            return (Ellipsis, self._sl(key), slice(None))
        raise NotImplementedError("linear indexing of a general matrix")

    def __getitem__(self, key):
        return NM(self.v[self._key(key)])

    def __setitem__(self, key, val):
        k = self._key(key)
        val = _raw(val)
        tgt = self.v[k]
        if val.ndim > tgt.ndim:   # a batched value stored into an unbatched matrix: grow the batch axis
            self.v = np.broadcast_to(self.v, val.shape[:-2] + self.v.shape[-2:]).copy()
        self.v[k] = val

    # -- arithmetic ------------------------------------------------------------------------------
    def _bin(self, other, f, swap=False):
        a, b = self.v, _raw(other)
        return NM(f(b, a) if swap else f(a, b))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    def __pow__(self, o): return self._bin(o, np.power)
    def __neg__(self): return NM(-self.v)
    def __pos__(self): return self
    def __abs__(self): return NM(np.abs(self.v))
    def __matmul__(self, o): return NM(np.matmul(self.v, _raw(o)))
    def __rmatmul__(self, o): return NM(np.matmul(_raw(o), self.v))

    # NumPy ufuncs applied to an NM (base.py:23 writes np.cos(x[2])) and ndarray (op) NM
    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__" or kw.get("out") is not None:
            return NotImplemented
        if ufunc in (np.less, np.less_equal, np.greater, np.greater_equal, np.equal):
            op = {np.less: "<", np.less_equal: "<=", np.greater: ">", np.greater_equal: ">=", np.equal: "=="}[ufunc]
            return Constraint(NM(_raw(inputs[0])), op, NM(_raw(inputs[1])))
        return NM(ufunc(*[_raw(i) for i in inputs], **kw))

    # numpy-style methods casadi objects also answer to
    def sin(self): return NM(np.sin(self.v))
    def cos(self): return NM(np.cos(self.v))
    def sqrt(self): return NM(np.sqrt(self.v))
    def squeeze(self): return np.squeeze(self.v)
    def reshape(self, *a): return NM(self.v.reshape(*a))

    # -- comparisons build constraints -------------------------------------------------------------
    def __lt__(self, o): return Constraint(self, "<", NM(_raw(o)))
    def __le__(self, o): return Constraint(self, "<=", NM(_raw(o)))
    def __gt__(self, o): return Constraint(self, ">", NM(_raw(o)))
    def __ge__(self, o): return Constraint(self, ">=", NM(_raw(o)))
    def __eq__(self, o): return Constraint(self, "==", NM(_raw(o)))
    __hash__ = None


class Constraint:
    """``lhs (op) rhs`` or ``lo <= mid <= hi`` (Opti.bounded), with numeric operands."""

    def __init__(self, lhs, op, rhs, mid=None):
        self.lhs, self.op, self.rhs, self.mid = lhs, op, rhs, mid

    def truth(self):
        a, b = self.lhs.v, self.rhs.v
        return {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal, "==": np.equal}[self.op](a, b)

    def __bool__(self):
        t = self.truth()
        assert t.size == 1, "truth value of a matrix constraint"
        return bool(t.reshape(-1)[0])


# DM / MX / SX all collapse onto NM
DM = MX = SX = NM


def _un(f):
    def g(x):
        if isinstance(x, NM):
            return NM(f(x.v))
        r = f(np.asarray(x, dtype=float))
        return float(r) if np.ndim(r) == 0 else NM(r)
    return g


sin, cos, tan, sqrt, exp, log, fabs = (_un(f) for f in (np.sin, np.cos, np.tan, np.sqrt, np.exp, np.log, np.abs))
atan2 = lambda a, b: NM(np.arctan2(_raw(a), _raw(b)))


def fmod(a, b):
    if isinstance(a, NM) or isinstance(b, NM):
        return NM(np.fmod(_raw(a), _raw(b)))
    return math.fmod(a, b)


def _cat(args, axis):
    # A 1-D NumPy operand takes the orientation that fits its neighbours.  (CasADi proper reads it as a column; the
    # reference's pybullet-free plant step, robot_models/mobile_manipulator.py:74 ``ca.horzcat(x_base_next, q_next)``
    # with a 1-D ``q_next`` next to a 1x6 row, needs the lenient reading -- see DESIGN.md "refshim".)
    vs = [np.asarray(a, dtype=float).reshape((1, -1) if axis == -1 else (-1, 1))
          if isinstance(a, np.ndarray) and a.ndim == 1 else _raw(a) for a in args]
    nb = max(v.ndim for v in vs)
    if nb > 2:
        bs = np.broadcast_shapes(*[v.shape[:-2] for v in vs])
        vs = [np.broadcast_to(v, bs + v.shape[-2:]) for v in vs]
    return NM(np.concatenate(vs, axis=axis))


def horzcat(*args):
    return _cat(args, -1)


def vertcat(*args):
    return _cat(args, -2)


def mtimes(*args):
    if len(args) == 1:
        args = tuple(args[0])
    out = _raw(args[0])
    for a in args[1:]:
        b = _raw(a)
        out = out * b if (out.shape[-2:] == (1, 1) or b.shape[-2:] == (1, 1)) else np.matmul(out, b)
    return NM(out)


def norm_2(x):
    v = _raw(x)
    assert 1 in v.shape[-2:], "norm_2 of a matrix is the spectral norm; only vectors occur on this path"
    r = np.sqrt(np.sum(v * v, axis=(-1, -2), keepdims=True))
    return NM(r)


def sumsqr(x):
    v = _raw(x)
    return NM(np.sum(v * v, axis=(-1, -2), keepdims=True))


def mmax(x):
    return NM(np.max(_raw(x), axis=(-1, -2), keepdims=True))


def mmin(x):
    return NM(np.min(_raw(x), axis=(-1, -2), keepdims=True))


def fmax(a, b):
    return NM(np.maximum(_raw(a), _raw(b)))


def fmin(a, b):
    return NM(np.minimum(_raw(a), _raw(b)))


def if_else(c, a, b):
    cv = c.truth() if isinstance(c, Constraint) else (_raw(c) != 0)
    if not any(isinstance(t, NM) for t in (a, b)) and cv.size == 1:   # plain floats in, plain float out (angleDiff :92-117)
        return a if bool(cv.reshape(-1)[0]) else b
    return NM(np.where(cv, _raw(a), _raw(b)))


class OptiSol:
    def __init__(self, values):
        self._values = values

    def value(self, x):
        return self._values(x)


class Opti:
    """Records what reset() issues.  ``Opti.FEED`` (class attribute) may hold, for the next instance, the numeric
    values of the variables and parameters in creation order:  {"variable": [arrays...], "parameter": [arrays...]};
    a missing or None entry is a zero matrix."""
    FEED = None
    SOLVE_HOOK = None

    def __init__(self):
        feed = Opti.FEED or {}
        self._feed = {k: list(v) for k, v in feed.items()}
        self.variables, self.parameters, self.constraints = [], [], []
        self.objective = None
        self.solver_name, self.solver_opts = None, None
        self.initial = {}
        self.debug = self

    def _make(self, kind, r, c):
        q = self._feed.get(kind) or []
        val = q.pop(0) if q else None
        if val is None:
            m = NM(np.zeros((r, c)))
        else:
            val = np.array(val, dtype=float)
            assert val.shape[-2:] == (r, c), (kind, val.shape, (r, c))
            m = NM(val)
        (self.variables if kind == "variable" else self.parameters).append(m)
        return m

    def variable(self, r=1, c=1):
        return self._make("variable", r, c)

    def parameter(self, r=1, c=1):
        return self._make("parameter", r, c)

    def set_value(self, p, val):
        val = _raw(val)
        if p.v.ndim == 2 or val.ndim == p.v.ndim:
            p.v[...] = np.broadcast_to(val, p.v.shape)
        else:
            p.v[...] = np.broadcast_to(val, p.v.shape[-2:])

    def set_initial(self, x, val):
        self.initial[id(x)] = np.array(_raw(val))

    def bounded(self, lo, x, hi):
        return Constraint(NM(_raw(lo)), "bounded", NM(_raw(hi)), mid=x if isinstance(x, NM) else NM(_raw(x)))

    def subject_to(self, con):
        assert isinstance(con, Constraint), type(con)
        self.constraints.append(con)

    def minimize(self, cost):
        self.objective = cost

    def solver(self, name, opts=None):
        self.solver_name, self.solver_opts = name, dict(opts or {})

    def value(self, x):
        return np.squeeze(_raw(x))

    def solve(self):
        if Opti.SOLVE_HOOK is None:
            raise RuntimeError("the casadi stand-in evaluates expressions; it does not solve NLPs")
        return Opti.SOLVE_HOOK(self)


def nlpsol(*a, **k):
    raise RuntimeError("the casadi stand-in has no nlpsol")
