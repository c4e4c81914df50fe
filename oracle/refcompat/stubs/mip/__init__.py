"""python-mip API subset on scipy.optimize.milp (HiGHS) — test infrastructure.

The reference's AutoFlow solver (easydist/autoflow/solver.py:21,233,271-362,...) needs python-mip
(CBC), which is not vendored and not installed here.  This provider implements
only the surface the solver uses so that the reference can produce plans in this container.
Equal-cost plans may differ from CBC's (plan parity is "unpinned").
"""
import numpy as np
from scipy.optimize import Bounds, LinearConstraint, milp
from scipy.sparse import lil_matrix

BINARY, INTEGER, CONTINUOUS = "B", "I", "C"


class SearchEmphasis:
    DEFAULT, FEASIBILITY, OPTIMALITY = 0, 1, 2


class OptimizationStatus:
    OPTIMAL, INFEASIBLE, FEASIBLE, NO_SOLUTION_FOUND = 0, 1, 3, 5


def _lx(o):
    return o._lx() if isinstance(o, Var) else o


class LinExpr:
    __slots__ = ("c", "k")
    __hash__ = None

    def __init__(self, c=None, k=0.0):
        self.c = c or {}
        self.k = k

    def _add(self, o, s=1.0):
        r = LinExpr(dict(self.c), self.k)
        if isinstance(o, (int, float)):
            r.k += s * o
        else:
            o = _lx(o)
            r.k += s * o.k
            for i, v in o.c.items():
                r.c[i] = r.c.get(i, 0.0) + s * v
        return r

    def __add__(self, o): return self._add(o)
    __radd__ = __add__
    def __sub__(self, o): return self._add(o, -1.0)
    def __rsub__(self, o): return (self * -1.0)._add(o)
    def __mul__(self, f): return LinExpr({i: v * f for i, v in self.c.items()}, self.k * f)
    __rmul__ = __mul__
    def __le__(self, o): return Constr(self - o, "<=")
    def __ge__(self, o): return Constr(self - o, ">=")
    def __eq__(self, o): return Constr(self - o, "==")


class Var:
    def __init__(self, m, i):
        self.m, self.idx, self.x = m, i, None

    def _lx(self): return LinExpr({self.idx: 1.0})
    def __add__(self, o): return self._lx() + o
    __radd__ = __add__
    def __sub__(self, o): return self._lx() - o
    def __rsub__(self, o): return (self._lx() * -1.0) + o
    def __mul__(self, f): return self._lx() * f
    __rmul__ = __mul__
    def __le__(self, o): return self._lx() <= o
    def __ge__(self, o): return self._lx() >= o
    def __eq__(self, o): return self._lx() == o
    def __hash__(self): return id(self)


class Constr:
    def __init__(self, e, s):
        self.e, self.s = e, s


def xsum(it):
    r = LinExpr()
    for t in it:
        if isinstance(t, (int, float)):
            r.k += t
            continue
        t = _lx(t)
        r.k += t.k
        for i, v in t.c.items():
            r.c[i] = r.c.get(i, 0.0) + v
    return r


class _Obj:
    def __init__(self, e):
        self.e = LinExpr(k=e) if isinstance(e, (int, float)) else _lx(e)


def minimize(e):
    return _Obj(e)


class Model:
    def __init__(self, name="", **kw):
        self.vars, self.cons, self.objective = [], [], None
        self.verbose, self.objective_value = 0, None

    def add_var(self, var_type=BINARY, **kw):
        v = Var(self, len(self.vars))
        self.vars.append(v)
        return v

    def __iadd__(self, c):
        assert isinstance(c, Constr)
        self.cons.append(c)
        return self

    def write(self, *a):
        pass

    def optimize(self, **kw):
        n = len(self.vars)
        c = np.zeros(n)
        for i, v in self.objective.e.c.items():
            c[i] = v
        A = lil_matrix((len(self.cons), n))
        lb = np.full(len(self.cons), -np.inf)
        ub = np.full(len(self.cons), np.inf)
        for r, con in enumerate(self.cons):
            for i, v in con.e.c.items():
                A[r, i] = v
            rhs = -con.e.k
            if con.s == "<=":
                ub[r] = rhs
            elif con.s == ">=":
                lb[r] = rhs
            else:
                lb[r] = ub[r] = rhs
        res = milp(c, constraints=LinearConstraint(A.tocsr(), lb, ub), integrality=np.ones(n),
                   bounds=Bounds(0, 1))
        assert res.success, res.message
        for v in self.vars:
            v.x = int(round(res.x[v.idx]))
        self.objective_value = res.fun + self.objective.e.k
        return OptimizationStatus.OPTIMAL
