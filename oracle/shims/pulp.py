"""Minimal stand-in for the PuLP modelling API, lowering to scipy.optimize.milp (HiGHS).

ORACLE INFRASTRUCTURE ONLY.  PuLP / Gurobi / CBC are not installed in the build
image; this shim lets /root/reference/saturn/solver/milp.py run
*unmodified* so that its outputs can be recorded as golden fixtures
(oracle/gen_golden.py).  Only the API surface milp.py touches is provided:
LpProblem, LpVariable, lpSum, LpMinimize, GUROBI_CMD, PULP_CBC_CMD and affine
arithmetic / comparisons on variables.  It is written from PuLP's documented
behaviour, not from its source.
"""
import numpy as np

LpMinimize = 1
LpMaximize = -1

__all__ = ["LpProblem", "LpVariable", "lpSum", "LpMinimize", "LpMaximize", "GUROBI_CMD",
           "PULP_CBC_CMD", "LpAffineExpression", "LpConstraint"]


class LpAffineExpression:
    __slots__ = ("terms", "const")

    def __init__(self, terms=None, const=0.0):
        self.terms = terms if terms is not None else {}
        self.const = float(const)

    @staticmethod
    def of(x):
        if isinstance(x, LpAffineExpression):
            return x
        if isinstance(x, LpVariable):
            return LpAffineExpression({x: 1.0}, 0.0)
        return LpAffineExpression({}, float(x))

    def copy(self):
        return LpAffineExpression(dict(self.terms), self.const)

    def _iadd(self, other, sign=1.0):
        other = LpAffineExpression.of(other)
        for v, c in other.terms.items():
            self.terms[v] = self.terms.get(v, 0.0) + sign * c
        self.const += sign * other.const
        return self

    def __add__(self, o):
        return self.copy()._iadd(o, 1.0)

    __radd__ = __add__

    def __sub__(self, o):
        return self.copy()._iadd(o, -1.0)

    def __rsub__(self, o):
        return (self * -1.0)._iadd(o, 1.0)

    def __neg__(self):
        return self * -1.0

    def __mul__(self, k):
        if isinstance(k, (LpVariable, LpAffineExpression)):
            k = LpAffineExpression.of(k)
            if k.terms and self.terms:
                raise TypeError("non-linear product")
            if k.terms:
                return k * self.const
            k = k.const
        k = float(k)
        return LpAffineExpression({v: c * k for v, c in self.terms.items()}, self.const * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self * (1.0 / float(k))

    def __le__(self, o):
        return LpConstraint(self - o, -1)

    def __ge__(self, o):
        return LpConstraint(self - o, 1)

    def __eq__(self, o):
        return LpConstraint(self - o, 0)

    __hash__ = None


class LpConstraint:
    """expr (sense) 0 with sense in {-1: <=, 0: ==, 1: >=}."""
    __slots__ = ("expr", "sense")

    def __init__(self, expr, sense):
        self.expr = expr
        self.sense = sense


class LpVariable:
    _counter = 0

    def __init__(self, name, lowBound=None, upBound=None, cat="Continuous"):
        self.name = name
        self.cat = cat
        if cat == "Binary":
            lowBound, upBound = 0, 1
        self.lowBound = lowBound
        self.upBound = upBound
        self.varValue = None
        self._init = None
        LpVariable._counter += 1
        self._id = LpVariable._counter

    def __hash__(self):
        return self._id

    def setInitialValue(self, v):
        self._init = v

    def value(self):
        return self.varValue

    def _e(self):
        return LpAffineExpression({self: 1.0}, 0.0)

    def __add__(self, o):
        return self._e() + o

    __radd__ = __add__

    def __sub__(self, o):
        return self._e() - o

    def __rsub__(self, o):
        return LpAffineExpression.of(o) - self._e()

    def __neg__(self):
        return self._e() * -1.0

    def __mul__(self, k):
        return self._e() * k

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self._e() / k

    def __le__(self, o):
        return self._e() <= o

    def __ge__(self, o):
        return self._e() >= o

    def __eq__(self, o):
        return self._e() == o


def lpSum(items):
    acc = LpAffineExpression()
    for it in items:
        acc._iadd(it, 1.0)
    return acc


class _Solver:
    def __init__(self, timeLimit=None, threads=None, warmStart=False, options=None, msg=False, **kw):
        self.timeLimit = timeLimit
        self.threads = threads
        self.options = options


class GUROBI_CMD(_Solver):
    pass


class PULP_CBC_CMD(_Solver):
    pass


class LpProblem:
    def __init__(self, name="prob", sense=LpMinimize):
        self.name = name
        self.sense = sense
        self.constraints = []
        self.objective = None
        self.status = None
        self.info = {}

    def __iadd__(self, c):
        if isinstance(c, LpConstraint):
            self.constraints.append(c)
        elif isinstance(c, (LpAffineExpression, LpVariable)):
            self.objective = LpAffineExpression.of(c)
        else:
            raise TypeError("cannot add %r to a problem" % (c,))
        return self

    def setObjective(self, e):
        self.objective = LpAffineExpression.of(e)

    def numVariables(self):
        return len(self._collect())

    def numConstraints(self):
        return len(self.constraints)

    def _collect(self):
        seen = {}
        for c in self.constraints:
            for v in c.expr.terms:
                seen.setdefault(v, len(seen))
        if self.objective is not None:
            for v in self.objective.terms:
                seen.setdefault(v, len(seen))
        return seen

    def solve(self, solver=None):
        from scipy.optimize import milp, LinearConstraint, Bounds
        from scipy.sparse import csr_matrix
        idx = self._collect()
        n = len(idx)
        rows, cols, vals, lo, hi = [], [], [], [], []
        for r, c in enumerate(self.constraints):
            for v, a in c.expr.terms.items():
                if a != 0.0:
                    rows.append(r)
                    cols.append(idx[v])
                    vals.append(a)
            rhs = -c.expr.const
            if c.sense < 0:
                lo.append(-np.inf); hi.append(rhs)
            elif c.sense > 0:
                lo.append(rhs); hi.append(np.inf)
            else:
                lo.append(rhs); hi.append(rhs)
        A = csr_matrix((vals, (rows, cols)), shape=(len(self.constraints), n))
        cvec = np.zeros(n)
        for v, a in self.objective.terms.items():
            cvec[idx[v]] = a * self.sense
        lb = np.full(n, -np.inf)
        ub = np.full(n, np.inf)
        integ = np.zeros(n)
        for v, i in idx.items():
            if v.lowBound is not None:
                lb[i] = v.lowBound
            if v.upBound is not None:
                ub[i] = v.upBound
            if v.cat in ("Binary", "Integer"):
                integ[i] = 1
        opts = {"disp": False}
        if solver is not None and solver.timeLimit is not None:
            opts["time_limit"] = float(solver.timeLimit)
        import os as _os
        if _os.environ.get("ORACLE_MIP_REL_GAP"):      # HiGHS' default relative gap is 1e-4: "optimal" within it
            opts["mip_rel_gap"] = float(_os.environ["ORACLE_MIP_REL_GAP"])
        res = milp(cvec, constraints=LinearConstraint(A, lo, hi), integrality=integ,
                   bounds=Bounds(lb, ub), options=opts)
        self.status = res.status
        self.info = {"status": int(res.status), "message": str(res.message),
                     "mip_gap": getattr(res, "mip_gap", None), "mip_rel_gap_limit": opts.get("mip_rel_gap", 1e-4),
                     "mip_dual_bound": getattr(res, "mip_dual_bound", None), "n_vars": n,
                     "n_cons": len(self.constraints)}
        LpProblem.last_info = self.info
        if res.x is not None:
            for v, i in idx.items():
                v.varValue = float(res.x[i])
        return res.status

    last_info = {}
