"""CPU ORACLE (test infrastructure — NOT product code): the reference's SPASE MILP, restated.

`/root/reference` does not exist on the GPU box, and the reference's solver path needs PuLP + a
Gurobi/CBC binary that are not in the image.  This module restates the MODEL that
`saturn/solver/milp.py:89-321` builds — same variables, same constraint families, one node of 8
GPUs (`milp.py:57-62`) — directly as a sparse matrix for `scipy.optimize.milp` (HiGHS), so that
"the reference's CPU MILP on the same T" can be timed next to the GPU search in the same run.
It is validated against the fixtures recorded from the UNMODIFIED reference
(`tests/golden/milp_cases.json`, see `tests/test_oracle.py::test_milp_port_matches_reference_runs`).

Deviation, deliberate and recorded: big-M.  The reference uses M = 1e10 (`milp.py:163`), which is
only sound under Gurobi's intFeasTol = 1e-9 and leaks under HiGHS; here
M = H + 1 with the horizon H = sum_t ceil(max_s rt[t][s]) — the substitution the golden "tight_m"
runs use — except in family (iii), whose rows for a NON-selected option of a different GPU count
need up to (8 - 1) * start of slack: those use M3 = 8 * M (sound for every start <= H).

Variables (milp.py line of the original in brackets)
  bss[t][s]  Binary   option s of task t selected                [96-111]
  bna[t]     Binary   task t on node 0 (N = 1 => fixed to 1)     [117-137]
  sta[g][t]  Integer >= 0  start of t on GPU g (0 if not there)  [139-149]
  mk         >= 0     makespan                                   [162]
  tga[t][g]  Binary   t occupies GPU g                           [184-195]
  boa[a][b]  Binary   b runs after a (a != b)                    [263-270]
Constraints
  (i)   mk >= sta[g][t] + rt[t][s] - M(1 - bss[t][s])                       [170-177]
  (ii)  sum_g tga[t][g] == k[t][s] when s selected (two inequalities)       [209-227]
  (iii) sum_g sta[g][t] / k[t][s] == sta[g][t] on occupied GPUs             [233-256]
  (iv)  pairwise exclusion on every GPU, ordered by boa                     [277-319]
"""
from __future__ import annotations

import math
import time

import numpy as np

G = 8  # GPUs per node, milp.py:62


class _Rows:
    def __init__(self):
        self.r, self.c, self.v, self.lo, self.hi = [], [], [], [], []
        self.n = 0

    def add(self, cols, vals, lo, hi):
        self.r.extend([self.n] * len(cols))
        self.c.extend(cols)
        self.v.extend(vals)
        self.lo.append(lo)
        self.hi.append(hi)
        self.n += 1


def build(gpu_time_tuples):
    J = len(gpu_time_tuples)
    M = float(sum(math.ceil(max(rt for (_k, rt) in tup)) for tup in gpu_time_tuples) + 1)
    M3 = 8.0 * M
    off = 0
    bss = []
    for tup in gpu_time_tuples:
        bss.append(list(range(off, off + len(tup))))
        off += len(tup)
    bna = list(range(off, off + J)); off += J
    sta = [[off + g * J + t for t in range(J)] for g in range(G)]; off += G * J
    mk = off; off += 1
    tga = [[off + t * G + g for g in range(G)] for t in range(J)]; off += J * G
    boa = {}
    for a in range(J):
        for b in range(J):
            if a != b:
                boa[(a, b)] = off
                off += 1
    nv = off
    R = _Rows()
    inf = np.inf
    for t in range(J):
        R.add(bss[t], [1.0] * len(bss[t]), 1.0, 1.0)          # one option     [110-111]
        R.add([bna[t]], [1.0], 1.0, 1.0)                       # one node       [136-137]
    for t, tup in enumerate(gpu_time_tuples):
        for s, (k, rt) in enumerate(tup):
            for g in range(G):                                 # (i)
                R.add([mk, sta[g][t], bss[t][s]], [1.0, -1.0, -M], rt - M, inf)
            # (ii): k - M(1-bss) - M(1-bna) <= sum tga <= k + M(1-bss) + M(1-bna)
            cols = tga[t] + [bss[t][s], bna[t]]
            R.add(cols, [1.0] * G + [-M, -M], k - 2 * M, inf)
            R.add(cols, [1.0] * G + [M, M], -inf, k + 2 * M)
            # (iii): target = sum_g sta / k ;  |target - sta[g][t]| <= M(1-tga) + M(1-bss) + M(1-bna)
            for g in range(G):
                coef = {sta[gg][t]: 1.0 / k for gg in range(G)}
                coef[sta[g][t]] -= 1.0
                cols3 = list(coef.keys()) + [tga[t][g], bss[t][s], bna[t]]
                R.add(cols3, list(coef.values()) + [M3, M3, M3], -inf, 3 * M3)
                R.add(cols3, list(coef.values()) + [-M3, -M3, -M3], -3 * M3, inf)
        # off-node: with N = 1 and bna fixed to 1 these rows are vacuous            [226-227]
    for g in range(G):                                         # (iv)
        for t in range(J):
            for p in range(J):
                if p == t:
                    continue
                b = boa[(p, t)]
                for s, (_k, rt) in enumerate(gpu_time_tuples[t]):
                    # sta[t] <= sta[p] - rt_t + M(1-tga_p) + M(1-tga_t) + M*boa[p][t] + M(1-bss[t][s])
                    R.add([sta[g][t], sta[g][p], tga[p][g], tga[t][g], b, bss[t][s]],
                          [1.0, -1.0, M, M, -M, M], -inf, -rt + 3 * M)
                for s, (_k, rt) in enumerate(gpu_time_tuples[p]):
                    # sta[t] >= sta[p] + rt_p - M(1-tga_t) - M(1-tga_p) - M(1-boa[p][t]) - M(1-bss[p][s])
                    R.add([sta[g][t], sta[g][p], tga[t][g], tga[p][g], b, bss[p][s]],
                          [1.0, -1.0, -M, -M, -M, -M], rt - 4 * M, inf)
    integrality = np.ones(nv)
    integrality[mk] = 0
    lb = np.zeros(nv)
    ub = np.ones(nv)
    for g in range(G):
        for t in range(J):
            ub[sta[g][t]] = M
    ub[mk] = np.inf
    idx = dict(bss=bss, bna=bna, sta=sta, mk=mk, tga=tga, boa=boa, nv=nv, M=M, J=J)
    return R, integrality, lb, ub, idx


def model_size(J, S, N=1):
    """Variable / constraint counts of the reference model (SURVEY §8a A2/A3 formulas)."""
    nvars = J * S + J * N + 2 * N * G * J + J * (J - 1) + 1
    ncons = 2 * J + N * G * J * S + 4 * J * N * S + 2 * J * N * S * G + 2 * S * N * G * J * (J - 1)
    return nvars, ncons


def solve(gpu_time_tuples, time_limit=60.0):
    """Returns dict(status, proven_optimal, makespan, start[J], mask[J], opt_idx[J], wall_s, build_s,
    n_vars, n_cons).  makespan is recomputed as max(start + rt) from the decoded plan."""
    from scipy.optimize import Bounds, LinearConstraint, milp
    from scipy.sparse import csr_matrix
    t0 = time.perf_counter()
    R, integrality, lb, ub, idx = build(gpu_time_tuples)
    A = csr_matrix((R.v, (R.r, R.c)), shape=(R.n, idx["nv"]))
    c = np.zeros(idx["nv"])
    c[idx["mk"]] = 1.0
    build_s = time.perf_counter() - t0
    t1 = time.perf_counter()
    res = milp(c, constraints=LinearConstraint(A, R.lo, R.hi), integrality=integrality, bounds=Bounds(lb, ub),
               options={"time_limit": float(time_limit), "disp": False})
    wall = time.perf_counter() - t1
    out = {"status": int(res.status), "proven_optimal": res.status == 0, "wall_s": wall, "build_s": build_s,
           "n_vars": idx["nv"], "n_cons": R.n, "makespan": None}
    if res.x is None:
        return out
    x = res.x
    J = idx["J"]
    start, mask, opt_idx = [], [], []
    for t in range(J):
        o = int(np.argmax([x[v] for v in idx["bss"][t]]))
        m = 0
        first = None
        for g in range(G):
            if round(x[idx["tga"][t][g]]) == 1:
                m |= 1 << g
                if first is None:
                    first = g
        start.append(float(round(x[idx["sta"][first][t]])) if first is not None else 0.0)
        mask.append(m)
        opt_idx.append(o)
    out.update(start=start, mask=mask, opt_idx=opt_idx,
               makespan=max(start[t] + gpu_time_tuples[t][opt_idx[t]][1] for t in range(J)))
    return out
