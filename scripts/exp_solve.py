"""solve() wall time on the C4 table: explicit 131072 chains against the wave-rounded default."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from saturn_b200 import Strategy, solve, solver, synth


class Task:
    def __init__(self, name, strategies):
        self.name, self.strategies, self.selected_strategy = name, strategies, None

    def select_strategy(self, s):
        self.selected_strategy = s


T, valid = synth.synth_table(256, 8, 8, seed=0)
tmin = np.where(valid, T, np.inf).min(axis=1)
tasks = [Task("t%d" % j, {g + 1: Strategy("x", g + 1, {}, float(tmin[j, g])) for g in range(8) if np.isfinite(tmin[j, g])})
         for j in range(256)]
solve(tasks, None, rounds=8)
print("wave", solver._engine().search_wave(reduced=True))
for chains in (131072, None, 131072, None) + (None,) * 12:
    t0 = time.perf_counter()
    out = solve(tasks, None, chains=chains, rounds=400)
    dt = time.perf_counter() - t0
    st = solver.last_stats
    print(f"chains={st['chains']} wall {dt*1e3:.1f} ms search {st['search_wall_s']*1e3:.1f} ms cand {st['candidates']:.3e} "
          f"rate {st['candidates']/dt:.3e} mk {out[5]:.1f} rounds {st['rounds']}", flush=True)
