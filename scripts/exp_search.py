"""Search-quality experiment (GPU): best makespan vs rounds for a few search settings."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from saturn_b200.engine import Engine
from saturn_b200.search import run_search
from saturn_b200.synth import synth_table

eng = Engine(0)
for (J, S, G) in [(64, 6, 8), (256, 8, 8)]:
    T, valid = synth_table(J, S, G, seed=0)
    eng.set_table(T)
    tmin, _ = eng.reduced_table()
    usable = np.where(tmin < 1e6, tmin, np.inf)
    lb = max(float((usable * np.arange(1, 9)[None, :]).min(axis=1).sum() / 8), float(usable.min(axis=1).max()))
    print("J=%d lower bound (area/8, longest job) %.1f" % (J, lb))
    for name, kw in [
        ("t2e-3 rs8", dict(t_start=2e-3, t_end=1e-5, resample_every=8)),
        ("t2e-3 rs0", dict(t_start=2e-3, t_end=1e-5, resample_every=0)),
        ("t5e-4 rs4", dict(t_start=5e-4, t_end=1e-6, resample_every=4)),
        ("t2e-2 rs8", dict(t_start=2e-2, t_end=1e-4, resample_every=8)),
        ("greedy rs4", dict(t_start=0.0, t_end=0.0, resample_every=4)),
        ("noseed t2e-3 rs8", dict(t_start=2e-3, t_end=1e-5, resample_every=8, heuristic_seeds=False)),
    ]:
        t0 = time.time()
        r = run_search(eng, chains=32768, rounds=200, seed=1, record_history=True, use_dist=False, **kw)
        h = r.history
        pts = [h[i][2] for i in (0, 10, 25, 50, 100, 150, len(h) - 1) if i < len(h)]
        print("  %-18s %s  gap %.2f%%  %.2fs  %.2e cand/s" % (name, " ".join("%.0f" % x for x in pts),
              100 * (r.makespan / lb - 1), time.time() - t0, r.evaluated / r.wall_s))
