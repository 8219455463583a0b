"""Resampling cadence: plan quality and wall time of sb_search_run on the C4 / C3-like tables."""
import sys
sys.path.insert(0, ".")
import numpy as np
from saturn_b200.engine import Engine
from saturn_b200 import synth
eng = Engine(0)
for J, S in ((256, 8), (128, 4)):
    T, valid = synth.synth_table(J, S, 8, seed=0)
    eng.set_table(T)
    chains = eng.search_wave(reduced=True) * 2
    for re_, rounds in ((1, 400), (2, 400), (3, 400), (1, 1600), (2, 1600), (2, 1800)):
        rs = [eng.search_run(chains, rounds, seed=s, reduced=True, resample_every=re_) for s in (1, 2, 3, 4, 5)]
        print(f"J={J} resample_every={re_:2d} rounds={rounds}: mk mean {np.mean([r['makespan'] for r in rs]):.1f} "
              f"min {min(r['makespan'] for r in rs):.1f}  wall {1e3*np.median([r['wall_s'] for r in rs]):.1f} ms", flush=True)
