"""Deterministic synthetic workloads of the benchmark configs (SURVEY §8d / BASELINE.md §3).

T[j][s][g-1] = base_j * beta_js / g ** alpha_js with base ~ LogUniform(600, 36000) s,
alpha ~ U(0.55, 0.95), beta ~ U(1, 1.5); feasibility mask mimicking the example executors
(strategy 0 = single-GPU spilling only at g = 1, the others only at g >= 2, 10 % random failures at
g <= 2) filled with the profiler's failure sentinel 1e8 (PerformanceEvaluator.py:106).
"""
import numpy as np

CONFIGS = {          # BASELINE.json configs -> (J, S, G, seed)
    "C1": (4, 2, 2, 0),
    "C2": (8, 3, 8, 103),
    "C3": (64, 6, 8, 0),
    "C4": (256, 8, 8, 0),
    "C5": (1024, 8, 8, 0),
}


def synth_table(J, S, G, seed=0, masked=True, dtype=np.float32):
    rng = np.random.default_rng(seed)
    base = np.exp(rng.uniform(np.log(600.0), np.log(36000.0), size=J))
    alpha = rng.uniform(0.55, 0.95, size=(J, S))
    beta = rng.uniform(1.0, 1.5, size=(J, S))
    g = np.arange(1, G + 1, dtype=np.float64)
    T = base[:, None, None] * beta[:, :, None] / g[None, None, :] ** alpha[:, :, None]
    valid = np.ones((J, S, G), dtype=bool)
    if masked:
        if S > 1:
            valid[:, 0, 1:] = False
            valid[:, 1:, 0] = False
        oom = rng.uniform(size=(J, S, G)) < 0.10
        oom[:, :, 2:] = False
        valid &= ~oom
        for j in range(J):
            if not valid[j].any():
                valid[j, 0, 0] = True
    T = np.where(valid, T, 1e8)
    return T.astype(dtype), valid
