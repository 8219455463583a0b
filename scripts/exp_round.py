"""Per-kernel time of one search round at a large population (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from saturn_b200.engine import Engine
from saturn_b200.synth import synth_table
eng = Engine(0)
T, valid = synth_table(256, 8, 8, seed=0)
eng.set_table(T)
eng.search_init(1 << 20, seed=0, reduced=True, t_start=5e-4, t_end=1e-6, total_rounds=20)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
eng.search_round(3)
ev[0].record(); eng.search_round(10); ev[1].record(); torch.cuda.synchronize()
print("10 rounds of 1M chains: %.3f ms/round" % (ev[0].elapsed_time(ev[1]) / 10))
