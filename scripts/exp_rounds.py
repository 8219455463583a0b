import sys, os, torch
sys.path.insert(0, ".")
from saturn_b200.engine import Engine
from saturn_b200 import synth
eng = Engine(0)
T, valid = synth.synth_table(256, 8, 8, seed=0)
eng.set_table(T)
for reduced, chains in ((False, 1 << 20), (True, 1 << 20), (True, 113664)):
    eng.search_init(chains, seed=1, reduced=reduced, t_start=5e-4, t_end=1e-6, total_rounds=64)
    eng.search_round(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.search_round(32); e1.record(); torch.cuda.synchronize()
    print(os.environ.get("SB_MAX_FUSED"), "reduced", reduced, "chains", chains, f"{e0.elapsed_time(e1)/32*1e3:.1f} us/round", flush=True)
