import sys, torch, numpy as np
sys.path.insert(0, ".")
from saturn_b200.engine import Engine, random_candidates, opt_by_position
from saturn_b200 import synth
eng = Engine(0)
for J, S, B in ((256, 8, 1022976), (64, 6, 4091904), (1024, 1, 262144), (1024, 8, 131072)):
    T, valid = synth.synth_table(J, S, 8, seed=0)
    eng.set_table(T)
    red = S == 1
    opt, prio = random_candidates(eng, B, valid, seed=1)
    op = opt_by_position(opt, prio)
    out = torch.empty(B, dtype=torch.float32, device="cuda")
    for name, o, kw in (("job-indexed", opt, {}), ("by-position", op, {"by_position": True})):
        try:
            for _ in range(3): eng.eval(o, prio, reduced=red, out=out, **kw)
        except RuntimeError as e:
            print(J, S, name, "refused:", str(e)[:60]); continue
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): eng.eval(o, prio, reduced=red, out=out, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"J={J} S={S} B={B} {name:12s} path {eng.last_eval_path()} {ms:.4f} ms  {B/ms*1e3:.3e} cand/s", flush=True)
