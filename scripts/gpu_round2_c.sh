mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log | tail -45
timeout 300 python scripts/exp_inc2.py 2>&1 | tee gpurun_out/r2c_inc2.md | tail -30
timeout 600 python scripts/exp_incremental.py > gpurun_out/r02_incremental.md 2> gpurun_out/r02_incremental.err; tail -3 gpurun_out/r02_incremental.err; tail -12 gpurun_out/r02_incremental.md
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench1.log 2> gpurun_out/r2c_bench1.err; echo "rc=$?"; tail -3 gpurun_out/r2c_bench1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_benchref.log 2> gpurun_out/r2c_benchref.err; head -c 400 gpurun_out/r2c_benchref.log
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 1 -c 1 -o gpurun_out/r02_search_inc -f python scripts/profile_r02.py search > gpurun_out/r02_ncu_search.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 3 -c 1 -o gpurun_out/r02_eval_tiles -f python scripts/profile_r02.py eval > gpurun_out/r02_ncu_eval.log 2>&1
du -sh gpurun_out; ls -la gpurun_out
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1; tail -2 gpurun_out/r02_launches_bench.log | head -c 300
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02_addr_ab.md
# A/B of the look-up address form in the headline kernel (CUDA events, 50 launches each, alternating)
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from saturn_b200.engine import Engine, random_candidates
from saturn_b200.synth import synth_table
eng = Engine(0)
T, valid = synth_table(256, 8, 8, seed=0)
eng.set_table(T)
B = 148 * 8 * 32 * 28
opt, prio = random_candidates(eng, B, valid, seed=1)
out = torch.empty(B, dtype=torch.float32, device="cuda")
res = {}
for rep in range(3):
    for name, plain in (("IMAD addresses (shipped)", False), ("IADD3 + LEA (round 1)", True)):
        for _ in range(3):
            eng.eval(opt, prio, out=out, _plain_addr=plain)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            eng.eval(opt, prio, out=out, _plain_addr=plain)
        e1.record(); torch.cuda.synchronize()
        res.setdefault(name, []).append(e0.elapsed_time(e1) / 50)
print("| look-up address arithmetic | ms / launch (3 x 50 launches, 1,060,864 candidates) | candidates / s |\n|---|---|---|")
for k, v in res.items():
    print("| %s | %s | %.3e |" % (k, ", ".join("%.4f" % x for x in v), B / (min(v) * 1e-3)))
PY
