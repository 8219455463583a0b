# final multi-GPU batch (N = visible GPUs): smoke, 2-GPU tests, the driver's scaling sequence, one-process solve / anneals
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multi_${N}gpu.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --solve-devices $N > gpurun_out/r02_bench_n1_of${N}.log 2> gpurun_out/r02_bench_n1_of${N}.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_n${n}_of${N}.log 2> gpurun_out/r02_bench_n${n}_of${N}.err
    fi
    echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_n${n}_of${N}.log") if l.startswith("{")][-1])
    print("N=%d value %.4e ms/step %.4f exchange_check %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("exchange_check")))
    if d.get("solve_api"): print(" solve_api", json.dumps(d["solve_api"])[:700])
except Exception as e:
    print("parse failed", e)
PY
    tail -2 gpurun_out/r02_bench_n${n}_of${N}.err | cut -c1-200
  fi
done
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/r02_solve_devices_${N}.md
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from saturn_b200 import Strategy, solve
from saturn_b200 import solver as S
from saturn_b200.synth import synth_table
class T_:
    def __init__(self, n, s): self.name, self.strategies, self.selected_strategy = n, s, None
    def select_strategy(self, s): self.selected_strategy = s
T, valid = synth_table(256, 8, 8, seed=0)
tmin = np.where(valid, T, np.inf).min(axis=1)
tasks = [T_("t%d" % j, {g + 1: Strategy("x", g + 1, {}, float(tmin[j, g])) for g in range(8) if np.isfinite(tmin[j, g])}) for j in range(256)]
print("| devices | rounds | wall ms (median of 5) | candidates | candidates / s | makespan | speed-up |\n|---|---|---|---|---|---|---|")
base = {}
for rounds in (200, 800):
    for n in [d for d in (1, 2, 4, 8) if d <= torch.cuda.device_count()]:
        solve(tasks, None, devices=n, rounds=16)
        ws = []
        for _ in range(5):
            t0 = time.perf_counter(); out = solve(tasks, None, devices=n, rounds=rounds); ws.append(time.perf_counter() - t0)
        w = float(np.median(ws)); c = S.last_stats["candidates"]
        base.setdefault(rounds, c / w)
        print("| %d | %d | %.2f | %.3e | %.3e | %.1f | %.2fx |" % (n, rounds, w * 1e3, c, c / w, out[5], (c / w) / base[rounds]), flush=True)
PY
timeout 300 python scripts/anneal.py --config C5 --chains 131072 --candidates $((N*125000000)) --devices $N > gpurun_out/r02_c5_anneal_${N}dev.md 2> gpurun_out/r02_c5_anneal_${N}dev.err; tail -3 gpurun_out/r02_c5_anneal_${N}dev.err; head -4 gpurun_out/r02_c5_anneal_${N}dev.md; tail -3 gpurun_out/r02_c5_anneal_${N}dev.md
timeout 300 python scripts/anneal.py --config C4 --chains 227328 --candidates $((N*500000000)) --devices $N > gpurun_out/r02_c4_anneal_${N}dev.md 2> gpurun_out/r02_c4_anneal_${N}dev.err; head -4 gpurun_out/r02_c4_anneal_${N}dev.md; tail -3 gpurun_out/r02_c4_anneal_${N}dev.md
