# final single-GPU batch: tests, bench (both arms), ncu captures, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.log 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench.log") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/r02_bench_ref.log") if l.startswith("{")][-1])
print("value %.4e frac %.4f e2e %.4e cpu %.4e ref %.4e same_config %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], r["value"], d["config"]==r["config"]))
print(json.dumps(d["configs"])[:1500]); print(json.dumps(d["search_round"])[:200]); print(json.dumps(d["solve_api"])[:300])
PY
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 1 -c 1 -o gpurun_out/r02_search_inc -f python scripts/profile_r02.py search > gpurun_out/r02_ncu_search.log 2>&1; tail -2 gpurun_out/r02_ncu_search.log
timeout 600 $NCU --set full -k regex:k_eval_groups -s 2 -c 1 -o gpurun_out/r02_alt_shape -f python - > gpurun_out/r02_ncu_alt.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from saturn_b200.engine import Engine, random_candidates
from saturn_b200.synth import synth_table
eng = Engine(0); T, valid = synth_table(256, 8, 8, seed=0); eng.set_table(T)
B = 148 * 8 * 32 * 4
opt, prio = random_candidates(eng, B, valid, seed=1)
for _ in range(4): eng.eval(opt, prio, alt_shape=True)
torch.cuda.synchronize(); print("ok")
PY
tail -2 gpurun_out/r02_ncu_alt.log
timeout 600 $NCU --metrics gpu__time_duration.sum -k regex:k_init_population --csv --log-file gpurun_out/r02_init_launches.csv python scripts/profile_r02.py init > /dev/null 2>&1; grep k_init gpurun_out/r02_init_launches.csv | awk -F'","' '{print $5, $NF}'
for tool in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $tool python scripts/sanitize.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize run ok|SUMMARY|Error|error" | head -8; done | tee gpurun_out/r02_sanitizer.txt
du -sh gpurun_out
