# end-of-round validation of HEAD on one B200: smoke, GPU tests, both bench arms (the driver's K/W), clocks under a long run
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.log 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.log 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.err
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 100 > gpurun_out/r02_clocks.csv &
SMI=$!
timeout 300 python bench.py --steps 4000 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02_bench_long.log 2> gpurun_out/r02_bench_long.err
kill $SMI
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench.log") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/r02_bench_ref.log") if l.startswith("{")][-1])
l=json.loads([l for l in open("gpurun_out/r02_bench_long.log") if l.startswith("{")][-1])
print("value %.4e frac %.4f e2e %.4e cpu %.4e ref %.4e e2e/ref %.1f same_config %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], r["value"], d["e2e"]["value"]/r["value"], d["config"]==r["config"]))
print("long run: value %.4e ms/step %.4f clocks %s" % (l["value"], l["ms_per_step"], l["clocks"]))
print(json.dumps(d["search_round"])[:160]); print(json.dumps(d["solve_api"])[:200]); print({k:(v["candidates_per_s"], v["frac"]) for k,v in d["configs"].items()})
PY
tail -3 gpurun_out/r02_clocks.csv
