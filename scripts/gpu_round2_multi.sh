# Multi-GPU batch: N = number of visible GPUs (2, 4 or 8)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -q -v 2>&1 | tail -25 | tee gpurun_out/r02_pytest_multi_${N}gpu.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 900 python bench.py --steps 20 --warmup 5 --solve-devices $N > gpurun_out/r02_bench_n1_of${N}.log 2> gpurun_out/r02_bench_n1_of${N}.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_n${n}_of${N}.log 2> gpurun_out/r02_bench_n${n}_of${N}.err
    fi
    echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_bench_n${n}_of${N}.log") if l.startswith("{")][-1])
    print("N=%d value %.4e ms/step %.4f exchange_check %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("exchange_check")))
    print(" per_rank", [(r["rank"], round(r["timed_ms"],3), round(r["kernel_ms_first"],4), round(r["kernel_ms_median"],4)) for r in d["run"]["per_rank"]])
    if d.get("solve_api"): print(" solve_api", json.dumps(d["solve_api"])[:600])
except Exception as e:
    print("parse failed", e)
PY
    tail -2 gpurun_out/r02_bench_n${n}_of${N}.err
  fi
done
timeout 600 python scripts/anneal.py --config C5 --chains 131072 --candidates 1e9 --devices $N > gpurun_out/r02_c5_anneal_${N}dev.md 2> gpurun_out/r02_c5_anneal_${N}dev.err; tail -3 gpurun_out/r02_c5_anneal_${N}dev.err; head -6 gpurun_out/r02_c5_anneal_${N}dev.md; tail -4 gpurun_out/r02_c5_anneal_${N}dev.md
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 scripts/anneal.py --config C5 --chains 131072 --candidates 1e9 > gpurun_out/r02_c5_anneal_${N}gpu.md 2> gpurun_out/r02_c5_anneal_${N}gpu.err; head -4 gpurun_out/r02_c5_anneal_${N}gpu.md
timeout 600 python scripts/anneal.py --config C4 --chains 227328 --candidates 4e9 --devices $N > gpurun_out/r02_c4_anneal_${N}dev.md 2> gpurun_out/r02_c4_anneal_${N}dev.err; head -4 gpurun_out/r02_c4_anneal_${N}dev.md; tail -2 gpurun_out/r02_c4_anneal_${N}dev.md
