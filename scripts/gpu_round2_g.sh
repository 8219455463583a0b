mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02g_n1.log 2>gpurun_out/r02g_n1.err
for n in 2 4 8; do if [ $n -le $N ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-e2e > gpurun_out/r02g_n$n.log 2> gpurun_out/r02g_n$n.err
fi; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02g_n*.log")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        pr=d["run"]["per_rank"]
        print(f, "N=%d value %.4e ms/step %.4f check %s kernel median max %.4f first %.4f" % (d["n_gpus"], d["value"], d["ms_per_step"], d.get("exchange_check"), max(r["kernel_ms_median"] for r in pr), max(r["kernel_ms_first"] for r in pr)))
    except Exception as e: print(f, "failed", e)
PY
