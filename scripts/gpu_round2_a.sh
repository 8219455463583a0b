mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench1.log 2> gpurun_out/r2a_bench1.err; echo "rc=$?" >> gpurun_out/r2a_bench1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_benchref.log 2> gpurun_out/r2a_benchref.err
if [ $(nvidia-smi -L | wc -l) -ge 2 ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2a_bench2.log 2> gpurun_out/r2a_bench2.err; echo "rc=$?" >> gpurun_out/r2a_bench2.err; fi
tail -3 gpurun_out/r2a_pytest.log; head -c 1500 gpurun_out/r2a_bench1.log; tail -5 gpurun_out/r2a_bench1.err; head -c 800 gpurun_out/r2a_bench2.log 2>/dev/null; tail -5 gpurun_out/r2a_bench2.err 2>/dev/null
timeout 600 python scripts/exp_incremental.py > gpurun_out/r02_incremental.md 2> gpurun_out/r02_incremental.err; tail -3 gpurun_out/r02_incremental.err; head -30 gpurun_out/r02_incremental.md
bash scripts/gpu_round2_b.sh > gpurun_out/r2b.log 2>&1; tail -12 gpurun_out/r2b.log
