mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_pytest.log | tail -45
timeout 300 python scripts/exp_inc2.py 2>&1 | tee gpurun_out/r2c_inc2.md | tail -30
timeout 600 python scripts/exp_incremental.py > gpurun_out/r02_incremental.md 2> gpurun_out/r02_incremental.err; tail -3 gpurun_out/r02_incremental.err; tail -12 gpurun_out/r02_incremental.md
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench1.log 2> gpurun_out/r2c_bench1.err; echo "rc=$?"; tail -3 gpurun_out/r2c_bench1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c_benchref.log 2> gpurun_out/r2c_benchref.err; head -c 400 gpurun_out/r2c_benchref.log
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 3 -c 1 -o gpurun_out/r02_search_inc -f python scripts/profile_r02.py search > gpurun_out/r02_ncu_search.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 3 -c 1 -o gpurun_out/r02_eval_tiles -f python scripts/profile_r02.py eval > gpurun_out/r02_ncu_eval.log 2>&1
du -sh gpurun_out; ls -la gpurun_out
