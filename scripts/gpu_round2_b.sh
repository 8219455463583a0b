# Round 2 profiling batch (1 GPU): launch list of the bench command + ncu --set full captures.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 3 -c 1 -o gpurun_out/r02_eval_tiles -f python scripts/profile_r02.py eval > gpurun_out/r02_ncu_eval.log 2>&1
timeout 600 $NCU --set full -k regex:k_eval_tiles -s 3 -c 1 -o gpurun_out/r02_eval_tiles_plain -f python scripts/profile_r02.py eval_plain > gpurun_out/r02_ncu_eval_plain.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_eval_tiles -s 1 -c 1 -o gpurun_out/r02_search_inc -f python scripts/profile_r02.py search > gpurun_out/r02_ncu_search.log 2>&1
timeout 600 $NCU --set full -k regex:k_eval_tiles -s 3 -c 1 -o gpurun_out/r02_search_full -f python scripts/profile_r02.py search_full > gpurun_out/r02_ncu_search_full.log 2>&1
timeout 600 $NCU --set full -k regex:k_search_pos -s 2 -c 1 -o gpurun_out/r02_search_pos -f python scripts/profile_r02.py pos > gpurun_out/r02_ncu_pos.log 2>&1
timeout 600 $NCU --set full -k regex:k_init_population -s 1 -c 1 -o gpurun_out/r02_init -f python scripts/profile_r02.py init > gpurun_out/r02_ncu_init.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out; tail -2 gpurun_out/r02_ncu_*.log
