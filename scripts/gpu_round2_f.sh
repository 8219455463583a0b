mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:k_init_population --csv --log-file gpurun_out/r02_init_launches.csv python scripts/profile_r02.py init > /dev/null 2>&1; grep k_init gpurun_out/r02_init_launches.csv | awk -F'","' '{print $5, $NF}'
timeout 600 python scripts/exp_incremental.py > gpurun_out/r02_incremental_bias.md 2> gpurun_out/r02_incremental_bias.err; tail -3 gpurun_out/r02_incremental_bias.err; cat gpurun_out/r02_incremental_bias.md | tail -32
