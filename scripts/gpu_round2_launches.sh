# launch list of the bench command at HEAD (one B200): ncu, one metric, no clock control
mkdir -p gpurun_out
timeout 150 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_launches.csv
