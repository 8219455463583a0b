mkdir -p gpurun_out
timeout 300 python scripts/exp_inc2.py 2>&1 | tee gpurun_out/r2c_inc2.md | tail -20
