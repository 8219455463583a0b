mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 300 python scripts/exp_inc2.py 2>&1 | tee gpurun_out/r2h_inc2.md | tail -24
timeout 600 python scripts/exp_incremental.py > gpurun_out/r02_incremental_stream.md 2> gpurun_out/r02_incremental_stream.err; tail -3 gpurun_out/r02_incremental_stream.err; tail -34 gpurun_out/r02_incremental_stream.md
