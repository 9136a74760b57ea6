mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_tests17.log
tail -4 gpurun_out/r2_tests17.log
python tools/config_bench.py cfg4 > gpurun_out/r2_cfg4e.log 2>&1; tail -22 gpurun_out/r2_cfg4e.log
