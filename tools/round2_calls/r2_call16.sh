mkdir -p gpurun_out
python -m pytest tests/test_gpu_cascade.py tests/test_gpu_boost.py tests/test_gpu_features.py -m gpu -q -k "hog or HOG or matrix" 2>&1 | grep -v "^$" | tail -20 > gpurun_out/r2_tests16.log
tail -4 gpurun_out/r2_tests16.log
python tools/config_bench.py cfg4 > gpurun_out/r2_cfg4d.log 2>&1; tail -22 gpurun_out/r2_cfg4d.log
