mkdir -p gpurun_out
python -m pytest tests/test_gpu_cascade.py tests/test_gpu_boost.py tests/test_gpu_features.py -m gpu -q -k "hog or HOG or matrix" 2>&1 | grep -v "^$" | tail -20 > gpurun_out/r2_tests12.log
tail -4 gpurun_out/r2_tests12.log
python tools/config_bench.py cfg4 > gpurun_out/r2_cfg4.log 2>&1; tail -25 gpurun_out/r2_cfg4.log
python tools/full_cascade.py 20 5000 10000 48 > gpurun_out/r2_full_cascade_e.log 2>&1
grep "cascade:\|timer\|mining:" gpurun_out/r2_full_cascade_e.log
