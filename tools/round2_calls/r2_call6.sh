mkdir -p gpurun_out
python -m pytest tests/test_gpu_boost.py -m gpu -q -k "wide or large" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_tests6.log
tail -5 gpurun_out/r2_tests6.log
python -m pytest tests/test_gpu_cascade.py -m gpu -q -k "cache or oracle or paths" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_tests6b.log
tail -3 gpurun_out/r2_tests6b.log
python tools/full_cascade.py 20 5000 10000 48 > gpurun_out/r2_full_cascade_d.log 2>&1
grep "cascade:\|timer\|mining:" gpurun_out/r2_full_cascade_d.log
