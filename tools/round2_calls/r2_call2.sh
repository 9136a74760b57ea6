mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests1.log
bash tools/verify_world_sizes.sh "1" > gpurun_out/r2_world1_cache.log 2>&1
python tools/full_cascade.py 20 5000 10000 48 > gpurun_out/r2_full_cascade_cache.log 2>&1
tail -5 gpurun_out/r2_tests1.log; cat gpurun_out/r2_world1_cache.log
