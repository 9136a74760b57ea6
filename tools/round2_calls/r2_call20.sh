mkdir -p gpurun_out
python -m pytest tests/test_gpu_cascade.py -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_tests20.log
tail -4 gpurun_out/r2_tests20.log
bash tools/verify_world_sizes.sh "1" > gpurun_out/r2_world1_final2.log 2>&1; cat gpurun_out/r2_world1_final2.log | cut -c1-100
python tools/full_cascade.py 20 5000 10000 48 > gpurun_out/r2_full_cascade_i.log 2>&1
grep "cascade:\|timer\|mining:\|stage 1[4-6]" gpurun_out/r2_full_cascade_i.log
