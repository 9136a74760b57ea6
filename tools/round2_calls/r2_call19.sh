mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_tests19.log
tail -4 gpurun_out/r2_tests19.log
bash tools/verify_world_sizes.sh "1" > gpurun_out/r2_world1_final.log 2>&1; cat gpurun_out/r2_world1_final.log | cut -c1-100
python tools/full_cascade.py 20 5000 10000 48 > gpurun_out/r2_full_cascade_h.log 2>&1
grep "cascade:\|timer\|mining:" gpurun_out/r2_full_cascade_h.log
