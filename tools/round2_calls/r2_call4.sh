mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -80 > gpurun_out/r2_tests4.log
tail -3 gpurun_out/r2_tests4.log
bash tools/verify_world_sizes.sh "1" > gpurun_out/r2_world1_c.log 2>&1; cat gpurun_out/r2_world1_c.log
