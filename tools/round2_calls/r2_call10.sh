mkdir -p gpurun_out
python -m pytest tests/test_gpu_createsamples.py -m gpu -q 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_tests10.log
tail -8 gpurun_out/r2_tests10.log
python -m pytest tests/test_gpu_cascade.py -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_tests10b.log
tail -4 gpurun_out/r2_tests10b.log
for b in 4194304 8388608 16777216; do
  echo "== CC_MINE_BATCH=$b" >> gpurun_out/r2_batch_sweep2.log
  CC_MINE_BATCH=$b python tools/mine_profile.py 1280 3 2>&1 | tail -2 >> gpurun_out/r2_batch_sweep2.log
done
cat gpurun_out/r2_batch_sweep2.log
