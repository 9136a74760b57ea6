mkdir -p gpurun_out
for b in 65536 131072 262144 524288 1048576 4194304; do
  echo "== CC_MINE_BATCH=$b" >> gpurun_out/r2_batch_sweep.log
  CC_MINE_BATCH=$b python tools/mine_profile.py 640 3 2>&1 | tail -2 >> gpurun_out/r2_batch_sweep.log
done
cat gpurun_out/r2_batch_sweep.log
python tools/mine_profile.py 320 2 > gpurun_out/r2_plain_mine.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/r2_launches_mine.csv python tools/mine_profile.py 320 2 > gpurun_out/r2_ncu_mine.log 2>&1
