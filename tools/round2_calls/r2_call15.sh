mkdir -p gpurun_out
python tools/config_bench.py cfg4 7 > gpurun_out/r2_cfg4_7.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 300 -c 500 --csv --log-file gpurun_out/r2_launches_hog.csv python tools/config_bench.py cfg4 7 > gpurun_out/r2_ncu_hog.log 2>&1
tail -12 gpurun_out/r2_cfg4_7.log
