# ncu evidence of the round (profiles/r02_*): launch lists of the bench and of a mining sweep, full captures of the top kernels
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mining > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mining > gpurun_out/r2_ncu_bench.log 2>&1
python tools/mine_profile.py 320 2 > gpurun_out/r2_plain_mine.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_mine.csv python tools/mine_profile.py 320 2 > gpurun_out/r2_ncu_mine.log 2>&1
python tools/profile_stage.py 5000 10000 > gpurun_out/r2_plain_stage.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_sort_rows_bucket|k_precalc_haar_sep|k_ss_bound" -c 4 -o gpurun_out/r2_prof_stage python tools/profile_stage.py 5000 10000 > gpurun_out/r2_ncu_stage.log 2>&1
python tools/mine_profile.py 320 1 > gpurun_out/r2_plain_mine1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_lvl_build|k_lvl_cascade" -c 5 -o gpurun_out/r2_prof_mine python tools/mine_profile.py 320 1 > gpurun_out/r2_ncu_mine_full.log 2>&1
cat gpurun_out/r2_plain_mine.log; ls -la gpurun_out/*.ncu-rep | tail -3
