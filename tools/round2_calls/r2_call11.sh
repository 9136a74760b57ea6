mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_tests11.log
tail -4 gpurun_out/r2_tests11.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -c 400 gpurun_out/r2_bench_c.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
head -c 1500 gpurun_out/r2_bench_ref.json
python tools/mine_profile.py 320 2 > gpurun_out/r2_plain_mine.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/r2_launches_mine.csv python tools/mine_profile.py 320 2 > gpurun_out/r2_ncu_mine.log 2>&1
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
