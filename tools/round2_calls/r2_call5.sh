mkdir -p gpurun_out
python -m pytest tests/test_gpu_cascade.py -m gpu -q 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_tests5.log
tail -3 gpurun_out/r2_tests5.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python tools/full_cascade.py 20 5000 10000 48 > gpurun_out/r2_full_cascade_c.log 2>&1
grep "cascade:\|timer\|mining:" gpurun_out/r2_full_cascade_c.log
