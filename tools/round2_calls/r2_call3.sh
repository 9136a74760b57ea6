mkdir -p gpurun_out
python -m pytest tests/test_gpu_cascade.py -m gpu -q -k "cache or stress" 2>&1 | grep -v "^$" | tail -150 > gpurun_out/r2_tests2.log
python -m pytest tests/test_gpu_boost.py -m gpu -q -k "stump or weights or cv2" 2>&1 | tail -60 > gpurun_out/r2_tests3.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -c 3000 gpurun_out/r2_bench_a.err
python tools/full_cascade.py 20 5000 10000 48 --save-stages gpurun_out/cfg3_cascade_stages.json > gpurun_out/r2_full_cascade_b.log 2>&1
tail -5 gpurun_out/r2_tests2.log gpurun_out/r2_tests3.log
