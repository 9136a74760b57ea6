mkdir -p gpurun_out
python -m pytest tests/test_gpu_features.py tests/test_gpu_fullsize.py -m gpu -q -k "sort or sorted or cfg3" 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r2_tests23.log
tail -3 gpurun_out/r2_tests23.log
python bench.py --steps 10 --warmup 3 --no-mining --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_d.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], {k: round(v["ms_per_step"],3) for k,v in d["kernels"]["families"].items()})
PY
