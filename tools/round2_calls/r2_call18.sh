mkdir -p gpurun_out
bash tools/verify_world_sizes.sh "2 4 8" > gpurun_out/r2_world248_final.log 2>&1
cat gpurun_out/r2_world248_final.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
tail -c 300 gpurun_out/r2_bench_n8_final.err; head -c 300 gpurun_out/r2_bench_n8_final.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 4 --steps 10 --warmup 3 --no-mining > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 2 --steps 10 --warmup 3 --no-mining > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
