mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r2_n8_gpus.log
bash tools/verify_world_sizes.sh "2 4 8" > gpurun_out/r2_world248.log 2>&1
cat gpurun_out/r2_world248.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/dist_check.py > gpurun_out/r2_dist_check8.log 2>&1
tail -2 gpurun_out/r2_dist_check8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -c 600 gpurun_out/r2_bench_n8.err; head -c 400 gpurun_out/r2_bench_n8.json
