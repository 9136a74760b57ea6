mkdir -p gpurun_out
bash tools/verify_world_sizes.sh "2" > gpurun_out/r2_world2_final2.log 2>&1; cat gpurun_out/r2_world2_final2.log | cut -c1-120
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2_dist_check2.log 2>&1; tail -1 gpurun_out/r2_dist_check2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 1 > gpurun_out/r2_dist_check2_lbp.log 2>&1; tail -1 gpurun_out/r2_dist_check2_lbp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py 2 > gpurun_out/r2_dist_check2_hog.log 2>&1; tail -1 gpurun_out/r2_dist_check2_hog.log
