mkdir -p gpurun_out
bash tools/verify_world_sizes.sh "8" > gpurun_out/r2_world8_final2.log 2>&1; cat gpurun_out/r2_world8_final2.log | cut -c1-120
bash tools/verify_world_sizes.sh "4" > gpurun_out/r2_world4_final2.log 2>&1; cat gpurun_out/r2_world4_final2.log | cut -c1-120
