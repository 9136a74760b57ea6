set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.log 2>&1
( pkg-config --modversion opencv4; pkg-config --modversion opencv; find / -name "core_c.h" -not -path "*/proc/*" 2>/dev/null | head; find / -name "libopencv_core*" -not -path "/proc/*" 2>/dev/null | head; ls /usr/include/opencv4 /usr/local/include/opencv4 2>&1 | head -3; which g++ cmake nproc; nproc ) > gpurun_out/r2_opencv_probe.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_tests0.log
bash tools/verify_world_sizes.sh "1" > gpurun_out/r2_world1.log 2>&1
python tools/mine_race.py train gpurun_out/race_cascade.pkl > gpurun_out/r2_race_train.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/mine_race.py scan gpurun_out/race_cascade.pkl 2 1920 1080 > gpurun_out/r2_racecheck_mine.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/mine_race.py scan gpurun_out/race_cascade.pkl 2 1920 1080 > gpurun_out/r2_memcheck_mine.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/mine_race.py stage > gpurun_out/r2_racecheck_stage.log 2>&1
tail -3 gpurun_out/r2_*.log
