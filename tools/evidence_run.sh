#!/bin/bash
# Round evidence on one GPU (run under gpurun): GPU tests, bench lines of both arms, the ncu launch list of the bench command
# and one ncu --set full capture of the level-0 / level-1 band kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.txt
cat gpurun_out/r02_gpu_tests.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_ncu.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_ -c 2 -f -o gpurun_out/r02_band python tools/time_clip.py --fps 30 --steps 1 > gpurun_out/ncu_log.txt 2>&1
cat gpurun_out/r02_bench_n1.json | head -c 1500
