#!/bin/bash
# one ncu --set full capture of the foveated level-0 band kernel (run under gpurun)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:band_ws -c 1 -f -o gpurun_out/prof_fov python tools/time_clip.py --fps 30 --foveated --steps 1 > gpurun_out/ncu_log.txt 2>&1
tail -2 gpurun_out/ncu_log.txt
